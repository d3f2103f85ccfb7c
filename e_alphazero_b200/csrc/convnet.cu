// convnet.cu -- the convolutional evaluators of the reference in inference mode:
//   EpistemicResidualAZNet  /root/reference/src/network/resnet.py:41-135   (every pgx env that is not DeepSea / Subleq / MinAtar)
//   EpistemicMinatarAZNet   /root/reference/src/network/minatar.py:11-114
// fp32 on the CUDA cores in ONE defined operation order (the EXACT contract of mlp.cu): a convolution output is the FMA chain
// acc = 0; for kh, kw, ci ascending: acc = fma(x, w, acc); then + b -- hk.Conv2D with SAME padding, NHWC / HWIO; hk.Linear is the chain
// over k ascending; hk.BatchNorm in inference is (x - mean) * (scale * 1/sqrt(var + 1e-5)) + offset with every operation rounded
// separately.  oracle/eaz_oracle.c restates the same order, so the two agree bit for bit; the golden files (the reference's own
// modules executed on the numpy stand-in) pin both to 1e-5.
//
// Three generic kernels, chained per network on the caller's stream (no allocation, no host sync, graph capturable):
//   conv3x3_kernel   one board per CTA: the zero-padded input tile (optionally bool observations, optionally BatchNorm + ReLU applied
//                    while staging -- the pre-activation of BlockV2) lives in shared memory, a thread owns 4 output channels of up to
//                    4 pixels (16 accumulators per weight vector load); epilogue: + bias, BatchNorm (BlockV1), + residual, ReLU
//   dense_kernel     rows x K @ K x N (+ b) with optional per-input BatchNorm + ReLU (the trunk's final BN feeding the 1x1 head
//                    convolutions), per-output BatchNorm, ReLU: serves hk.Linear AND the 1x1 convolutions (rows = B * H * W)
//   convnet_finish_kernel   tanh / softplus, the hash-count novelty of the float32 observation (XXHash, hashes.py:162-229) and
//                    max(novelty, u) (resnet.py:126-128, minatar.py:101-104)
#include "common.cuh"

namespace eaz {

struct BnDev {
  const float *scale, *offset, *mean, *var;  // scale == nullptr: no BatchNorm
};
__device__ __forceinline__ float bn_inv(const BnDev& bn, int c) {  // hk.BatchNorm: inv = scale * rsqrt(var + eps)
  return __fmul_rn(bn.scale[c], __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(bn.var[c], 1e-5f))));
}
__device__ __forceinline__ float bn_apply(const BnDev& bn, int c, float x) {  // (x - mean) * inv + offset
  return __fadd_rn(__fmul_rn(__fsub_rn(x, bn.mean[c]), bn_inv(bn, c)), bn.offset[c]);
}

// ------------------------------------------------------------------------------------------------ 3x3 convolution, SAME padding
// in: fp32 [B,H,W,Cin] or (obs != nullptr) bool [B,H,W,Cin]; w: [3,3,Cin,Cout]; out: fp32 [B,H,W,Cout]; Cout % 4 == 0, Cout / 4 | 256
__global__ void __launch_bounds__(256) conv3x3_kernel(const float* __restrict__ in, const uint8_t* __restrict__ obs, int H, int W, int Cin, int Cout,
                                                      const float* __restrict__ w, const float* __restrict__ bias, BnDev pre, BnDev post,
                                                      const float* __restrict__ residual, int relu_out, float* __restrict__ out) {
  extern __shared__ __align__(16) float s_in[];  // [(H + 2) * (W + 2)][Cin], zero halo
  const int b = blockIdx.x, HW = H * W, W2 = W + 2;
  const size_t base = (size_t)b * HW * Cin;
  for (int i = threadIdx.x; i < (H + 2) * W2 * Cin; i += blockDim.x) {
    const int c = i % Cin, pp = i / Cin, y = pp / W2 - 1, x = pp % W2 - 1;
    float v = 0.0f;
    if (y >= 0 && y < H && x >= 0 && x < W) {
      const size_t g = base + (size_t)(y * W + x) * Cin + c;
      v = obs ? (obs[g] ? 1.0f : 0.0f) : in[g];  // x.astype(float32), resnet.py:69
      if (pre.scale) v = fmaxf(bn_apply(pre, c, v), 0.0f);  // BlockV2: BatchNorm -> relu -> conv (:36-41); the padding stays zero
    }
    s_in[i] = v;
  }
  __syncthreads();
  const int ngrp = Cout >> 2;           // groups of 4 output channels
  const int cg = threadIdx.x % ngrp;    // this thread's group
  const int pl = threadIdx.x / ngrp, npl = blockDim.x / ngrp;
  if (pl >= npl) return;
  for (int p0 = pl; p0 < HW; p0 += 4 * npl) {
    float acc[4][4];
    int off[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int p = min(p0 + j * npl, HW - 1);  // (clamped duplicates are computed and dropped)
      off[j] = ((p / W) * W2 + (p % W)) * Cin;   // top-left corner of the 3x3 window in the padded tile
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[j][q] = 0.0f;
    }
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) {
        const float* wp = w + (size_t)((kh * 3 + kw) * Cin) * Cout + 4 * cg;
        const int tap = (kh * W2 + kw) * Cin;
        for (int ci = 0; ci < Cin; ++ci) {
          const float4 w4 = __ldg(reinterpret_cast<const float4*>(wp + (size_t)ci * Cout));
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float xv = s_in[off[j] + tap + ci];
            acc[j][0] = __fmaf_rn(xv, w4.x, acc[j][0]);
            acc[j][1] = __fmaf_rn(xv, w4.y, acc[j][1]);
            acc[j][2] = __fmaf_rn(xv, w4.z, acc[j][2]);
            acc[j][3] = __fmaf_rn(xv, w4.w, acc[j][3]);
          }
        }
      }
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + 4 * cg));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int p = p0 + j * npl;
      if (p >= HW) break;
      const size_t o = ((size_t)b * HW + p) * Cout + 4 * cg;
      float4 r = make_float4(__fadd_rn(acc[j][0], b4.x), __fadd_rn(acc[j][1], b4.y), __fadd_rn(acc[j][2], b4.z), __fadd_rn(acc[j][3], b4.w));
      if (post.scale)  // BlockV1 / the v1 stem: conv -> BatchNorm (:17-23,73-75)
        r = make_float4(bn_apply(post, 4 * cg, r.x), bn_apply(post, 4 * cg + 1, r.y), bn_apply(post, 4 * cg + 2, r.z), bn_apply(post, 4 * cg + 3, r.w));
      if (residual) {  // x + i (:24,43)
        const float4 i4 = *reinterpret_cast<const float4*>(residual + o);
        r = make_float4(__fadd_rn(r.x, i4.x), __fadd_rn(r.y, i4.y), __fadd_rn(r.z, i4.z), __fadd_rn(r.w, i4.w));
      }
      if (relu_out) r = make_float4(fmaxf(r.x, 0.0f), fmaxf(r.y, 0.0f), fmaxf(r.z, 0.0f), fmaxf(r.w, 0.0f));
      *reinterpret_cast<float4*>(out + o) = r;
    }
  }
}

// ------------------------------------------------------------------------------------------------ rows x K @ K x N
enum { kActNone = 0, kActRelu = 1 };
__global__ void __launch_bounds__(128) dense_kernel(const float* __restrict__ x, int R, int K, int N, int rows_per_cta, const float* __restrict__ w,
                                                    const float* __restrict__ b, BnDev pre, BnDev post, int act, float* __restrict__ y) {
  extern __shared__ __align__(16) float s_x[];  // [rows_per_cta][K | 1] (odd stride: the lanes of a warp may read different rows)
  const int r0 = blockIdx.x * rows_per_cta, nr = min(rows_per_cta, R - r0), ld = K | 1;
  for (int i = threadIdx.x; i < nr * K; i += blockDim.x) {
    float v = x[(size_t)r0 * K + i];
    if (pre.scale) v = fmaxf(bn_apply(pre, i % K, v), 0.0f);
    s_x[(i / K) * ld + i % K] = v;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < nr * N; idx += blockDim.x) {
    const int n = idx % N, r = idx / N;
    const float* xr = s_x + r * ld;
    float acc = 0.0f;
    for (int k = 0; k < K; ++k) acc = __fmaf_rn(xr[k], __ldg(w + (size_t)k * N + n), acc);
    acc = __fadd_rn(acc, b[n]);
    if (post.scale) acc = bn_apply(post, n, acc);
    if (act == kActRelu) acc = fmaxf(acc, 0.0f);
    y[(size_t)(r0 + r) * N + n] = acc;
  }
}

// ------------------------------------------------------------------------------------------------ output transforms + novelty
// XXHash of the float32 observation (obs bool -> 1.0f / 0.0f bit patterns), 4 threads per sample = the 4 lanes of hashes.py:210-229
__global__ void __launch_bounds__(128) convnet_finish_kernel(int kind, const uint8_t* __restrict__ obs, int D, int B, const uint8_t* __restrict__ binary_set,
                                                             int hash_bits, float max_u, float novelty_scale, float local_unc_scale,
                                                             const float* __restrict__ v_raw, const float* __restrict__ u_raw,
                                                             float* __restrict__ value, float* __restrict__ ube, float* __restrict__ novelty) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 2, lane = threadIdx.x & 3;
  const bool on = b < B;
  const int L = D >> 2;
  uint32_t acc = xx_init(lane);
  if (on) {
    const uint8_t* o = obs + (size_t)b * D + (size_t)lane * L;
    for (int i = 0; i < L; ++i) acc = xx_round(acc, o[i] ? EAZ_XX_ONE : 0u);
  }
  const unsigned m = 0xffffffffu;
  const int base = (threadIdx.x & 31) & ~3;
  const uint32_t a0 = __shfl_sync(m, acc, base), a1 = __shfl_sync(m, acc, base + 1), a2 = __shfl_sync(m, acc, base + 2), a3 = __shfl_sync(m, acc, base + 3);
  if (!on || lane != 0) return;
  const uint32_t idx = xx_finish(a0, a1, a2, a3, L, hash_bits);
  const int seen = binary_set ? ((binary_set[idx >> 3] >> (idx & 7u)) & 1) : 0;
  const float nov = __fmul_rn(seen ? 0.0f : 1.0f, novelty_scale);  // (~hash_obj(x)) * max_reward_epistemic_variance
  float v = v_raw[b], u = u_raw[b];
  if (kind == EAZ_CONVNET_RESNET) {
    v = eaz_tanh(v);                                            // resnet.py:102
    u = __fmul_rn(0.5f, __fadd_rn(eaz_tanh(u), 1.0f));          // :114
    u = eaz_max(nov, u);                                        // :126-128 (is_training False)
  } else {
    u = eaz_softplus(u);                                        // minatar.py:91
    u = eaz_max(__fmul_rn(nov, local_unc_scale), u);            // :101-103
    u = eaz_min(eaz_max(u, 0.0f), max_u);                       // :104
  }
  if (value) value[b] = v;
  if (ube) ube[b] = u;
  if (novelty) novelty[b] = nov;
}

// ------------------------------------------------------------------------------------------------ host side
static BnDev bn_of(const eaz_bn& b) { return BnDev{b.scale, b.offset, b.mean, b.var}; }
static const BnDev kNoBn{nullptr, nullptr, nullptr, nullptr};

static int launch_conv3x3(const float* in, const uint8_t* obs, int B, int H, int W, int Cin, int Cout, const eaz_conv& c, BnDev pre, BnDev post,
                          const float* residual, int relu_out, float* out, cudaStream_t st) {
  const size_t smem = (size_t)(H + 2) * (W + 2) * Cin * sizeof(float);
  if (smem > 48 * 1024)
    if (cudaError_t e = cudaFuncSetAttribute(conv3x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); e != cudaSuccess)
      return cuda_fail(e, "conv3x3_kernel shared memory");
  conv3x3_kernel<<<B, 256, smem, st>>>(in, obs, H, W, Cin, Cout, c.w, c.b, pre, post, residual, relu_out, out);
  EAZ_CHECK_LAUNCH("conv3x3_kernel");
  return 0;
}
static int launch_dense(const float* x, int R, int K, int N, const eaz_conv& l, BnDev pre, BnDev post, int act, float* y, cudaStream_t st) {
  int rows = max(1, 128 / N);
  rows = max(rows, 8);
  while (rows > 1 && (size_t)rows * (K | 1) * sizeof(float) > 64 * 1024) rows >>= 1;
  const size_t smem = (size_t)rows * (K | 1) * sizeof(float);
  if (smem > 48 * 1024)
    if (cudaError_t e = cudaFuncSetAttribute(dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024); e != cudaSuccess)
      return cuda_fail(e, "dense_kernel shared memory");
  dense_kernel<<<ceil_div(R, rows), 128, smem, st>>>(x, R, K, N, rows, l.w, l.b, pre, post, act, y);
  EAZ_CHECK_LAUNCH("dense_kernel");
  return 0;
}

static int check_convnet(const eaz_convnet_params* n) {
  EAZ_CHECK_ARG(n != nullptr, "convnet: NULL parameters");
  EAZ_CHECK_ARG(n->kind == EAZ_CONVNET_RESNET || n->kind == EAZ_CONVNET_MINATAR, "convnet: unknown kind %d", n->kind);
  EAZ_CHECK_ARG(n->height >= 1 && n->width >= 1 && n->height <= 32 && n->width <= 32 && n->in_channels >= 1, "convnet: bad observation shape");
  EAZ_CHECK_ARG(n->num_actions >= 1 && n->num_channels >= 4 && n->num_channels % 4 == 0 && 256 % (n->num_channels / 4) == 0 && n->num_channels <= 256,
                "convnet: num_channels %d must be a multiple of 4 that divides 1024", n->num_channels);
  EAZ_CHECK_ARG(n->hidden >= 1 && n->hidden <= 1024, "convnet: bad hidden width");
  EAZ_CHECK_ARG((n->height * n->width * n->in_channels) % 4 == 0, "hash input length %d is not a multiple of 4 (hashes.py:210)",
                n->height * n->width * n->in_channels);
  EAZ_CHECK_ARG(n->hash_bits > 0 && n->hash_bits <= 32, "bits_per_hash %d outside (0, 32] (hashes.py:154)", n->hash_bits);
  if (n->kind == EAZ_CONVNET_RESNET) EAZ_CHECK_ARG(n->num_blocks >= 0 && n->num_blocks <= EAZ_CONVNET_MAX_BLOCKS, "convnet: num_blocks outside [0, 8]");
  const size_t tile = (size_t)(n->height + 2) * (n->width + 2) * (size_t)max(n->in_channels, n->num_channels) * sizeof(float);
  if (tile > 200 * 1024) {
    set_error("convnet: a %dx%d board with %d channels does not fit the one-board-per-CTA convolution", n->height, n->width, max(n->in_channels, n->num_channels));
    return EAZ_ERR_UNSUPPORTED;
  }
  return 0;
}

struct ConvnetLayout {
  size_t act, small, total;  // three activation buffers of `act` floats, then `small` floats of head scratch (bump-allocated)
};
static ConvnetLayout convnet_layout(const eaz_convnet_params* n, int B) {
  ConvnetLayout L;
  const size_t HW = (size_t)n->height * n->width, wide = (size_t)max(n->num_channels, n->hidden);
  L.act = ((size_t)B * HW * n->num_channels + 63) & ~(size_t)63;
  L.small = (size_t)B * (4 * HW * 2 + 10 * wide + 16) + 64 * 24;
  L.total = (3 * L.act + L.small) * sizeof(float);
  return L;
}

}  // namespace eaz

using namespace eaz;

extern "C" {

size_t eaz_convnet_workspace_bytes(const eaz_convnet_params* net, int32_t B) {
  if (check_convnet(net) || B < 1) return 0;
  return convnet_layout(net, B).total;
}

int eaz_convnet_forward(const eaz_convnet_params* net, const uint8_t* observation, int32_t B, float* exploit_logits, float* explore_logits,
                        float* value, float* ube, float* novelty, void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_convnet(net)) return rc;
  EAZ_CHECK_ARG(observation != nullptr && B >= 0, "convnet forward: observation is NULL or negative batch");
  if (B == 0) return 0;
  const ConvnetLayout L = convnet_layout(net, B);
  if (!workspace || workspace_bytes < L.total || ((uintptr_t)workspace & 15)) {
    set_error("convnet workspace: need %zu bytes, 16-byte aligned", L.total);
    return EAZ_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int H = net->height, W = net->width, C0 = net->in_channels, C = net->num_channels, A = net->num_actions, HW = H * W, Hd = net->hidden;
  float* buf[3] = {(float*)workspace, (float*)workspace + L.act, (float*)workspace + 2 * L.act};
  float* sm = (float*)workspace + 3 * L.act;
  size_t used = 0;
  auto take = [&](size_t nfloats) {  // bump allocator over the head scratch (convnet_layout sized it)
    float* p = sm + used;
    used += (nfloats + 63) & ~(size_t)63;
    return p;
  };
  float* v_raw = nullptr;
  float* u_raw = nullptr;
  if (net->kind == EAZ_CONVNET_RESNET) {
    // ---- trunk (resnet.py:69-82)
    const bool v2 = net->resnet_v2 != 0;
    if (int rc = launch_conv3x3(nullptr, observation, B, H, W, C0, C, net->stem, kNoBn, v2 ? kNoBn : bn_of(net->stem_bn), nullptr, v2 ? 0 : 1, buf[0], st))
      return rc;
    int cur = 0;
    for (int i = 0; i < net->num_blocks; ++i) {
      const int t = (cur + 1) % 3, o = (cur + 2) % 3;
      if (v2) {  // BlockV2 (:28-43): bn -> relu -> conv -> bn -> relu -> conv, + input
        if (int rc = launch_conv3x3(buf[cur], nullptr, B, H, W, C, C, net->block_conv[i][0], bn_of(net->block_bn[i][0]), kNoBn, nullptr, 0, buf[t], st)) return rc;
        if (int rc = launch_conv3x3(buf[t], nullptr, B, H, W, C, C, net->block_conv[i][1], bn_of(net->block_bn[i][1]), kNoBn, buf[cur], 0, buf[o], st)) return rc;
      } else {   // BlockV1 (:11-24): conv -> bn -> relu -> conv -> bn, relu(x + input)
        if (int rc = launch_conv3x3(buf[cur], nullptr, B, H, W, C, C, net->block_conv[i][0], kNoBn, bn_of(net->block_bn[i][0]), nullptr, 1, buf[t], st)) return rc;
        if (int rc = launch_conv3x3(buf[t], nullptr, B, H, W, C, C, net->block_conv[i][1], kNoBn, bn_of(net->block_bn[i][1]), buf[cur], 1, buf[o], st)) return rc;
      }
      cur = o;
    }
    // ---- heads (:84-124): 1x1 conv (on relu(bn(x1)) for v2, :80-82) -> bn -> relu -> flatten -> linear [-> relu -> linear]
    const BnDev trunk_bn = v2 ? bn_of(net->final_bn) : kNoBn;
    for (int h = 0; h < 4; ++h) {
      const int k = h < 2 ? 2 : 1;
      float* dst = h == 0 ? exploit_logits : (h == 1 ? explore_logits : nullptr);
      if (h < 2 && !dst) continue;
      float* hc = take((size_t)B * HW * k);
      if (int rc = launch_dense(buf[cur], B * HW, C, k, net->head_conv[h], trunk_bn, bn_of(net->head_bn[h]), kActRelu, hc, st)) return rc;
      if (h < 2) {
        if (int rc = launch_dense(hc, B, HW * k, A, net->head_fc[h], kNoBn, kNoBn, kActNone, dst, st)) return rc;
      } else {
        float* hf = take((size_t)B * C);
        float* ho = take((size_t)B);
        if (int rc = launch_dense(hc, B, HW * k, C, net->head_fc[h], kNoBn, kNoBn, kActRelu, hf, st)) return rc;
        if (int rc = launch_dense(hf, B, C, 1, net->head_out[h], kNoBn, kNoBn, kActNone, ho, st)) return rc;
        (h == 2 ? v_raw : u_raw) = ho;
      }
    }
  } else {
    // ---- minatar.py:55-95: two towers conv -> relu -> flatten -> linear -> relu -> linear -> relu
    float* towers[2];
    for (int tw = 0; tw < 2; ++tw) {
      if (int rc = launch_conv3x3(nullptr, observation, B, H, W, C0, C, net->tower_conv[tw], kNoBn, kNoBn, nullptr, 1, buf[tw], st)) return rc;
      float* f1 = take((size_t)B * Hd);
      towers[tw] = take((size_t)B * Hd);
      if (int rc = launch_dense(buf[tw], B, HW * C, Hd, net->tower_fc[tw][0], kNoBn, kNoBn, kActRelu, f1, st)) return rc;
      if (int rc = launch_dense(f1, B, Hd, Hd, net->tower_fc[tw][1], kNoBn, kNoBn, kActRelu, towers[tw], st)) return rc;
    }
    // heads: [0] main policy (x1), [1] value (x1), [2] exploration policy (x2), [3] ube (x2): Linear(hidden) -> relu -> Linear(out)
    for (int h = 0; h < 4; ++h) {
      const bool policy = (h == 0 || h == 2);
      float* dst = policy ? (h == 0 ? exploit_logits : explore_logits) : take((size_t)B);
      if (policy && !dst) continue;
      float* hh = take((size_t)B * Hd);
      if (int rc = launch_dense(towers[h >> 1], B, Hd, Hd, net->mhead_fc[h][0], kNoBn, kNoBn, kActRelu, hh, st)) return rc;
      if (int rc = launch_dense(hh, B, Hd, policy ? A : 1, net->mhead_fc[h][1], kNoBn, kNoBn, kActNone, dst, st)) return rc;
      if (h == 1) v_raw = dst;
      if (h == 3) u_raw = dst;
    }
  }
  if (used > L.small) {
    set_error("convnet: head scratch overrun (%zu > %zu floats)", used, L.small);
    return EAZ_ERR_WORKSPACE;
  }
  convnet_finish_kernel<<<ceil_div(B * 4, 128), 128, 0, st>>>(net->kind, observation, HW * C0, B, net->binary_set, net->hash_bits, net->max_u,
                                                              net->novelty_scale, net->local_unc_scale, v_raw, u_raw, value, ube, novelty);
  EAZ_CHECK_LAUNCH("convnet_finish_kernel");
  return 0;
}

}  // extern "C"

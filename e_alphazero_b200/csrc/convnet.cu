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
#include "mlp.cuh"
#include "umma.cuh"

namespace eaz {
using namespace umma;

struct BnDev {
  const float *scale, *offset, *mean, *var;  // scale == nullptr: no BatchNorm
};
__device__ __forceinline__ float bn_inv(const BnDev& bn, int c) {  // hk.BatchNorm: inv = scale * rsqrt(var + eps)
  return __fmul_rn(bn.scale[c], __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(bn.var[c], 1e-5f))));
}
__device__ __forceinline__ float bn_apply(const BnDev& bn, int c, float x) {  // (x - mean) * inv + offset
  return __fadd_rn(__fmul_rn(__fsub_rn(x, bn.mean[c]), bn_inv(bn, c)), bn.offset[c]);
}

// ---- TENSOR mode operand format: per pixel [hi 64 halves | lo 64 halves] of relu(bn_next(x)) * 16 (the scaled 3xFP16 split of mlp.cuh)
constexpr int kAct16Bytes = 256;
constexpr int kAct16Stride = 272;  // bytes between pixels, in global memory as in the staged tile: thread-per-row 16-byte shared-memory reads
                                   // are conflict-free at this stride, and a CTA's whole input tile is ONE contiguous range (bulk copy)
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) { return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16); }
__device__ __forceinline__ void act16_split(const BnDev& bn, int c, float x, __half& hi, __half& lo, uint32_t* num_flags) {
  float v = __fmul_rn(fmaxf(bn_apply(bn, c, x), 0.0f), kActScale);
  if (!(v <= 65504.0f)) {  // range guard: clamped AND reported (eaz_convnet_numeric_status)
    atomicOr(num_flags, kNumActSaturated);
    v = 65504.0f;
  }
  split_f16(v, hi, lo);
}

// ------------------------------------------------------------------------------------------------ 3x3 convolution, SAME padding
// in: fp32 [B,H,W,Cin] or (obs != nullptr) bool [B,H,W,Cin]; w: [3,3,Cin,Cout]; out: fp32 [B,H,W,Cout]; Cout % 4 == 0, Cout / 4 | 256
__global__ void __launch_bounds__(256, 4) conv3x3_kernel(const float* __restrict__ in, const uint8_t* __restrict__ obs, int H, int W, int Cin, int Cout,
                                                      const float* __restrict__ w, const float* __restrict__ bias, BnDev pre, BnDev post,
                                                      const float* __restrict__ residual, int relu_out, float* __restrict__ out,
                                                      uint8_t* __restrict__ act16_out, BnDev next_bn, uint32_t* num_flags) {
  extern __shared__ __align__(16) float s_in[];  // [(H + 2) * (W + 2)][Cin], zero halo
  const int b = blockIdx.x, HW = H * W, W2 = W + 2;
  float* const s_pre = s_in + (H + 2) * W2 * Cin;  // [3][Cin]: the pre-activation BatchNorm per input channel: inv, mean, offset
  const size_t base = (size_t)b * HW * Cin;
  if (pre.scale) {
    for (int c = threadIdx.x; c < Cin; c += blockDim.x) {
      s_pre[c] = bn_inv(pre, c);
      s_pre[Cin + c] = pre.mean[c];
      s_pre[2 * Cin + c] = pre.offset[c];
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < (H + 2) * W2 * Cin; i += blockDim.x) {
    const int c = i % Cin, pp = i / Cin, y = pp / W2 - 1, x = pp % W2 - 1;
    float v = 0.0f;
    if (y >= 0 && y < H && x >= 0 && x < W) {
      const size_t g = base + (size_t)(y * W + x) * Cin + c;
      v = obs ? (obs[g] ? 1.0f : 0.0f) : in[g];  // x.astype(float32), resnet.py:69
      // BlockV2: BatchNorm -> relu -> conv (:36-41), the operations of bn_apply; the padding stays zero
      if (pre.scale) v = fmaxf(__fadd_rn(__fmul_rn(__fsub_rn(v, s_pre[Cin + c]), s_pre[c]), s_pre[2 * Cin + c]), 0.0f);
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
      if (act16_out) {  // TENSOR mode: the next convolution's A operand, pre-activated and split (conv_tensor_kernel below)
        const float rr[4] = {r.x, r.y, r.z, r.w};
        __half hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) act16_split(next_bn, 4 * cg + q, rr[q], hi[q], lo[q], num_flags);
        uint8_t* a16 = act16_out + ((size_t)b * HW + p) * kAct16Stride + (size_t)(4 * cg) * 2;
        *reinterpret_cast<uint2*>(a16) = make_uint2(pack_h2(hi[0], hi[1]), pack_h2(hi[2], hi[3]));
        *reinterpret_cast<uint2*>(a16 + kAct16Bytes / 2) = make_uint2(pack_h2(lo[0], lo[1]), pack_h2(lo[2], lo[3]));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ 3x3 convolution 64 -> 64 on tcgen05
// Implicit GEMM: one CTA = 256 consecutive output pixels (two M = 128 groups sharing every weight chunk) x 64 output channels (N),
// K = 9 taps x 64 input channels = 18 chunks of 32.
//   * the pre-activated fp16 hi / lo rows of the CTA's pixels PLUS a halo of W + 1 pixels on either side (everything a valid tap can
//     touch; written once by the producing layer's epilogue) are staged in shared memory ONCE by a few large bulk async copies (the
//     operand buffer keeps the same 272-byte pixel stride that makes thread-per-row 16-byte reads conflict-free) -- the first version gathered every tap from L2 (9 x re-read, 0.9 GB per
//     layer at 4096 boards: 242 us per layer);
//   * 8 producer warps (thread = row = TMEM lane) copy the shifted pixel's 64 B hi + 64 B lo of the chunk from the staged tile straight
//     into TENSOR MEMORY (tcgen05.st; zeros for taps outside the board): the A operand never crosses the shared-memory port twice,
//     2-stage ring of 32 columns per group;
//   * weight chunk images (tile_weights.cu: [hi 64x32 | lo 64x32], 8 KB) by 1-D bulk async copies through a 4-stage ring, each used by
//     both groups; one elected lane issues tcgen05.mma kind::f16 M=128 N=64 K=16 with A from TMEM, three split products per K-step,
//     fp32 accumulators in 2 x 64 TMEM columns;
//   * epilogue (the same 8 warps, out of TMEM): x unscale + bias, + residual, fp32 store, and the NEXT layer's operand:
//     relu(bn_next(.)) x 16, clamp + flag, hi / lo split.
namespace ct {
constexpr int kTM = 128, kGroups = 2, kPix = kTM * kGroups, kC = 64, kCK = 32, kChunks = 9 * kC / kCK;
constexpr int kStagesA = 2, kStagesB = 4;
constexpr int kBHalf = kC * kCK * 2, kBStage = 2 * kBHalf;  // 4 KB hi + 4 KB lo
constexpr int kSBO = (kCK * 2 / 16) * kCoreBytes;           // 512 B
constexpr int kRowStride = kAct16Stride;                     // staged pixel row: [hi 128 B | lo 128 B] + 16 B
constexpr int kTmemCols = 256, kTmemA = kGroups * kC;        // D: group g at 64 g; A: stage s, group g at kTmemA + (2 s + g) * 32: [hi 16 | lo 16]
struct Smem {
  uint64_t full_a[kStagesA], empty_a[kStagesA], full_b[kStagesB], empty_b[kStagesB], acc_done, staged;
  uint32_t tmem_base;
  alignas(16) float bias[kC];
  alignas(16) float nb_inv[kC], nb_mean[kC], nb_off[kC];  // the consumer's BatchNorm, per channel (same operations as bn_apply)
};
__host__ __device__ inline int staged_pixels(int W) { return kPix + 2 * (W + 1); }
__host__ __device__ inline size_t smem_bytes(int W) { return 1024 + (size_t)kStagesB * kBStage + (size_t)staged_pixels(W) * kRowStride + sizeof(Smem) + 16; }
}  // namespace ct

struct ConvTensorArgs {
  const uint8_t* act16_in;  // [P][hi 64 halves | lo 64 halves]
  const uint8_t* wimg;      // 18 chunk images
  const float* wscale;      // the power-of-two scale the image carries
  const float* bias;        // [64]
  const float* residual;    // fp32 [P][64] or nullptr
  float* out_raw;           // fp32 [P][64] or nullptr
  uint8_t* act16_out;       // or nullptr
  BnDev next_bn;            // pre-activation of the consumer of act16_out
  uint32_t* num_flags;
  int P, H, W;              // P = B * H * W pixels
};

__global__ void __launch_bounds__(320, 2) conv_tensor_kernel(const ConvTensorArgs a) {
  using namespace ct;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = smem;
  uint8_t* sX = smem + kStagesB * kBStage;  // staged pixels [first, first + n_staged), kRowStride bytes each
  const int nst = staged_pixels(a.W);
  Smem* sh = reinterpret_cast<Smem*>(sX + (size_t)nst * kRowStride);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P0 = blockIdx.x * kPix, HW = a.H * a.W;
  const int first = P0 - (a.W + 1);  // global pixel index of staged row 0
  const int lo_pix = max(first, 0), hi_pix = min(first + nst, a.P);  // the rows that exist

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStagesA; ++s) {
      mbar_init(&sh->full_a[s], 8);  // one elected arrive per producer warp
      mbar_init(&sh->empty_a[s], 1);
    }
    for (int s = 0; s < kStagesB; ++s) {
      mbar_init(&sh->full_b[s], 1);
      mbar_init(&sh->empty_b[s], 1);
    }
    mbar_init(&sh->acc_done, 1);
    mbar_init(&sh->staged, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < kC) {
    sh->bias[threadIdx.x] = __ldg(a.bias + threadIdx.x);
    if (a.act16_out) {
      sh->nb_inv[threadIdx.x] = bn_inv(a.next_bn, threadIdx.x);
      sh->nb_mean[threadIdx.x] = a.next_bn.mean[threadIdx.x];
      sh->nb_off[threadIdx.x] = a.next_bn.offset[threadIdx.x];
    }
  }
  if (warp == 8) {
    tmem_alloc(&sh->tmem_base, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh->tmem_base;

  if (warp == 9) {
    // ================= copy warp: the staged pixel rows (once), then the weight chunk ring =================
    if (lane == 0) {
      // the staged rows are one contiguous range of the (272-byte strided) operand buffer: a few large bulk copies
      const uint32_t total = (uint32_t)(hi_pix - lo_pix) * kRowStride;
      mbar_arrive_expect_tx(&sh->staged, total);
      for (uint32_t off = 0; off < total; off += 32768u)
        bulk_g2s(sX + (size_t)(lo_pix - first) * kRowStride + off, a.act16_in + (size_t)lo_pix * kAct16Stride + off, min(32768u, total - off), &sh->staged);
    }
    if (lane == 0) {
      for (int c = 0; c < kChunks; ++c) {
        const int s = c % kStagesB;
        if (c >= kStagesB) mbar_wait(&sh->empty_b[s], ((c / kStagesB) & 1) ^ 1);
        mbar_arrive_expect_tx(&sh->full_b[s], (uint32_t)kBStage);
        bulk_g2s(sB + s * kBStage, a.wimg + (size_t)c * kBStage, kBStage, &sh->full_b[s]);
      }
    }
    __syncwarp();
  } else if (warp == 8) {
    // ================= MMA-issue warp: warp-uniform loop, one elected lane issues (A from tensor memory) =================
    const uint32_t desc_hi = (uint32_t)(kSBO >> 4) | (1u << 14);  // SBO [32,46) + version=1 [46,48)
    const uint32_t lbo_bits = (uint32_t)(kCoreBytes >> 4) << 16;  // LBO [16,30)
    auto mk = [&](uint32_t lo) { return ((uint64_t)desc_hi << 32) | (uint64_t)lo; };
    const uint32_t idesc = idesc_f16(kTM, kC);
    const uint32_t b_base = ((smem_u32(sB) & 0x3FFFFu) >> 4) | lbo_bits;
#pragma unroll 1
    for (int c = 0; c < kChunks; ++c) {
      const int sa = c % kStagesA, sb = c % kStagesB;
      mbar_wait(&sh->full_b[sb], (c / kStagesB) & 1);
      mbar_wait(&sh->full_a[sa], (c / kStagesA) & 1);
      tc_fence_after();
      const uint32_t bl = b_base + (uint32_t)((sb * kBStage) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int g = 0; g < kGroups; ++g) {
          const uint32_t d = tmem + (uint32_t)(g * kC);
          const uint32_t a_hi = tmem + (uint32_t)(kTmemA + (2 * sa + g) * 32), a_lo = a_hi + 16;
#pragma unroll
          for (int j = 0; j < kCK / 16; ++j) {
            const uint32_t o = (uint32_t)(j * kKStepBytes) >> 4;
            mma_f16_ts(d, a_hi + 8 * j, mk(bl + o), idesc, (c | j) != 0);
            mma_f16_ts(d, a_hi + 8 * j, mk(bl + (kBHalf >> 4) + o), idesc, 1);
            mma_f16_ts(d, a_lo + 8 * j, mk(bl + o), idesc, 1);
          }
        }
        mma_commit(&sh->empty_a[sa]);
        mma_commit(&sh->empty_b[sb]);
        if (c == kChunks - 1) mma_commit(&sh->acc_done);
      }
      __syncwarp();
    }
  } else {
    // ================= workers: thread = output pixel = TMEM lane; warps 0-3 group 0, warps 4-7 group 1 =================
    const int g = warp >> 2, wq = warp & 3;
    const int row = 32 * wq + lane, Pr = P0 + g * kTM + row;
    const bool live = Pr < a.P;
    const int p = live ? Pr % HW : 0;
    const int y = live ? p / a.W : -(1 << 20), x = p % a.W;  // dead rows: every tap falls outside the board
    const uint8_t* const xrow = sX + (size_t)(Pr - first) * kRowStride;  // this pixel's own staged row
    const uint32_t tlane = tmem + ((uint32_t)(32 * wq) << 16);
    mbar_wait_warp(&sh->staged, 0);
#pragma unroll 1
    for (int c = 0; c < kChunks; ++c) {
      const int sa = c % kStagesA;
      const int tap = c >> 1, kh = tap / 3 - 1, kw = tap % 3 - 1, ci0 = (c & 1) * kCK;
      const int yy = y + kh, xx = x + kw;
      const bool in = yy >= 0 && yy < a.H && xx >= 0 && xx < a.W;  // SAME padding: zeros outside the board
      uint4 h[4], l[4];
      if (in) {
        const uint4* src = reinterpret_cast<const uint4*>(xrow + (kh * a.W + kw) * kRowStride + ci0 * 2);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          h[q] = src[q];
          l[q] = src[(kAct16Bytes / 2) / 16 + q];
        }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) h[q] = l[q] = make_uint4(0u, 0u, 0u, 0u);
      }
      if (c >= kStagesA) mbar_wait_warp(&sh->empty_a[sa], ((c / kStagesA) & 1) ^ 1);  // the MMAs of chunk c - 2 have read this stage
      tc_fence_after();
      const uint32_t ta = tlane + (uint32_t)(kTmemA + (2 * sa + g) * 32);
      tmem_st8(ta, h[0], h[1]);
      tmem_st8(ta + 8, h[2], h[3]);
      tmem_st8(ta + 16, l[0], l[1]);
      tmem_st8(ta + 24, l[2], l[3]);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->full_a[sa]);
    }

    // ---- epilogue: this thread's pixel, 64 channels in two halves of 32
    const float unscale = 1.0f / (kActScale * __ldg(a.wscale));  // exact: powers of two
    mbar_wait_warp(&sh->acc_done, 0);
    tc_fence_after();
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      const int cbase = half * 32;
      uint32_t ra[16], rb[16];
      tmem_ld16(tlane + (uint32_t)(g * kC + cbase), ra);
      tmem_ld16(tlane + (uint32_t)(g * kC + cbase + 16), rb);
      tmem_ld_wait();
      if (live) {
        float val[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          val[i] = __fmaf_rn(__uint_as_float(ra[i]), unscale, sh->bias[cbase + i]);
          val[16 + i] = __fmaf_rn(__uint_as_float(rb[i]), unscale, sh->bias[cbase + 16 + i]);
        }
        const size_t o = (size_t)Pr * kC + cbase;
        if (a.residual) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 r4 = *reinterpret_cast<const float4*>(a.residual + o + 4 * q);
            val[4 * q] = __fadd_rn(val[4 * q], r4.x); val[4 * q + 1] = __fadd_rn(val[4 * q + 1], r4.y);
            val[4 * q + 2] = __fadd_rn(val[4 * q + 2], r4.z); val[4 * q + 3] = __fadd_rn(val[4 * q + 3], r4.w);
          }
        }
        if (a.out_raw) {
#pragma unroll
          for (int q = 0; q < 8; ++q) *reinterpret_cast<float4*>(a.out_raw + o + 4 * q) = make_float4(val[4 * q], val[4 * q + 1], val[4 * q + 2], val[4 * q + 3]);
        }
        if (a.act16_out) {
          uint8_t* a16 = a.act16_out + (size_t)Pr * kAct16Stride + (size_t)cbase * 2;
#pragma unroll
          for (int q = 0; q < 4; ++q) {  // 8 channels = 16 bytes of hi and of lo per store
            uint32_t hh[4], ll[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              __half hl[4];  // h0, l0, h1, l1
#pragma unroll
              for (int t = 0; t < 2; ++t) {
                const int ch = cbase + 8 * q + 2 * e + t;
                // relu(bn(x)) * 16 with the per-channel terms from shared memory: the operations of bn_apply / act16_split
                float v = __fmul_rn(fmaxf(__fadd_rn(__fmul_rn(__fsub_rn(val[8 * q + 2 * e + t], sh->nb_mean[ch]), sh->nb_inv[ch]), sh->nb_off[ch]), 0.0f), kActScale);
                if (!(v <= 65504.0f)) {
                  atomicOr(a.num_flags, kNumActSaturated);
                  v = 65504.0f;
                }
                split_f16(v, hl[2 * t], hl[2 * t + 1]);
              }
              hh[e] = pack_h2(hl[0], hl[2]);
              ll[e] = pack_h2(hl[1], hl[3]);
            }
            *reinterpret_cast<uint4*>(a16 + 16 * q) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
            *reinterpret_cast<uint4*>(a16 + kAct16Bytes / 2 + 16 * q) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, ct::kTmemCols);
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

// ------------------------------------------------------------------------------------------------ the four 1x1 head convolutions, one pass
// resnet.py:84-124: every head starts with hk.Conv2D(k, kernel_shape=1) -> BatchNorm -> relu on the trunk output (k = 2, 2, 1, 1).  One
// thread per pixel reads its 64 trunk channels once (applying the trunk's final BatchNorm + relu for v2, :80-82) and produces all six
// head channels; each is the same FMA chain over ci ascending as dense_kernel, so the results are bit-identical to four dense passes
// (which read the 64-channel activations four times: 416 us of a 2 ms forward at 4096 boards).
struct HeadConvArgs {
  const float* w[4];  // [C][k]
  const float* b[4];
  BnDev bn[4];
  float* out[4];      // [P][k]; nullptr = head not requested
  int k[4];
};
__global__ void __launch_bounds__(256) head_conv_kernel(const float* __restrict__ x, int P, int C, BnDev trunk_bn, HeadConvArgs hc) {
  extern __shared__ __align__(16) float s_hc[];  // [3][C] trunk BN (inv, mean, offset) | [6][C] weights, output-major
  float* s_w = s_hc + 3 * C;
  __shared__ float s_post[4][6];  // per head channel: conv bias, BN inv, mean, offset
  int col0[4], ncol = 0;
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    col0[h] = ncol;
    ncol += hc.k[h];
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    if (trunk_bn.scale) {
      s_hc[i] = bn_inv(trunk_bn, i);
      s_hc[C + i] = trunk_bn.mean[i];
      s_hc[2 * C + i] = trunk_bn.offset[i];
    }
#pragma unroll
    for (int h = 0; h < 4; ++h)
      for (int j = 0; j < hc.k[h]; ++j) s_w[(col0[h] + j) * C + i] = hc.w[h][(size_t)i * hc.k[h] + j];
  }
  if (threadIdx.x < 6) {
    int h = 0;
    while (h < 3 && (int)threadIdx.x >= col0[h + 1]) ++h;
    const int j = threadIdx.x - col0[h];
    if (j < hc.k[h]) {
      s_post[0][threadIdx.x] = hc.b[h][j];
      s_post[1][threadIdx.x] = bn_inv(hc.bn[h], j);
      s_post[2][threadIdx.x] = hc.bn[h].mean[j];
      s_post[3][threadIdx.x] = hc.bn[h].offset[j];
    }
  }
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  float acc[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)p * C);
  for (int q = 0; q < C / 4; ++q) {
    const float4 v4 = xr[q];
    float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int ci = 4 * q + e;
      float xv = v[e];
      if (trunk_bn.scale) xv = fmaxf(__fadd_rn(__fmul_rn(__fsub_rn(xv, s_hc[C + ci]), s_hc[ci]), s_hc[2 * C + ci]), 0.0f);
#pragma unroll
      for (int o = 0; o < 6; ++o) acc[o] = __fmaf_rn(xv, s_w[o * C + ci], acc[o]);
    }
  }
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    if (!hc.out[h]) continue;
    for (int j = 0; j < hc.k[h]; ++j) {
      const int o = col0[h] + j;
      float r = 0.0f;
#pragma unroll
      for (int t = 0; t < 6; ++t)
        if (t == o) r = acc[t];
      r = __fadd_rn(r, s_post[0][o]);
      r = __fadd_rn(__fmul_rn(__fsub_rn(r, s_post[2][o]), s_post[1][o]), s_post[3][o]);  // bn_apply
      hc.out[h][(size_t)p * hc.k[h] + j] = fmaxf(r, 0.0f);
    }
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
                          const float* residual, int relu_out, float* out, cudaStream_t st, uint8_t* act16_out = nullptr, BnDev next_bn = BnDev{},
                          uint32_t* num_flags = nullptr) {
  const size_t smem = ((size_t)(H + 2) * (W + 2) + 3) * Cin * sizeof(float);
  if (smem > 48 * 1024)
    if (cudaError_t e = cudaFuncSetAttribute(conv3x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); e != cudaSuccess)
      return cuda_fail(e, "conv3x3_kernel shared memory");
  conv3x3_kernel<<<B, 256, smem, st>>>(in, obs, H, W, Cin, Cout, c.w, c.b, pre, post, residual, relu_out, out, act16_out, next_bn, num_flags);
  EAZ_CHECK_LAUNCH("conv3x3_kernel");
  return 0;
}
static int launch_dense(const float* x, int R, int K, int N, const eaz_conv& l, BnDev pre, BnDev post, int act, float* y, cudaStream_t st) {
  int rows = max(1, 128 / N);
  rows = max(rows, 8);
  while (rows > 8 && ceil_div(R, rows) < 296) rows >>= 1;  // small problems: at least two CTAs per SM
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
  EAZ_CHECK_ARG(n->height >= 1 && n->width >= 1 && n->height <= 32 && n->width <= 32 && n->in_channels >= 1 && n->in_channels <= 256, "convnet: bad observation shape");
  EAZ_CHECK_ARG(n->num_actions >= 1 && n->num_channels >= 4 && n->num_channels % 4 == 0 && 256 % (n->num_channels / 4) == 0 && n->num_channels <= 256,
                "convnet: num_channels %d must be a multiple of 4 that divides 1024", n->num_channels);
  EAZ_CHECK_ARG(n->hidden >= 1 && n->hidden <= 1024, "convnet: bad hidden width");
  EAZ_CHECK_ARG((n->height * n->width * n->in_channels) % 4 == 0, "hash input length %d is not a multiple of 4 (hashes.py:210)",
                n->height * n->width * n->in_channels);
  EAZ_CHECK_ARG(n->hash_bits > 0 && n->hash_bits <= 32, "bits_per_hash %d outside (0, 32] (hashes.py:154)", n->hash_bits);
  if (n->kind == EAZ_CONVNET_RESNET) EAZ_CHECK_ARG(n->num_blocks >= 0 && n->num_blocks <= EAZ_CONVNET_MAX_BLOCKS, "convnet: num_blocks outside [0, 8]");
  EAZ_CHECK_ARG(n->mlp_mode == EAZ_MLP_EXACT || n->mlp_mode == EAZ_MLP_TENSOR, "convnet: unknown mlp_mode %d", n->mlp_mode);
  if (n->mlp_mode == EAZ_MLP_TENSOR && !(n->kind == EAZ_CONVNET_RESNET && n->resnet_v2 && n->num_channels == 64)) {
    set_error("convnet: mlp_mode TENSOR covers EpistemicResidualAZNet v2 with 64 channels (the reference configuration); use EXACT for this network");
    return EAZ_ERR_UNSUPPORTED;
  }
  const size_t tile = (size_t)(n->height + 2) * (n->width + 2) * (size_t)max(n->in_channels, n->num_channels) * sizeof(float);
  if (tile > 200 * 1024) {
    set_error("convnet: a %dx%d board with %d channels does not fit the one-board-per-CTA convolution", n->height, n->width, max(n->in_channels, n->num_channels));
    return EAZ_ERR_UNSUPPORTED;
  }
  return 0;
}

struct ConvnetLayout {
  size_t act, small, total;  // three activation buffers of `act` floats, then `small` floats of head scratch (bump-allocated)
  // TENSOR mode, behind them (byte offsets from the workspace base): two act16 buffers, the weight images, zero page + status blocks
  size_t act16_off, act16_bytes, wimg_off, wimg_bytes, zero_off, status_off;
};
constexpr size_t kConvImgBytes = (size_t)ct::kChunks * ct::kBStage;  // 147 456 B per 64 -> 64 convolution
static bool convnet_tensor(const eaz_convnet_params* n) { return n->mlp_mode == EAZ_MLP_TENSOR; }
static ConvnetLayout convnet_layout(const eaz_convnet_params* n, int B) {
  ConvnetLayout L{};
  const size_t HW = (size_t)n->height * n->width, wide = (size_t)max(n->num_channels, n->hidden);
  L.act = ((size_t)B * HW * n->num_channels + 63) & ~(size_t)63;
  L.small = (size_t)B * (4 * HW * 2 + 10 * wide + 16) + 64 * 24;
  L.total = (3 * L.act + L.small) * sizeof(float);
  if (convnet_tensor(n)) {
    L.total = (L.total + 255) & ~(size_t)255;
    L.act16_off = L.total;
    L.act16_bytes = ((size_t)B * HW * kAct16Stride + 255) & ~(size_t)255;
    L.wimg_off = L.act16_off + 2 * L.act16_bytes;
    L.wimg_bytes = (size_t)2 * EAZ_CONVNET_MAX_BLOCKS * kConvImgBytes;
    L.zero_off = L.wimg_off + L.wimg_bytes;
    L.status_off = L.zero_off + 256;
    L.total = L.status_off + 2 * sizeof(NumStatus);
  }
  return L;
}

static int launch_conv_tensor(const ConvTensorArgs& a, cudaStream_t st) {
  const size_t smem = ct::smem_bytes(a.W);
  if (cudaError_t e = cudaFuncSetAttribute(conv_tensor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ct::smem_bytes(32)); e != cudaSuccess)
    return cuda_fail(e, "conv_tensor_kernel shared memory");
  conv_tensor_kernel<<<ceil_div(a.P, ct::kPix), 320, smem, st>>>(a);
  EAZ_CHECK_LAUNCH("conv_tensor_kernel");
  return 0;
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
    int cur = 0;
    if (convnet_tensor(net) && net->num_blocks > 0) {
      // ---- TENSOR mode: the 2 x num_blocks 64 -> 64 convolutions as tcgen05 implicit GEMMs (conv_tensor_kernel)
      uint8_t* wsb = (uint8_t*)workspace;
      uint8_t* a16[2] = {wsb + L.act16_off, wsb + L.act16_off + L.act16_bytes};
      NumStatus* ns = reinterpret_cast<NumStatus*>(wsb + L.status_off);
      const int nconv = 2 * net->num_blocks;
      for (int g = 0; g * 12 < nconv; ++g) {  // per-tensor power-of-two scales from max |w| (tile_weights.cu), 12 tensors per status block
        const float* wl[12];
        int nl[12];
        const int cnt = min(12, nconv - 12 * g);
        for (int m = 0; m < cnt; ++m) {
          wl[m] = net->block_conv[(12 * g + m) >> 1][(12 * g + m) & 1].w;
          nl[m] = 9 * C * C;
        }
        if (int rc = launch_weight_scales_list(wl, nl, cnt, ns + g, st)) return rc;
      }
      auto scale_of = [&](int k) { return &ns[k / 12].wscale[(k % 12) / 3][(k % 12) % 3]; };
      for (int k = 0; k < nconv; ++k)
        if (int rc = launch_tile_weights_f16(net->block_conv[k >> 1][k & 1].w, 9 * C, C, 9 * C, C, scale_of(k), wsb + L.wimg_off + (size_t)k * kConvImgBytes, st))
          return rc;
      uint32_t* flags = &ns[0].flags;
      // stem (fp32: K = 9 x in_channels is tiny) -> raw x_0 and the first block's operand relu(bn_0,0(x_0))
      if (int rc = launch_conv3x3(nullptr, observation, B, H, W, C0, C, net->stem, kNoBn, kNoBn, nullptr, 0, buf[0], st, a16[0], bn_of(net->block_bn[0][0]), flags))
        return rc;
      for (int i = 0; i < net->num_blocks; ++i) {  // BlockV2: x_{i+1} = conv2(relu(bn2(conv1(relu(bn1(x_i)))))) + x_i
        const int o = cur ^ 1;
        ConvTensorArgs c1{a16[0], wsb + L.wimg_off + (size_t)(2 * i) * kConvImgBytes, scale_of(2 * i), net->block_conv[i][0].b, nullptr, nullptr,
                          a16[1], bn_of(net->block_bn[i][1]), flags, B * HW, H, W};
        if (int rc = launch_conv_tensor(c1, st)) return rc;
        const bool last = i + 1 == net->num_blocks;
        ConvTensorArgs c2{a16[1], wsb + L.wimg_off + (size_t)(2 * i + 1) * kConvImgBytes, scale_of(2 * i + 1), net->block_conv[i][1].b, buf[cur],
                          buf[o], last ? nullptr : a16[0], last ? kNoBn : bn_of(net->block_bn[i + 1][0]), flags, B * HW, H, W};
        if (int rc = launch_conv_tensor(c2, st)) return rc;
        cur = o;
      }
    } else {
    if (int rc = launch_conv3x3(nullptr, observation, B, H, W, C0, C, net->stem, kNoBn, v2 ? kNoBn : bn_of(net->stem_bn), nullptr, v2 ? 0 : 1, buf[0], st))
      return rc;
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
    }
    // ---- heads (:84-124): 1x1 conv (on relu(bn(x1)) for v2, :80-82) -> bn -> relu -> flatten -> linear [-> relu -> linear]
    const BnDev trunk_bn = v2 ? bn_of(net->final_bn) : kNoBn;
    HeadConvArgs hca{};
    float* hcbuf[4];
    for (int h = 0; h < 4; ++h) {
      const int k = h < 2 ? 2 : 1;
      float* dst = h == 0 ? exploit_logits : (h == 1 ? explore_logits : nullptr);
      hcbuf[h] = (h < 2 && !dst) ? nullptr : take((size_t)B * HW * k);
      hca.w[h] = net->head_conv[h].w;
      hca.b[h] = net->head_conv[h].b;
      hca.bn[h] = bn_of(net->head_bn[h]);
      hca.out[h] = hcbuf[h];
      hca.k[h] = k;
    }
    head_conv_kernel<<<ceil_div(B * HW, 256), 256, (size_t)9 * C * sizeof(float), st>>>(buf[cur], B * HW, C, trunk_bn, hca);
    EAZ_CHECK_LAUNCH("head_conv_kernel");
    for (int h = 0; h < 4; ++h) {
      const int k = h < 2 ? 2 : 1;
      float* hc = hcbuf[h];
      if (!hc) continue;
      if (h < 2) {
        float* dst = h == 0 ? exploit_logits : explore_logits;
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

int eaz_convnet_numeric_status(const eaz_convnet_params* net, int32_t B, const void* workspace, size_t workspace_bytes, void* stream,
                               int32_t* flags_out) {
  if (int rc = check_convnet(net)) return rc;
  if (flags_out) *flags_out = 0;
  if (!convnet_tensor(net) || B < 1) return 0;  // the fp32 path has no scaled split
  const ConvnetLayout L = convnet_layout(net, B);
  EAZ_CHECK_ARG(workspace && workspace_bytes >= L.total, "convnet numeric status: workspace too small");
  if (cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream); e != cudaSuccess) return cuda_fail(e, "convnet numeric status sync");
  uint32_t flags = 0;
  for (int g = 0; g < 2; ++g) {
    uint32_t f = 0;
    const uint8_t* p = (const uint8_t*)workspace + L.status_off + g * sizeof(NumStatus) + offsetof(NumStatus, flags);
    if (cudaError_t e = cudaMemcpy(&f, p, sizeof(f), cudaMemcpyDeviceToHost); e != cudaSuccess) return cuda_fail(e, "convnet numeric status read");
    if (g == 0 || 12 * g < 2 * net->num_blocks) flags |= f;
  }
  if (flags_out) *flags_out = (int32_t)flags;
  if (flags == 0) return 0;
  set_error("tensor-core convolution path out of range:%s%s -- use mlp_mode EXACT for this model",
            (flags & kNumWeightsNonFinite) ? " a weight tensor holds inf / nan or |w| > 2^20;" : "",
            (flags & kNumActSaturated) ? " an activation exceeded 4094 (fp16 range of the 16x-scaled split) and was clamped;" : "");
  return EAZ_ERR_UNSUPPORTED;
}

}  // extern "C"

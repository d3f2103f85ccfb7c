// mlp_tensor.cu -- tensor-core (tcgen05) evaluation of EpistemicFullyConnectedAZNet
// (network/fully_connected.py:41-101) for the expand step of the search.
//
// One CTA = one head x 128 rows (nodes).  All three hk.Linear layers of the head run as
// tcgen05.mma kind::tf32 with 3xTF32 split precision (umma.cuh) and fp32 accumulators in TMEM:
//   D1 (TMEM cols   0..255) = x  @ W1      Subleq / dense observations only: x is 0/1, exact in TF32,
//                                          so two products (x*W1_hi + x*W1_lo) suffice.  One-hot DeepSea
//                                          observations skip the GEMM: h1 = relu(W1[cell] + b1) is a row gather.
//   D2 (TMEM cols 256..511) = h1 @ W2      h1 = relu(D1 + b1), read back with tcgen05.ld by the 4 worker warps
//   D3 (TMEM cols   0..Np-1) = h2 @ W3     (thread = row), split hi/lo and written to the A stage in shared memory
// Weights are pre-split and pre-tiled once per search (tile_weights_kernel) so that a K-chunk of B is one
// contiguous block fetched with a single 1-D bulk async copy (UBLKCP) that signals an mbarrier.
// Warp roles: warps 0-3 produce A chunks / run the epilogue (TMEM lane == row), warp 4 issues copies and MMAs.
// 2-stage pipeline: A(32 KB) + B(64 KB) per stage.
//
// Accuracy: <= ~1e-6 relative to the fp32 EXACT contract (tests: 1e-5); NOT bit-identical to it, so search
// parity in this mode is proven by replaying the GPU's per-node network outputs through the oracle.
#include "mlp.cuh"
#include "umma.cuh"

namespace eaz {
using namespace umma;

int launch_tile_weights(const float* W, int K, int N, int Kpad, int Npad, uint32_t* out, cudaStream_t st);

constexpr int kTM = 128;                          // rows per CTA
constexpr int kAStage = 2 * kTM * kChunkK * 4;    // hi + lo
constexpr int kBStageMax = 2 * 256 * kChunkK * 4; // hi + lo, N = 256
constexpr int kH = 256;

struct TensorSmem {
  uint64_t full_a[2], full_b[2], empty[2], acc_done[3];
  uint32_t tmem_base;
  float bias[3][kH];
};
constexpr size_t kTensorSmemBytes = 2 * kAStage + 2 * kBStageMax + sizeof(TensorSmem) + 128;

__device__ __forceinline__ int sq_word_g(const uint8_t* st, int row, int ws, int trow) {
  if (row < ws) return st[EAZ_SQ_HDR + row];
  const int r = row - ws, part = r >> 3, i = r & 7;
  const uint16_t* h = reinterpret_cast<const uint16_t*>(st);
  if (part == 0) return sq_test_in(trow, 0, i, ws);
  if (part == 1) return h[i];
  if (part == 2) return sq_test_out(trow, 0, i, ws);
  return h[8 + i];
}
__device__ __forceinline__ int sq_bit_g(int v, int c, int w, int ws, int binary) {
  if (binary) {
    if (c == w - 1) return v == ws;
    const unsigned m = (unsigned)floormod(v, ws) & 0xffu;
    return c < 8 ? (int)((m >> c) & 1u) : 0;
  }
  return c == (v == ws ? ws : floormod(v, ws));
}

// write 32 fp32 values (one K-chunk of this thread's row) into the A stage as hi / lo tiles
__device__ __forceinline__ void store_a_chunk(uint8_t* stage, int row, const float (&v)[32], bool with_lo) {
  uint8_t* hi = stage;
  uint8_t* lo = stage + kTM * kChunkK * 4;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    uint4 h, l;
    split_tf32(v[4 * q + 0], h.x, l.x);
    split_tf32(v[4 * q + 1], h.y, l.y);
    split_tf32(v[4 * q + 2], h.z, l.z);
    split_tf32(v[4 * q + 3], h.w, l.w);
    const int off = tile_offset(row, q * 4);
    *reinterpret_cast<uint4*>(hi + off) = h;
    if (with_lo) *reinterpret_cast<uint4*>(lo + off) = l;
  }
}

struct TensorHeads {
  int n;
  int head[4];
};

__global__ void __launch_bounds__(160, 1) mlp_tensor_kernel(NetDesc net, EnvDesc env, MlpSource src, TensorWeights tw, int B, TensorHeads heads,
                                                            MlpOutputs out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + 2 * kAStage;
  TensorSmem* sh = reinterpret_cast<TensorSmem*>(sB + 2 * kBStageMax);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = heads.head[blockIdx.y];
  const int r0 = blockIdx.x * kTM;
  const int nrows = min(kTM, B - r0);
  const bool gather = env.kind == EAZ_ENV_DEEPSEA && src.compact != nullptr;
  const bool policy = head >= EAZ_HEAD_EXPLOIT;
  const int nout = policy ? net.A : 1;
  const int np3 = (nout + 15) & ~15;
  const int n1 = gather ? 0 : tw.k1pad / kChunkK;  // layer-1 chunks
  const int total = n1 + 16;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sh->full_a[s], kTM);
      mbar_init(&sh->full_b[s], 1);
      mbar_init(&sh->empty[s], 1);
    }
    for (int i = 0; i < 3; ++i) mbar_init(&sh->acc_done[i], 1);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 3 * kH; i += blockDim.x) {
    const int l = i / kH, j = i % kH;
    sh->bias[l][j] = (l < 2 || j < nout) ? __ldg(net.b[head][l] + j) : 0.0f;
  }
  if (warp == 4) {
    tmem_alloc(&sh->tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh->tmem_base;

  if (warp == 4) {
    // ================= control warp: B copies + MMA issue =================
    if (lane == 0) {
      for (int t = 0; t < total; ++t) {
        const int s = t & 1, ph = (t >> 1) & 1;
        const int layer = t < n1 ? 0 : (t < n1 + 8 ? 1 : 2);
        const int c = layer == 0 ? t : (layer == 1 ? t - n1 : t - n1 - 8);
        const int npad = layer == 2 ? np3 : kH;
        const uint32_t bbytes = (uint32_t)(2 * npad * kChunkK * 4);
        mbar_wait(&sh->empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&sh->full_b[s], bbytes);
        bulk_g2s(sB + s * kBStageMax, tw.img[head][layer] + (size_t)c * (bbytes / 4), bbytes, &sh->full_b[s]);
        mbar_wait(&sh->full_b[s], ph);
        mbar_wait(&sh->full_a[s], ph);
        tc_fence_after();
        const uint32_t idesc = idesc_tf32(kTM, npad);
        const uint32_t d = tmem + (layer == 1 ? 256u : 0u);
        const uint32_t a_hi = smem_u32(sA + s * kAStage), a_lo = a_hi + kTM * kChunkK * 4;
        const uint32_t b_hi = smem_u32(sB + s * kBStageMax), b_lo = b_hi + npad * kChunkK * 4;
#pragma unroll
        for (int j = 0; j < kKSteps; ++j) {
          const uint32_t o = j * kKStepBytes;
          mma_tf32(d, smem_desc(a_hi + o), smem_desc(b_hi + o), idesc, (c | j) != 0);
          mma_tf32(d, smem_desc(a_hi + o), smem_desc(b_lo + o), idesc, 1);
          if (layer != 0) mma_tf32(d, smem_desc(a_lo + o), smem_desc(b_hi + o), idesc, 1);  // x is exact in TF32: no lo part
        }
        mma_commit(&sh->empty[s]);
        if (t == n1 - 1) mma_commit(&sh->acc_done[0]);
        if (t == n1 + 7) mma_commit(&sh->acc_done[1]);
        if (t == total - 1) mma_commit(&sh->acc_done[2]);
      }
    }
    __syncwarp();
  } else {
    // ================= worker warps: A producer (thread == row) + epilogue =================
    const int row = threadIdx.x;
    const bool live = row < nrows;
    const int b = r0 + (live ? row : 0);
    const uint8_t* st = nullptr;
    int cell = 0, trow = 0;
    if (src.compact) {
      const size_t slot = src.node_index ? ((size_t)src.node_index[b] * B + b) : (size_t)b;
      st = src.compact + slot * env.compact_bytes;
      if (env.kind == EAZ_ENV_DEEPSEA) cell = deepsea_obs_index(*reinterpret_cast<const uint32_t*>(st), env.size);
      else trow = sq_task_row(st[34]);
    }
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int t = 0; t < total; ++t) {
      const int s = t & 1, ph = (t >> 1) & 1;
      const int layer = t < n1 ? 0 : (t < n1 + 8 ? 1 : 2);
      const int c = layer == 0 ? t : (layer == 1 ? t - n1 : t - n1 - 8);
      float v[32];
      if (layer == 0) {  // observation bits of K-chunk c
        const int k0 = c * kChunkK, w = env.obs_cols, ws = env.ws;
        if (st) {
          int orow = k0 / w, oc = k0 - orow * w;
          int word = (live && k0 < net.D) ? sq_word_g(st, orow, ws, trow) : 0;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const bool in = live && (k0 + i) < net.D;
            v[i] = (in && sq_bit_g(word, oc, w, ws, env.binary)) ? 1.0f : 0.0f;
            if (++oc == w) {
              oc = 0;
              ++orow;
              if (in && k0 + i + 1 < net.D) word = sq_word_g(st, orow, ws, trow);
            }
          }
        } else {
          const uint8_t* o = src.dense + (size_t)b * net.D + k0;
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = (live && k0 + i < net.D && o[i]) ? 1.0f : 0.0f;
        }
      } else if (layer == 1 && gather) {  // h1 = relu(W1[cell] + b1): one-hot observation
        const float4* wrow = reinterpret_cast<const float4*>(net.w[head][0] + (size_t)cell * kH + c * kChunkK);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 x = live ? __ldg(wrow + q) : make_float4(0.f, 0.f, 0.f, 0.f);
          v[4 * q + 0] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = live ? fmaxf(__fadd_rn(v[i], sh->bias[0][c * kChunkK + i]), 0.0f) : 0.0f;
      } else {  // relu(previous accumulator + bias) read back from TMEM
        if (c == 0) {
          mbar_wait(&sh->acc_done[layer - 1], 0);
          tc_fence_after();
        }
        uint32_t r[32];
        tmem_ld32(tmem + lane_base + (layer == 1 ? 0u : 256u) + c * kChunkK, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = live ? fmaxf(__fadd_rn(__uint_as_float(r[i]), sh->bias[layer - 1][c * kChunkK + i]), 0.0f) : 0.0f;
      }
      mbar_wait(&sh->empty[s], ph ^ 1);
      store_a_chunk(sA + s * kAStage, row, v, layer != 0);
      fence_proxy_async();
      mbar_arrive(&sh->full_a[s]);
    }

    // ---- novelty probe (UBE head): fully_connected.py:83-90
    int seen = 0;
    if (head == EAZ_HEAD_UBE && live) {
      if (gather && src.ds_seen) {
        seen = src.ds_seen[cell];
      } else {
        const int L = net.hash_dim >> 2;
        uint32_t a[4];
#pragma unroll
        for (int l = 0; l < 4; ++l) a[l] = xx_init(l);
        const int kbeg0 = net.D - net.hash_dim;
        if (gather) {
          for (int i = 0; i < L; ++i)
#pragma unroll
            for (int l = 0; l < 4; ++l) a[l] = xx_round(a[l], (kbeg0 + l * L + i == cell) ? EAZ_XX_ONE : 0u);
        } else if (st) {
          const int w = env.obs_cols, ws = env.ws;
#pragma unroll
          for (int l = 0; l < 4; ++l) {
            const int kb = kbeg0 + l * L;
            int orow = kb / w, oc = kb - orow * w;
            int word = sq_word_g(st, orow, ws, trow);
            for (int i = 0; i < L; ++i) {
              a[l] = xx_round(a[l], sq_bit_g(word, oc, w, ws, env.binary) ? EAZ_XX_ONE : 0u);
              if (++oc == w) {
                oc = 0;
                ++orow;
                if (i + 1 < L) word = sq_word_g(st, orow, ws, trow);
              }
            }
          }
        } else {
          const uint8_t* o = src.dense + (size_t)b * net.D + kbeg0;
          for (int i = 0; i < L; ++i)
#pragma unroll
            for (int l = 0; l < 4; ++l) a[l] = xx_round(a[l], o[l * L + i] ? EAZ_XX_ONE : 0u);
        }
        const uint32_t idx = xx_finish(a[0], a[1], a[2], a[3], L, net.hash_bits);
        seen = (net.bset[idx >> 3] >> (idx & 7u)) & 1u;
      }
    }

    // ---- layer-3 epilogue
    mbar_wait(&sh->acc_done[2], 0);
    tc_fence_after();
    float* logits = policy ? out.logits[head - EAZ_HEAD_EXPLOIT] : nullptr;
    for (int n0 = 0; n0 < np3; n0 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem + lane_base + n0, r);
      tmem_ld_wait();
      if (!live) continue;
      if (policy) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (n0 + i < nout) logits[(size_t)b * nout + n0 + i] = __fadd_rn(__uint_as_float(r[i]), sh->bias[2][n0 + i]);
      } else {
        const float y = __fadd_rn(__uint_as_float(r[0]), sh->bias[2][0]);
        if (head == EAZ_HEAD_VALUE) {
          out.value[b] = eaz_tanh(y);
        } else {
          float u = __fmul_rn(0.5f, __fadd_rn(eaz_tanh(y), 1.0f));
          const float nov = __fmul_rn(seen ? 0.0f : 1.0f, net.novelty_scale);
          u = __fmul_rn(u, net.max_u);
          u = eaz_max(nov, u);
          u = eaz_min(eaz_max(u, 0.0f), net.max_u);
          if (out.ube) out.ube[b] = u;
          if (out.novelty) out.novelty[b] = nov;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------- host side
size_t tensor_weights_bytes(const NetDesc& net, const EnvDesc& env) {
  const int k1pad = (net.D + kChunkK - 1) / kChunkK * kChunkK;
  const int np3 = (net.A + 15) & ~15;
  const size_t l1 = env.kind == EAZ_ENV_DEEPSEA ? 0 : (size_t)k1pad * 2 * kH * 4;
  const size_t l2 = (size_t)kH * 2 * kH * 4;
  const size_t per_head = l1 + l2 + (size_t)kH * 2 * np3 * 4;
  return 4 * ((per_head + 255) & ~(size_t)255);
}

int prepare_tensor_weights(const NetDesc& net, const EnvDesc& env, int heads_mask, void* buf, TensorWeights* tw, cudaStream_t st) {
  if (net.H != kH) {
    set_error("tensor network path needs hidden size %d (got %d); use mlp_mode EXACT", kH, net.H);
    return EAZ_ERR_UNSUPPORTED;
  }
  const int k1pad = (net.D + kChunkK - 1) / kChunkK * kChunkK;
  const bool has_l1 = env.kind != EAZ_ENV_DEEPSEA;
  const size_t l1 = has_l1 ? (size_t)k1pad * 2 * kH * 4 : 0, l2 = (size_t)kH * 2 * kH * 4;
  const size_t per_head = tensor_weights_bytes(net, env) / 4;
  tw->k1pad = k1pad;
  for (int h = 0; h < 4; ++h) {
    uint8_t* p = (uint8_t*)buf + (size_t)h * per_head;
    tw->img[h][0] = (const uint32_t*)p;
    tw->img[h][1] = (const uint32_t*)(p + l1);
    tw->img[h][2] = (const uint32_t*)(p + l1 + l2);
    if (!(heads_mask & (1 << h))) continue;
    const int nout = h >= EAZ_HEAD_EXPLOIT ? net.A : 1;
    if (has_l1)
      if (int rc = launch_tile_weights(net.w[h][0], net.D, kH, k1pad, kH, (uint32_t*)p, st)) return rc;
    if (int rc = launch_tile_weights(net.w[h][1], kH, kH, kH, kH, (uint32_t*)(p + l1), st)) return rc;
    if (int rc = launch_tile_weights(net.w[h][2], kH, nout, kH, (nout + 15) & ~15, (uint32_t*)(p + l1 + l2), st)) return rc;
  }
  return 0;
}

int launch_mlp_tensor(const NetDesc& net, const EnvDesc& env, const MlpSource& src, const TensorWeights& tw, int B, int heads_mask,
                      const MlpOutputs& out, cudaStream_t stream) {
  TensorHeads hl{0, {0, 0, 0, 0}};
  for (int h = 0; h < 4; ++h)
    if (heads_mask & (1 << h)) hl.head[hl.n++] = h;
  if (hl.n == 0 || B == 0) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(mlp_tensor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTensorSmemBytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mlp_tensor_kernel)");
    attr_set = true;
  }
  dim3 grid(ceil_div(B, kTM), hl.n);
  mlp_tensor_kernel<<<grid, 160, kTensorSmemBytes, stream>>>(net, env, src, tw, B, hl, out);
  EAZ_CHECK_LAUNCH("mlp_tensor_kernel");
  return 0;
}

}  // namespace eaz

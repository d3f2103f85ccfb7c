// mlp.cu -- EpistemicFullyConnectedAZNet inference (network/fully_connected.py:41-101):
// 4 independent heads x (D -> H relu -> H relu -> out), value tanh, UBE 0.5(tanh+1)
// combined with the hash-count novelty probe (:83-96).
//
// EXACT mode: every hk.Linear is the fp32 contract of the CPU oracle
//   acc = 0; for k ascending: acc = fma(x[k], w[k][j], acc); y = acc + b[j]
// so results are bit-identical to oracle/eaz_oracle.c:orc_linear.  One CTA
// evaluates one head for a tile of R rows; activations live transposed in shared
// memory ([k][row], row stride R+4 floats so that float4 row-vectors are aligned
// and the transposing stores spread over banks); each thread owns one output
// column and R accumulators, so a weight element is loaded once per R FMAs.
// One-hot DeepSea observations skip layer 1's GEMM: x@W1 is the row W1[cell].
#include "mlp.cuh"

namespace eaz {

constexpr int kR = 32;        // rows per CTA
constexpr int kRS = kR + 4;   // padded row stride of transposed activations
constexpr int kKC = 64;       // layer-1 feature chunk
constexpr int kThreads = 256;
constexpr int kStatePad = 304;  // >= 40 + 256, multiple of 8

struct MlpSmem {
  float xT[EAZ_FC_HIDDEN_MAX * kRS];  // activations of the current layer, [k][row]
  float xc[kKC * kRS];                // layer-1 input chunk, [k][row] as 0.0f / 1.0f
  uint8_t st[kR * kStatePad];         // staged compact Subleq states
  int cell[kR];                       // DeepSea observation cell per row
  uint8_t seen[kR];
};

int make_net_desc(const eaz_fc_params* net, const EnvDesc* env, NetDesc* d) {
  EAZ_CHECK_ARG(net != nullptr, "net is NULL");
  EAZ_CHECK_ARG(net->hidden >= 1 && net->hidden <= EAZ_FC_HIDDEN_MAX, "hidden size %d outside [1,%d]", net->hidden, EAZ_FC_HIDDEN_MAX);
  EAZ_CHECK_ARG(net->num_actions >= 1 && net->num_actions <= 256, "num_actions %d outside [1,256]", net->num_actions);
  EAZ_CHECK_ARG(net->in_dim >= 1, "in_dim must be positive");
  EAZ_CHECK_ARG(net->hash_bits > 0 && net->hash_bits <= 32, "bits_per_hash %d violates 0 < bits <= 32 (hashes.py:154)", net->hash_bits);
  d->D = net->in_dim;
  d->H = net->hidden;
  d->A = net->num_actions;
  for (int h = 0; h < 4; ++h)
    for (int l = 0; l < 3; ++l) {
      EAZ_CHECK_ARG(net->w[h][l] && net->b[h][l], "net: w[%d][%d] / b[%d][%d] is NULL", h, l, h, l);
      d->w[h][l] = net->w[h][l];
      d->b[h][l] = net->b[h][l];
    }
  d->bset = net->binary_set;
  EAZ_CHECK_ARG(d->bset != nullptr, "net: binary_set is NULL");
  d->hash_bits = net->hash_bits;
  d->hash_io = net->hash_io;
  d->max_u = net->max_u;
  d->novelty_scale = net->novelty_scale;
  d->hash_dim = d->D;
  if (!env && d->hash_io && net->word_size > 0) {  // dense observations: rows word_size.. of [ws+32, w]
    EAZ_CHECK_ARG(d->D % (net->word_size + 32) == 0, "in_dim %d is not (word_size+32) x cols", d->D);
    d->hash_dim = 32 * (d->D / (net->word_size + 32));
  }
  if (env) {
    EAZ_CHECK_ARG(env->obs_dim == d->D, "net in_dim %d != env observation size %d", d->D, env->obs_dim);
    EAZ_CHECK_ARG(env->num_actions == d->A, "net num_actions %d != env num_actions %d", d->A, env->num_actions);
    if (env->kind == EAZ_ENV_SUBLEQ && d->hash_io) d->hash_dim = 32 * env->obs_cols;  // fully_connected.py:85-89
  }
  EAZ_CHECK_ARG(d->hash_dim % 4 == 0, "hash input length %d is not a multiple of 4 (hashes.py:210)", d->hash_dim);
  return 0;
}

// ---- observation bits straight from a compact Subleq state (Subleq._observe, subleq.py:679-707)
__device__ __forceinline__ int sq_word_compact(const uint8_t* st, int row, int ws, int trow) {
  if (row < ws) return st[EAZ_SQ_HDR + row];
  const int r = row - ws, part = r >> 3, i = r & 7;
  const uint16_t* h = reinterpret_cast<const uint16_t*>(st);
  if (part == 0) return sq_test_in(trow, 0, i, ws);
  if (part == 1) return h[i];
  if (part == 2) return sq_test_out(trow, 0, i, ws);
  return h[8 + i];
}
__device__ __forceinline__ int sq_obs_bit(int v, int c, int w, int ws, int binary) {
  if (binary) {  // subleq.py:88-97
    if (c == w - 1) return v == ws;
    const unsigned m = (unsigned)floormod(v, ws) & 0xffu;
    return c < 8 ? (int)((m >> c) & 1u) : 0;
  }
  return c == (v == ws ? ws : floormod(v, ws));  // subleq.py:51-55
}

__device__ __forceinline__ const uint8_t* state_ptr(const MlpSource& src, int b, int B, int S) {
  const size_t slot = src.node_index ? ((size_t)src.node_index[b] * B + b) : (size_t)b;
  return src.compact + slot * S;
}

// acc[r] = fma(x[k][r], W[k][col], acc[r]) for k ascending over [0,K)
__device__ __forceinline__ void gemm_rows(const float* __restrict__ xT, const float* __restrict__ W, int ldw, int col, int K,
                                          float (&acc)[kR]) {
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float w = __ldg(W + (size_t)k * ldw + col);
    const float4* xr = reinterpret_cast<const float4*>(xT + k * kRS);
#pragma unroll
    for (int q = 0; q < kR / 4; ++q) {
      const float4 x = xr[q];
      acc[4 * q + 0] = __fmaf_rn(x.x, w, acc[4 * q + 0]);
      acc[4 * q + 1] = __fmaf_rn(x.y, w, acc[4 * q + 1]);
      acc[4 * q + 2] = __fmaf_rn(x.z, w, acc[4 * q + 2]);
      acc[4 * q + 3] = __fmaf_rn(x.w, w, acc[4 * q + 3]);
    }
  }
}

__device__ __forceinline__ void store_relu_T(float* xT, int col, const float (&acc)[kR], float bias) {
  float4* dst = reinterpret_cast<float4*>(xT + col * kRS);
#pragma unroll
  for (int q = 0; q < kR / 4; ++q) {
    float4 v;
    v.x = fmaxf(__fadd_rn(acc[4 * q + 0], bias), 0.0f);
    v.y = fmaxf(__fadd_rn(acc[4 * q + 1], bias), 0.0f);
    v.z = fmaxf(__fadd_rn(acc[4 * q + 2], bias), 0.0f);
    v.w = fmaxf(__fadd_rn(acc[4 * q + 3], bias), 0.0f);
    dst[q] = v;
  }
}

struct HeadList {
  int n;
  int head[4];
};

__global__ void __launch_bounds__(kThreads) mlp_exact_kernel(NetDesc net, EnvDesc env, MlpSource src, int B, HeadList heads,
                                                             MlpOutputs out) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  MlpSmem& sm = *reinterpret_cast<MlpSmem*>(smem_raw);
  const int head = heads.head[blockIdx.y];
  const int r0 = blockIdx.x * kR;
  const int nrows = min(kR, B - r0);
  const int tid = threadIdx.x;
  const int H = net.H, D = net.D;
  const bool compact = src.compact != nullptr;
  const bool gather = compact && env.kind == EAZ_ENV_DEEPSEA;
  const int S = env.compact_bytes;

  // ---- stage the rows' env states
  if (compact) {
    if (env.kind == EAZ_ENV_DEEPSEA) {
      if (tid < kR) sm.cell[tid] = tid < nrows ? deepsea_obs_index(*reinterpret_cast<const uint32_t*>(state_ptr(src, r0 + tid, B, S)), env.size) : 0;
    } else {
      const int words = S / 8;
      for (int e = tid; e < nrows * words; e += kThreads) {
        const int r = e / words, i = e % words;
        reinterpret_cast<uint2*>(sm.st + r * kStatePad)[i] = reinterpret_cast<const uint2*>(state_ptr(src, r0 + r, B, S))[i];
      }
    }
  }
  __syncthreads();

  float acc[kR];
  // ---- layer 1
  const float* W1 = net.w[head][0];
  if (gather) {
    if (tid < H) {
#pragma unroll
      for (int r = 0; r < kR; ++r) acc[r] = r < nrows ? __ldg(W1 + (size_t)sm.cell[r] * H + tid) : 0.0f;
    }
  } else {
#pragma unroll
    for (int r = 0; r < kR; ++r) acc[r] = 0.0f;
    const int w = env.obs_cols, ws = env.ws;
    for (int k0 = 0; k0 < D; k0 += kKC) {
      const int kc = min(kKC, D - k0);
      for (int e = tid; e < kKC * kR; e += kThreads) {
        const int kk = e % kKC, r = e / kKC;
        float x = 0.0f;
        if (kk < kc && r < nrows) {
          const int k = k0 + kk;
          if (compact) {
            const uint8_t* st = sm.st + r * kStatePad;
            const int row = k / w, c = k - row * w;
            x = sq_obs_bit(sq_word_compact(st, row, ws, sq_task_row(st[34])), c, w, ws, env.binary) ? 1.0f : 0.0f;
          } else {
            x = src.dense[(size_t)(r0 + r) * D + k] ? 1.0f : 0.0f;
          }
        }
        sm.xc[kk * kRS + r] = x;
      }
      __syncthreads();
      if (tid < H) gemm_rows(sm.xc, W1 + (size_t)k0 * H, H, tid, kc, acc);
      __syncthreads();
    }
  }
  if (tid < H) store_relu_T(sm.xT, tid, acc, __ldg(net.b[head][0] + tid));
  __syncthreads();

  // ---- layer 2
  if (tid < H) {
#pragma unroll
    for (int r = 0; r < kR; ++r) acc[r] = 0.0f;
    gemm_rows(sm.xT, net.w[head][1], H, tid, H, acc);
  }
  __syncthreads();  // everyone finished reading h1
  if (tid < H) store_relu_T(sm.xT, tid, acc, __ldg(net.b[head][1] + tid));

  // ---- novelty probe for the UBE head: 4 threads (xxhash lanes) per row
  if (head == EAZ_HEAD_UBE && tid < 4 * kR) {
    const int r = tid >> 2, lane = tid & 3;
    int seen = 0;
    if (gather && src.ds_seen) {
      if (lane == 0 && r < nrows) seen = src.ds_seen[sm.cell[r]];
    } else {
      const int L = net.hash_dim >> 2;
      uint32_t a = xx_init(lane);
      if (r < nrows) {
        const int kbeg = D - net.hash_dim + lane * L;  // trailing hash_dim elements (fully_connected.py:85-89)
        if (gather) {
          const int cell = sm.cell[r];
          for (int i = 0; i < L; ++i) a = xx_round(a, (kbeg + i == cell) ? EAZ_XX_ONE : 0u);
        } else if (compact) {
          const uint8_t* st = sm.st + r * kStatePad;
          const int w = env.obs_cols, ws = env.ws, trow = sq_task_row(st[34]);
          int row = kbeg / w, c = kbeg - row * w;
          int v = sq_word_compact(st, row, ws, trow);
          for (int i = 0; i < L; ++i) {
            a = xx_round(a, sq_obs_bit(v, c, w, ws, env.binary) ? EAZ_XX_ONE : 0u);
            if (++c == w) {
              c = 0;
              ++row;
              if (i + 1 < L) v = sq_word_compact(st, row, ws, trow);
            }
          }
        } else {
          const uint8_t* o = src.dense + (size_t)(r0 + r) * D + kbeg;
          for (int i = 0; i < L; ++i) a = xx_round(a, o[i] ? EAZ_XX_ONE : 0u);
        }
      }
      const unsigned base = (tid & 31u) & ~3u;
      const uint32_t a0 = __shfl_sync(0xffffffffu, a, base + 0), a1 = __shfl_sync(0xffffffffu, a, base + 1),
                     a2 = __shfl_sync(0xffffffffu, a, base + 2), a3 = __shfl_sync(0xffffffffu, a, base + 3);
      if (lane == 0 && r < nrows) {
        const uint32_t idx = xx_finish(a0, a1, a2, a3, L, net.hash_bits);
        seen = (net.bset[idx >> 3] >> (idx & 7u)) & 1u;
      }
    }
    if (lane == 0) sm.seen[r] = (uint8_t)seen;
  }
  __syncthreads();

  // ---- layer 3 + head epilogue
  const bool policy = head >= EAZ_HEAD_EXPLOIT;
  const int nout = policy ? net.A : 1;
  const float* W3 = net.w[head][2];
  const float* b3 = net.b[head][2];
  float* logits = policy ? out.logits[head - EAZ_HEAD_EXPLOIT] : nullptr;
  if (nout > 32) {
    if (tid < nout) {
#pragma unroll
      for (int r = 0; r < kR; ++r) acc[r] = 0.0f;
      gemm_rows(sm.xT, W3, nout, tid, H, acc);
      const float bias = __ldg(b3 + tid);
#pragma unroll
      for (int r = 0; r < kR; ++r)
        if (r < nrows) logits[(size_t)(r0 + r) * nout + tid] = __fadd_rn(acc[r], bias);
    }
    return;
  }
  const int total = kR * nout;
  for (int base = 0; base < total; base += kThreads * 4) {
    float a4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    int row[4], col[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = min(base + tid + kThreads * i, total - 1);
      col[i] = p % nout;
      row[i] = p / nout;
    }
    for (int k = 0; k < H; ++k) {
#pragma unroll
      for (int i = 0; i < 4; ++i) a4[i] = __fmaf_rn(sm.xT[k * kRS + row[i]], __ldg(W3 + (size_t)k * nout + col[i]), a4[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = base + tid + kThreads * i;
      if (p >= total || row[i] >= nrows) continue;
      const int b = r0 + row[i];
      const float y = __fadd_rn(a4[i], __ldg(b3 + col[i]));
      if (policy) {
        logits[(size_t)b * nout + col[i]] = y;
      } else if (head == EAZ_HEAD_VALUE) {
        out.value[b] = eaz_tanh(y);  // fully_connected.py:55
      } else {
        float u = __fmul_rn(0.5f, __fadd_rn(eaz_tanh(y), 1.0f));             // :64
        const float nov = __fmul_rn(sm.seen[row[i]] ? 0.0f : 1.0f, net.novelty_scale);  // :90
        u = __fmul_rn(u, net.max_u);                                         // :93
        u = eaz_max(nov, u);                                                 // :95
        u = eaz_min(eaz_max(u, 0.0f), net.max_u);                            // :96
        if (out.ube) out.ube[b] = u;
        if (out.novelty) out.novelty[b] = nov;
      }
    }
  }
}

// One thread per grid cell: hash of the one-hot observation with that cell set.
__global__ void deepsea_seen_table_kernel(NetDesc net, int D, uint8_t* __restrict__ seen) {
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= D) return;
  const int L = D >> 2;
  uint32_t a[4];
#pragma unroll
  for (int l = 0; l < 4; ++l) a[l] = xx_init(l);
  const int hot_lane = cell / L, hot_i = cell - hot_lane * L;
  for (int i = 0; i < L; ++i) {
#pragma unroll
    for (int l = 0; l < 4; ++l) a[l] = xx_round(a[l], (l == hot_lane && i == hot_i) ? EAZ_XX_ONE : 0u);
  }
  const uint32_t idx = xx_finish(a[0], a[1], a[2], a[3], L, net.hash_bits);
  seen[cell] = (uint8_t)((net.bset[idx >> 3] >> (idx & 7u)) & 1u);
}

int launch_deepsea_seen_table(const NetDesc& net, const EnvDesc& env, uint8_t* seen, cudaStream_t stream) {
  deepsea_seen_table_kernel<<<ceil_div(env.obs_dim, 64), 64, 0, stream>>>(net, env.obs_dim, seen);
  EAZ_CHECK_LAUNCH("deepsea_seen_table");
  return 0;
}

int mlp_num_launches(int mode) { return 1; }

int launch_mlp(const NetDesc& net, const EnvDesc& env, const MlpSource& src, int B, int heads_mask, const MlpOutputs& out, int mode,
               cudaStream_t stream, const TensorWeights* tw) {
  if (mode == EAZ_MLP_TENSOR) {
    if (!tw) {
      set_error("mlp_mode TENSOR needs prepared weight images");
      return EAZ_ERR_INVALID_ARG;
    }
    return launch_mlp_tensor(net, env, src, *tw, B, heads_mask, out, stream);
  }
  if (mode != EAZ_MLP_EXACT) {
    set_error("unknown mlp_mode %d", mode);
    return EAZ_ERR_INVALID_ARG;
  }
  HeadList hl{0, {0, 0, 0, 0}};
  for (int h = 0; h < 4; ++h)
    if (heads_mask & (1 << h)) hl.head[hl.n++] = h;
  if (hl.n == 0 || B == 0) return 0;
  static bool attr_set = false;  // idempotent; a race only repeats the call
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(mlp_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MlpSmem));
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mlp_exact_kernel)");
    attr_set = true;
  }
  dim3 grid(ceil_div(B, kR), hl.n);
  mlp_exact_kernel<<<grid, kThreads, sizeof(MlpSmem), stream>>>(net, env, src, B, hl, out);
  EAZ_CHECK_LAUNCH("mlp_exact_kernel");
  return 0;
}

}  // namespace eaz

using namespace eaz;

extern "C" {

static int mlp_entry(const eaz_fc_params* net, const eaz_env* env, const MlpSource& src, int32_t B, float* exploit_logits,
                     float* explore_logits, float* value, float* ube, float* novelty, const EnvDesc& ed, void* stream) {
  NetDesc nd;
  if (int rc = make_net_desc(net, env ? &ed : nullptr, &nd)) return rc;
  MlpOutputs out{{exploit_logits, explore_logits}, value, ube, novelty};
  int mask = 0;
  if (value) mask |= 1 << EAZ_HEAD_VALUE;
  if (ube || novelty) mask |= 1 << EAZ_HEAD_UBE;
  if (exploit_logits) mask |= 1 << EAZ_HEAD_EXPLOIT;
  if (explore_logits) mask |= 1 << EAZ_HEAD_EXPLORE;
  return launch_mlp(nd, ed, src, B, mask, out, EAZ_MLP_EXACT, (cudaStream_t)stream);
}

int eaz_mlp_forward(const eaz_fc_params* net, const uint8_t* observation, int32_t B, float* exploit_logits, float* explore_logits,
                    float* value, float* ube, float* novelty, void* stream) {
  EAZ_CHECK_ARG(observation != nullptr && B >= 0, "eaz_mlp_forward: bad arguments");
  EnvDesc ed{};
  ed.kind = -1;
  MlpSource src{observation, nullptr, nullptr, nullptr, nullptr};
  return mlp_entry(net, nullptr, src, B, exploit_logits, explore_logits, value, ube, novelty, ed, stream);
}

}  // extern "C"

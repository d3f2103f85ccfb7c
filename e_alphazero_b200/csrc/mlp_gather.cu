// mlp_gather.cu -- tensor-core (tcgen05) evaluation of EpistemicFullyConnectedAZNet
// (network/fully_connected.py:41-101) for ONE-HOT observations (DeepSea, deep_sea.py:83-85) inside the search.
//
// For a one-hot row x (cell c), layer 1 is a row gather: h1 = relu(W1[c] + b1).  That row -- already activated,
// scaled and split into the fp16 hi / lo pair the 3xFP16 scheme of umma.cuh multiplies -- is precomputed once per
// search for every cell (h1_table_kernel), so inside the per-simulation kernel layer 1 is a pure 1 KB copy per row
// and the kernel is one split-precision GEMM plus a CUDA-core tail.  One CTA = one head x 128 rows:
//
//   D (TMEM, 256 cols) = A @ W2      8 K-chunks of 32, tcgen05.mma kind::f16 M=128 N=256, 3 products per K-step; A chunks
//                                     (gathered h1 rows) and W2 chunk images (1-D bulk async copies) go through 4-stage rings
//   y = relu(D + b2) @ W3             fp32 FMAs straight out of TMEM (all DeepSea heads have <= 4 outputs)
// (N is not split: with operands in shared memory an N=128 MMA costs as much as an N=256 one -- measured, the A tile
// re-read dominates -- and the kernel is L2->SM bandwidth-adjacent: 128 KB of rows + 256 KB of weights per CTA.)
//
// Warp roles: warps 0-7 = gather producers and layer-3 / epilogue workers, warp 8 issues the MMAs, warp 9 the
// weight copies.  Producer mapping: one warp instruction moves 8 rows x 4 sixteen-byte pieces = one 512-byte row
// group of the tile (conflict-free in shared memory, two full 32-byte sectors per row in global memory); warps 0-3
// fetch the hi halves, 4-7 the lo halves.  Barriers are polled by ONE lane per warp: hundreds of spinning threads
// measurably starve the shared-memory pipe the copies go through (profiles/r1_summary.md).
#include "mlp.cuh"
#include "umma.cuh"

namespace eaz {
using namespace umma;

namespace gk {
constexpr int kTM = 128;                  // rows per CTA
constexpr int kH = 256;
constexpr int kCK = 32;                   // K (halves) per chunk
constexpr int kChunks = kH / kCK;         // 8
constexpr int kStagesMax = 4;  // ring depth: template parameter of the kernel (4: one CTA per SM; 2: two CTAs per SM)
constexpr int kAHalf = kTM * kCK * 2;     // 8 KB: hi or lo tile of one A chunk
constexpr int kAStage = 2 * kAHalf;
constexpr int kBHalf = kH * kCK * 2;      // 16 KB: hi or lo tile of one W2 chunk
constexpr int kBStage = 2 * kBHalf;
constexpr int kSBO = (kCK * 2 / 16) * kCoreBytes;  // 512 B
constexpr int kOutMax = 4;
// (kActScale = 16, mlp.cuh; the W2 image carries a per-matrix power-of-two scale, tile_weights.cu)

struct Smem {
  uint64_t full_a[kStagesMax], full_b[kStagesMax], empty[kStagesMax], acc_done;
  uint32_t tmem_base;
  alignas(16) float b2[kH];
  alignas(16) float w3t[kOutMax][kH];  // layer-3 weights, output-major: four consecutive k per 16-byte load
  alignas(16) float part[kTM][kOutMax];  // partial sums of column group 1 (warps 4-7)
};
constexpr size_t smem_bytes(int stages) { return (size_t)stages * (kAStage + kBStage) + sizeof(Smem) + 1024; }
}  // namespace gk
using namespace gk;

// h1 table: out[cell][hi 256 halves | lo 256 halves] = split(clamp(relu(W1[cell] + b1) * kActScale))
__global__ void h1_table_kernel(const float* __restrict__ W1, const float* __restrict__ b1, int D, __half* __restrict__ out, uint32_t* num_flags) {
  const int total = D * kH;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int cell = i / kH, k = i % kH;
    // same operations as the in-kernel layer-1 epilogue of mlp_tensor.cu: fma(w, S, b*S) == (w + b) * S for S = 2^k
    float v = fmaxf(__fmaf_rn(W1[i], kActScale, b1[k] * kActScale), 0.0f);
    if (!(v <= 65504.0f)) {  // range guard: clamped AND reported (eaz_search_numeric_status)
      atomicOr(num_flags, kNumActSaturated);
      v = 65504.0f;
    }
    __half hi, lo;
    split_f16(v, hi, lo);
    out[(size_t)cell * 2 * kH + k] = hi;
    out[(size_t)cell * 2 * kH + kH + k] = lo;
  }
}

struct GatherHeads {
  int n;
  int head[4];
};
// Optional timeline (eaz_debug_set_gather_trace): CTA (0,0,0) records clock64() at its milestones.
static unsigned long long* g_gather_trace = nullptr;

template <int kStages>
__global__ void __launch_bounds__(320, 2) mlp_gather_kernel(NetDesc net, EnvDesc env, MlpSource src, TensorWeights tw, int B, GatherHeads heads,
                                                            MlpOutputs out, unsigned long long* tl, unsigned long long* trace) {
  const bool tr = trace && blockIdx.x == 0 && blockIdx.y == 0;
  if (tr && threadIdx.x == 0) trace[0] = clock64();
  const bool tl_on = tl && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0;
  const unsigned long long tl_entry = tl_on ? globaltimer_ns() : 0ull;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                      // [stage][hi tile | lo tile]
  uint8_t* sB = smem + kStages * kAStage;  // [stage][hi tile | lo tile]
  Smem* sh = reinterpret_cast<Smem*>(sB + kStages * kBStage);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = heads.head[blockIdx.y];
  const int r0 = blockIdx.x * kTM;
  const int nrows = min(kTM, B - r0);
  const bool policy = head >= EAZ_HEAD_EXPLOIT;
  const int nout = policy ? net.A : 1;

  // ---- prologue: touches only weights, overlaps the previous kernel under PDL
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sh->full_a[s], 8);  // one elected arrive per producer warp
      mbar_init(&sh->full_b[s], 1);
      mbar_init(&sh->empty[s], 1);
    }
    mbar_init(&sh->acc_done, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < kH) {
    const int j = threadIdx.x;
    const float b2 = __ldg(net.b[head][1] + j);
    float w[kOutMax] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int o = 0; o < kOutMax; ++o)
      if (o < nout) w[o] = __ldg(net.w[head][2] + (size_t)j * nout + o);
    sh->b2[j] = b2;
    sh->w3t[0][j] = w[0]; sh->w3t[1][j] = w[1]; sh->w3t[2][j] = w[2]; sh->w3t[3][j] = w[3];
  }
  if (warp == 8) {
    tmem_alloc(&sh->tmem_base, kH);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh->tmem_base;
  const float kUnscale = 1.0f / (kActScale * __ldg(tw.wscale + 3 * head + 1));  // exact: both scales are powers of two
  const uint8_t* w2img = reinterpret_cast<const uint8_t*>(tw.img[head][1]);  // per chunk: [hi tile 256 x 32 | lo tile]
  auto issue_b = [&](int c) {
    const int s = c % kStages;
    mbar_arrive_expect_tx(&sh->full_b[s], (uint32_t)kBStage);
    bulk_g2s(sB + s * kBStage, w2img + (size_t)c * kBStage, kBStage, &sh->full_b[s]);
  };
  if (warp == 9 && lane == 0)
    for (int c = 0; c < kStages; ++c) issue_b(c);  // weights never depend on the previous kernel

  // Every thread waits for the previous kernel (tree step) and only then lets the next tree kernel launch (tree_step.cuh).
  if (tr && threadIdx.x == 0) trace[1] = clock64();
  if (src.tile_done && src.tree_epoch > 0) {  // tile flags: this tile's trees are done (one polling lane per warp)
    if (threadIdx.x == 0) wait_counter(src.tile_done + blockIdx.x, src.tree_epoch * nrows);
    __syncthreads();
  } else {
    pdl_wait();
  }
  pdl_trigger();
  if (tr && threadIdx.x == 0) trace[2] = clock64();
  const unsigned long long tl_wait = tl_on ? globaltimer_ns() : 0ull;

  float y3[kOutMax] = {0.0f, 0.0f, 0.0f, 0.0f};
  const int row = threadIdx.x & (kTM - 1);
  const bool live = row < nrows;
  const int b = r0 + (live ? row : 0);

  if (warp == 9) {
    // ================= weight-copy warp =================
    if (lane == 0) {
      for (int c = kStages; c < kChunks; ++c) {
        mbar_wait(&sh->empty[c % kStages], ((c / kStages) & 1) ^ 1);
        issue_b(c);
      }
    }
    __syncwarp();
  } else if (warp == 8) {
    // ================= MMA-issue warp: warp-uniform loop, one elected lane issues (umma.cuh: elect_one) =================
    {
      const uint32_t desc_hi = (uint32_t)(kSBO >> 4) | (1u << 14);    // SBO [32,46) + version=1 [46,48)
      const uint32_t lbo_bits = (uint32_t)(kCoreBytes >> 4) << 16;    // LBO [16,30)
      auto mk = [&](uint32_t lo) { return ((uint64_t)desc_hi << 32) | (uint64_t)lo; };
      const uint32_t idesc = idesc_f16(kTM, kH);
      const uint32_t a_base = ((smem_u32(sA) & 0x3FFFFu) >> 4) | lbo_bits, b_base = ((smem_u32(sB) & 0x3FFFFu) >> 4) | lbo_bits;
#pragma unroll 1
      for (int c = 0; c < kChunks; ++c) {
        const int s = c % kStages, ph = (c / kStages) & 1;
        mbar_wait(&sh->full_b[s], ph);
        if (tr && lane == 0) trace[16 + c] = clock64();
        mbar_wait(&sh->full_a[s], ph);
        tc_fence_after();
        if (tr && lane == 0) trace[32 + c] = clock64();
        const uint32_t al = a_base + (uint32_t)((s * kAStage) >> 4), bl = b_base + (uint32_t)((s * kBStage) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < kCK / 16; ++j) {
            const uint32_t o = (uint32_t)(j * kKStepBytes) >> 4;
            mma_f16(tmem, mk(al + o), mk(bl + o), idesc, (c | j) != 0);
            mma_f16(tmem, mk(al + o), mk(bl + (kBHalf >> 4) + o), idesc, 1);
            mma_f16(tmem, mk(al + (kAHalf >> 4) + o), mk(bl + o), idesc, 1);
          }
          mma_commit(&sh->empty[s]);
          if (c == kChunks - 1) mma_commit(&sh->acc_done);
        }
        __syncwarp();
      }
    }
  } else {
    // ================= workers: layer-1 row gather, then layer 3 out of TMEM =================
    const int part = warp >> 2, wq = warp & 3;  // hi / lo halves; rows [32 wq, 32 wq + 32)
    const int piece = lane & 3;
    const uint4* grow[4];
    uint32_t doff[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int rj = 32 * wq + 8 * j + (lane >> 2);
      const int bj = r0 + (rj < nrows ? rj : 0);
      int cj;
      if (src.cell_index && src.node_index) {
        cj = src.cell_index[bj];
      } else {
        const uint8_t* stj = src.compact + (src.node_index ? ((size_t)src.node_index[bj] * B + bj) : (size_t)bj) * env.compact_bytes;
        cj = deepsea_obs_index(*reinterpret_cast<const uint32_t*>(stj), env.size);
      }
      grow[j] = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(tw.h1[head]) + ((size_t)cj * 2 + part) * (kH * 2) + piece * 16);
      doff[j] = (uint32_t)(part * kAHalf + tile_offset_h32(rj, piece * 8));
    }
    int seen = 0;  // UBE head: the novelty bit of this thread's own row (fully_connected.py:83-90), fetched early
    if (part == 0 && head == EAZ_HEAD_UBE && live) {
      int cell;
      if (src.cell_index && src.node_index) cell = src.cell_index[b];
      else cell = deepsea_obs_index(*reinterpret_cast<const uint32_t*>(src.compact + (src.node_index ? ((size_t)src.node_index[b] * B + b) : (size_t)b) * env.compact_bytes), env.size);
      if (src.ds_seen) {
        seen = src.ds_seen[cell];
      } else {  // hash of the one-hot row (hashes.py:162-229)
        const int Lq = net.hash_dim >> 2;
        uint32_t a[4];
#pragma unroll
        for (int l = 0; l < 4; ++l) a[l] = xx_init(l);
        const int kbeg0 = net.D - net.hash_dim;
        for (int i = 0; i < Lq; ++i)
#pragma unroll
          for (int l = 0; l < 4; ++l) a[l] = xx_round(a[l], (kbeg0 + l * Lq + i == cell) ? EAZ_XX_ONE : 0u);
        const uint32_t idx = xx_finish(a[0], a[1], a[2], a[3], Lq, net.hash_bits);
        seen = (net.bset[idx >> 3] >> (idx & 7u)) & 1u;
      }
    }
    if (tr && threadIdx.x == 0) trace[3] = clock64();
    uint4 v[3][4];  // chunks in flight: two ahead of the one being stored
    auto load_a = [&](int c, uint4 (&x)[4]) {
#pragma unroll
      for (int j = 0; j < 4; ++j) x[j] = __ldg(grow[j] + c * (kCK * 2 / 16));
    };
    auto store_a = [&](int c, const uint4 (&x)[4]) {
      const int s = c % kStages;
      if (c >= kStages) {  // the stage is free once the MMAs of chunk c - kStages have completed; one polling lane per warp
        mbar_wait_warp(&sh->empty[s], ((c / kStages) & 1) ^ 1);
      }
      uint8_t* dst = sA + s * kAStage;
#pragma unroll
      for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(dst + doff[j]) = x[j];
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->full_a[s]);
    };
    load_a(0, v[0]);
    load_a(1, v[1]);
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      if (c + 2 < kChunks) load_a(c + 2, v[(c + 2) % 3]);
      store_a(c, v[c % 3]);
    }
    if (tr && threadIdx.x == 0) trace[5] = clock64();

    // ---- layer 3 on the CUDA cores: y = relu(D + b2) @ W3; worker group `part` takes half of the 256 columns
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    mbar_wait_warp(&sh->acc_done, 0);  // one polling lane per warp, warp-uniform loop
    tc_fence_after();
    if (tr && threadIdx.x == 0) trace[6] = clock64();
    const int cbase = part * (kH / 2);
    uint32_t ra[16], rb[16];
    // 16 accumulator columns: h = relu(acc * unscale + b2), y[o] += h * W3[k][o]; biases / weights come as 16-byte vectors
    auto consume16 = [&](const uint32_t (&r)[16], int k0) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 bq = *reinterpret_cast<const float4*>(&sh->b2[k0 + 4 * q]);
        const float h0 = fmaxf(__fmaf_rn(__uint_as_float(r[4 * q + 0]), kUnscale, bq.x), 0.0f);
        const float h1 = fmaxf(__fmaf_rn(__uint_as_float(r[4 * q + 1]), kUnscale, bq.y), 0.0f);
        const float h2 = fmaxf(__fmaf_rn(__uint_as_float(r[4 * q + 2]), kUnscale, bq.z), 0.0f);
        const float h3 = fmaxf(__fmaf_rn(__uint_as_float(r[4 * q + 3]), kUnscale, bq.w), 0.0f);
#pragma unroll
        for (int o = 0; o < kOutMax; ++o) {
          if (o < nout) {
            const float4 w = *reinterpret_cast<const float4*>(&sh->w3t[o][k0 + 4 * q]);
            y3[o] = __fmaf_rn(h0, w.x, y3[o]);
            y3[o] = __fmaf_rn(h1, w.y, y3[o]);
            y3[o] = __fmaf_rn(h2, w.z, y3[o]);
            y3[o] = __fmaf_rn(h3, w.w, y3[o]);
          }
        }
      }
    };
    tmem_ld16(tmem + lane_base + (uint32_t)cbase, ra);
#pragma unroll
    for (int kk = 0; kk < kH / 2; kk += 32) {  // double-buffered: the next 16 columns are in flight while these are consumed
      tmem_ld_wait();
      tmem_ld16(tmem + lane_base + (uint32_t)(cbase + kk + 16), rb);
      consume16(ra, cbase + kk);
      tmem_ld_wait();
      if (kk + 32 < kH / 2) tmem_ld16(tmem + lane_base + (uint32_t)(cbase + kk + 32), ra);
      consume16(rb, cbase + kk + 16);
    }
    if (tr && threadIdx.x == 0) trace[7] = clock64();
    if (part == 1) *reinterpret_cast<float4*>(sh->part[row]) = make_float4(y3[0], y3[1], y3[2], y3[3]);
    asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 worker warps
    if (part == 0 && live) {
      const float4 o = *reinterpret_cast<const float4*>(sh->part[row]);
      y3[0] = __fadd_rn(y3[0], o.x); y3[1] = __fadd_rn(y3[1], o.y); y3[2] = __fadd_rn(y3[2], o.z); y3[3] = __fadd_rn(y3[3], o.w);
      if (policy) {
        float* logits = out.logits[head - EAZ_HEAD_EXPLOIT];
#pragma unroll
        for (int o2 = 0; o2 < kOutMax; ++o2)
          if (o2 < nout) logits[(size_t)b * nout + o2] = __fadd_rn(y3[o2], __ldg(net.b[head][2] + o2));
      } else {
        const float y = __fadd_rn(y3[0], __ldg(net.b[head][2]));
        if (head == EAZ_HEAD_VALUE) {
          out.value[b] = eaz_tanh(y);
        } else {  // fully_connected.py:92-96
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
  if (src.mlp_done && threadIdx.x == 0) signal_counter(src.mlp_done + blockIdx.x);  // (after the barrier: one cumulative release fence)
  if (warp == 8) tmem_dealloc(tmem, kH);
  if (tr && threadIdx.x == 0) trace[10] = clock64();
  if (tl_on) {
    const unsigned long long i = atomicAdd(tl, 1ull);
    if (i < 2000) { tl[8 + 4 * i] = tl_entry; tl[9 + 4 * i] = tl_wait; tl[10 + 4 * i] = globaltimer_ns(); tl[11 + 4 * i] = 1; }
  }
}

// ---------------------------------------------------------------- host side
size_t gather_table_bytes(const NetDesc& net) { return (size_t)net.D * 2 * kH * sizeof(__half); }

int prepare_gather_table(const NetDesc& net, int head, void* buf, uint32_t* num_flags, cudaStream_t st) {
  const int total = net.D * kH;
  h1_table_kernel<<<min(ceil_div(total, 256), 148 * 8), 256, 0, st>>>(net.w[head][0], net.b[head][0], net.D, (__half*)buf, num_flags);
  EAZ_CHECK_LAUNCH("h1_table_kernel");
  return 0;
}

int launch_mlp_gather(const NetDesc& net, const EnvDesc& env, const MlpSource& src, const TensorWeights& tw, int B, int heads_mask,
                      const MlpOutputs& out, cudaStream_t stream) {
  GatherHeads hl{0, {0, 0, 0, 0}};
  for (int h = 0; h < 4; ++h)
    if (heads_mask & (1 << h)) hl.head[hl.n++] = h;
  if (hl.n == 0 || B == 0) return 0;
  // Ring depth: 4 stages (197 KB: one CTA per SM) when the launch fits the machine in one wave, 2 stages (101 KB: two CTAs per SM, 256 TMEM
  // columns each) when there are more (tile, head) units than SMs -- BASELINE C4: 64 tiles x 3 heads = 192 CTAs would otherwise run as
  // 148 + 44.  Measured at C4 (ms / step, same box): 4 stages 5.17 (one stream) / 4.97 (3 sub-batch streams of 66 CTAs), 2 stages 4.86 / 5.05.
  static const int force = getenv("EAZ_GATHER_STAGES") ? atoi(getenv("EAZ_GATHER_STAGES")) : 0;  // measurement knob
  const int units = ceil_div(B, kTM) * hl.n;
  const bool two = force ? force == 2 : units > 148;
  static bool attr_set = false;  // idempotent; a race only repeats the calls
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(mlp_gather_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(4));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_gather_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(2));
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mlp_gather_kernel)");
    attr_set = true;
  }
  cudaError_t le = two ? launch_pdl(mlp_gather_kernel<2>, dim3(ceil_div(B, kTM), hl.n), dim3(320), smem_bytes(2), stream, net, env, src, tw, B, hl, out, g_timeline,
                                    g_gather_trace)
                       : launch_pdl(mlp_gather_kernel<4>, dim3(ceil_div(B, kTM), hl.n), dim3(320), smem_bytes(4), stream, net, env, src, tw, B, hl, out, g_timeline,
                                    g_gather_trace);
  if (le != cudaSuccess) return cuda_fail(le, "mlp_gather_kernel launch");
  return 0;
}

}  // namespace eaz

// Debug hook (not in the public header): device buffer of >= 64 u64 receiving the milestones of CTA (0,0,0).
extern "C" void eaz_debug_set_gather_trace(unsigned long long* device_buffer) { eaz::g_gather_trace = device_buffer; }

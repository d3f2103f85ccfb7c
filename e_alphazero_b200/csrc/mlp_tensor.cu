// mlp_tensor.cu -- tensor-core (tcgen05) evaluation of EpistemicFullyConnectedAZNet
// (network/fully_connected.py:41-101) for the expand step of the search.
//
// One CTA = one head x 128 rows (nodes).  All three hk.Linear layers of the head run as
// tcgen05.mma kind::f16 with scaled 3xFP16 split precision (umma.cuh: x*S = hi + lo, three products, S a power of
// two removed exactly in the epilogue; same ~22-bit accuracy as 3xTF32 at half the bytes and twice the MMA rate)
// and fp32 accumulators in TMEM:
//   D1 (TMEM cols   0..255) = x  @ W1      Subleq / dense observations only: x is 0/1, exact in FP16,
//                                          so two products (x*W1_hi + x*W1_lo) suffice.  One-hot DeepSea
//                                          observations skip the GEMM: h1 = relu(W1[cell] + b1) is a row gather.
//   D2 (TMEM cols 256..511) = h1 @ W2      h1 = relu(D1 + b1), read back with tcgen05.ld by the 4 worker warps
//   D3 (TMEM cols   0..Np-1) = h2 @ W3     (thread = row), split hi/lo and written to the A stage in shared memory
// Weights are pre-split and pre-tiled once per search (tile_weights_kernel) so that a K-chunk of B is one
// contiguous block fetched with a single 1-D bulk async copy (UBLKCP) that signals an mbarrier.
// Warp roles: warps 0-7 = two producer groups (thread == row; group g produces chunks t % 2 == g, prefetching its
// next chunk's operands before storing the current one) that also run the epilogue, warp 8 issues the MMAs,
// warp 9 issues the weight copies.  mbarrier pipeline of K=32 chunks, A 16 KB + B 32 KB per stage (kStages below).
//
// Accuracy: <= ~1e-6 relative to the fp32 EXACT contract (tests: 1e-5); NOT bit-identical to it, so search
// parity in this mode is proven by replaying the GPU's per-node network outputs through the oracle.
#include "mlp.cuh"
#include "umma.cuh"

namespace eaz {
using namespace umma;

int launch_weight_scales(const NetDesc& net, int heads_mask, NumStatus* ns, cudaStream_t st);                                                     // tile_weights.cu

constexpr int kTM = 128;                       // rows per CTA
constexpr int kCK = 32;                        // K elements (fp16) per pipeline stage: 64 B per row
#ifndef EAZ_MT_STAGES
#define EAZ_MT_STAGES 2
#endif
// Pipeline stages (A 16 KB + B 32 KB each).  TWO, measured: with four a CTA holds 205 KB of shared memory and owns its SM, so under the
// sub-batch streams (EAZ_FLAG_STREAMS) the tree kernel's blocks of another sub-batch cannot use the issue slots the ten network warps
// leave idle; with two (109 KB) four tree blocks -- or the next network CTA's prologue -- fit beside it.  C3 (3 streams) 4.72 -> 4.50 ms
// per step (three stages: 4.62), one stream 5.39 -> 5.38; C5 16 384 x 128: 102 -> 107 M simulations/s, 65 536 x 128: 106 -> 108.
constexpr int kStages = EAZ_MT_STAGES;
constexpr int kSBO = (kCK * 2 / 16) * kCoreBytes;  // 8-row group stride inside a chunk tile (512 B)
constexpr int kAHalf = kTM * kCK * 2;          // one of hi / lo
constexpr int kAStage = 2 * kAHalf;
constexpr int kBStageMax = 2 * 256 * kCK * 2;  // hi + lo, N = 256
// (kActScale = 16, mlp.cuh: activations |h| < 4094 representable, lo part normal down to 2^-7; the weight scale is a per-matrix power
//  of two chosen from max |w| -- tile_weights.cu -- and removed exactly in the epilogues)
constexpr int kH = 256;
constexpr int kLayerChunks = kH / kCK;         // chunks of layers 2 and 3
constexpr int kBitsWordsMax = 40;              // observation bit-strings cached in smem up to 40*32 bits per row
constexpr int kSimtOutMax = 4;                 // heads with <= 4 outputs run layer 3 on the CUDA cores (see epilogue)

struct TensorSmem {
  uint64_t full_a[kStages], full_b[kStages], empty[kStages], acc_done[3];
  uint32_t tmem_base;
  alignas(16) float bias[3][kH];
  alignas(16) float bias_s[2][kH];  // b1, b2 pre-multiplied by kActScale: relu(acc*un + b) * S == relu(fma(acc, un*S, b*S)) exactly (S = 2^k)
  alignas(16) float w3[kH][kSimtOutMax];    // layer-3 weights of narrow heads
  alignas(16) float part[kTM][kSimtOutMax];  // partial dot products of producer group 1
  alignas(16) int16_t example[6][16];        // Subleq: example input [0..8) / example output [8..16) words per task row (test case 0)
};
constexpr size_t kTensorSmemFixed = (size_t)kStages * (kAStage + kBStageMax) + sizeof(TensorSmem) + 128;

__device__ __forceinline__ int sq_word_g(const uint8_t* st, int row, int ws, int trow) {
  if (row < ws) return st[EAZ_SQ_HDR + row];
  const int r = row - ws, part = r >> 3, i = r & 7;
  const uint16_t* h = reinterpret_cast<const uint16_t*>(st);
  if (part == 0) return sq_test_in(trow, 0, i, ws);
  if (part == 1) return h[i];
  if (part == 2) return sq_test_out(trow, 0, i, ws);
  return h[8 + i];
}
// same, with the example input / output rows taken from the per-CTA table (they cost a runtime modulo each otherwise)
__device__ __forceinline__ int sq_word_t(const uint8_t* st, int row, int ws, const int16_t* example) {
  if (row < ws) return st[EAZ_SQ_HDR + row];
  const int r = row - ws, part = r >> 3, i = r & 7;
  const uint16_t* h = reinterpret_cast<const uint16_t*>(st);
  if (part == 0) return example[i];
  if (part == 1) return h[i];
  if (part == 2) return example[8 + i];
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

// write kCK values (one K-chunk of this thread's row), already scaled, into the A stage as fp16 hi / lo tiles
__device__ __forceinline__ void store_a_chunk(uint8_t* stage, int row, const float (&v)[kCK], bool with_lo) {
#pragma unroll
  for (int q = 0; q < kCK / 8; ++q) {  // one 16-byte core-matrix row (8 halves) per store
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // two elements per packed conversion (cvt.rn.f16x2.f32)
      const float2 x = make_float2(v[8 * q + 2 * i], v[8 * q + 2 * i + 1]);
      const __half2 hi = __float22half2_rn(x);
      const float2 hf = __half22float2(hi);
      const __half2 lo = __float22half2_rn(make_float2(__fsub_rn(x.x, hf.x), __fsub_rn(x.y, hf.y)));
      h[i] = *reinterpret_cast<const uint32_t*>(&hi);
      l[i] = *reinterpret_cast<const uint32_t*>(&lo);
    }
    const int off = tile_offset_h32(row, q * 8);
    *reinterpret_cast<uint4*>(stage + off) = make_uint4(h[0], h[1], h[2], h[3]);
    if (with_lo) *reinterpret_cast<uint4*>(stage + kAHalf + off) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// Barrier wait for a whole warp with ONE polling lane: hundreds of threads spinning on mbarrier.try_wait starve the
// shared-memory pipe that the operand stores and TMEM loads of the other warps go through (measured in mlp_gather.cu).
__device__ __forceinline__ void warp_wait(uint64_t* bar, uint32_t parity) { mbar_wait_warp(bar, parity); }  // (umma.cuh: warp-uniform loop)

struct TensorHeads {
  int n;
  int head[4];
};

// Optional timeline trace (eaz_debug_set_mlp_trace): CTA (0,0) records clock64() per chunk.
// layout: [0]=start, [1]=end, [8+t]=MMA thread saw chunk t ready, [264+t]=producer lane 0 of the chunk's group arrived,
// [520+t]=MMA thread issued+committed chunk t
static unsigned long long* g_mlp_trace = nullptr;

__global__ void __launch_bounds__(320, 1) mlp_tensor_kernel(NetDesc net, EnvDesc env, MlpSource src, TensorWeights tw, int B, TensorHeads heads,
                                                            MlpOutputs out, int bits_words, unsigned long long* trace, unsigned long long* tl) {
  const bool tl_on = tl && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0;
  const unsigned long long tl_entry = tl_on ? globaltimer_ns() : 0ull;
  unsigned long long tl_wait = 0ull;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * kAStage;
  TensorSmem* sh = reinterpret_cast<TensorSmem*>(sB + kStages * kBStageMax);
  uint32_t* sbits = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(sh) + ((sizeof(TensorSmem) + 15) & ~15));  // [word][row]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = heads.head[blockIdx.y];
  const int r0 = blockIdx.x * kTM;
  const int nrows = min(kTM, B - r0);
  const bool gather = env.kind == EAZ_ENV_DEEPSEA && src.compact != nullptr;
  const bool policy = head >= EAZ_HEAD_EXPLOIT;
  const int nout = policy ? net.A : 1;
  const int np3 = (nout + 15) & ~15;
  const int n1 = gather ? 0 : tw.k1pad / kCK;  // layer-1 chunks
  // Narrow heads (value, UBE, 2-action policy): layer 3 is 256 x nout -- as tcgen05.mma it would cost the same ~100
  // issue slots as a 256-wide layer for no math, so it runs as fp32 FMAs straight out of the TMEM accumulator instead.
  const bool simt3 = nout <= kSimtOutMax;
  const int total = n1 + kLayerChunks + (simt3 ? 0 : kLayerChunks);
  const bool cached_bits = bits_words > 0;
  const bool tr = trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
  if (tr && threadIdx.x == 0) trace[0] = clock64();

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sh->full_a[s], 4);  // one elected arrive per producer warp of the group
      mbar_init(&sh->full_b[s], 1);
      mbar_init(&sh->empty[s], 1);
    }
    for (int i = 0; i < 3; ++i) mbar_init(&sh->acc_done[i], 1);
    fence_mbar_init();
  }
  if (threadIdx.x < kH) {  // biases and narrow-head W3 into shared memory: all loads issued before the first store
    const int j = threadIdx.x;
    const float b0 = __ldg(net.b[head][0] + j), b1 = __ldg(net.b[head][1] + j);
    const float b2 = j < nout ? __ldg(net.b[head][2] + j) : 0.0f;
    float w[kSimtOutMax] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (simt3) {
#pragma unroll
      for (int o = 0; o < kSimtOutMax; ++o)
        if (o < nout) w[o] = __ldg(net.w[head][2] + (size_t)j * nout + o);
    }
    sh->bias[0][j] = b0;
    sh->bias[1][j] = b1;
    sh->bias_s[0][j] = b0 * kActScale;
    sh->bias_s[1][j] = b1 * kActScale;
    sh->bias[2][j] = b2;
    if (simt3) *reinterpret_cast<float4*>(sh->w3[j]) = make_float4(w[0], w[1], w[2], w[3]);
  }
  if (env.kind == EAZ_ENV_SUBLEQ && threadIdx.x < 96) {  // example rows of Subleq._observe (subleq.py:698-704) for the 6 task rows
    const int tr6 = threadIdx.x >> 4, idx = threadIdx.x & 15;
    sh->example[tr6][idx] = (int16_t)(idx < 8 ? sq_test_in(tr6, 0, idx, env.ws) : sq_test_out(tr6, 0, idx - 8, env.ws));
  }
  if (warp == 8) {
    tmem_alloc(&sh->tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh->tmem_base;
  // the power-of-two scales the three weight images of this head carry (tile_weights.cu), removed exactly in the epilogues
  const float un_bits = 1.0f / __ldg(tw.wscale + 3 * head + 0);                 // layer 1 on 0/1 observations: A unscaled
  const float un_l2 = 1.0f / (kActScale * __ldg(tw.wscale + 3 * head + 1));   // layers fed by (16x-scaled) activations
  const float un_l3 = 1.0f / (kActScale * __ldg(tw.wscale + 3 * head + 2));
  // everything above reads only weights: under PDL it overlaps the tail of the previous kernel.  From here on the
  // kernel consumes the previous kernel's outputs (leaf indices, states).
  const int row = threadIdx.x & (kTM - 1);
  const bool live = row < nrows;
  const int b = r0 + (live ? row : 0);
  const uint8_t* st = nullptr;
  int cell = 0, trow = 0;
  // Every thread waits for the previous kernel (the tree / env step of this simulation) and only THEN lets the next
  // tree kernel launch: that kernel stages tree data before its own wait, which is safe only once the previous tree
  // kernel is complete (tree_step.cuh).
  pdl_wait();
  pdl_trigger();
  if (tl_on) tl_wait = globaltimer_ns();
  if (tr && threadIdx.x == 0) trace[2] = clock64();
  if (warp < 8 && src.compact) {
    const size_t slot = src.node_index ? ((size_t)src.node_index[b] * B + b) : (size_t)b;
    st = src.compact + slot * env.compact_bytes;
    if (env.kind == EAZ_ENV_DEEPSEA) cell = (src.cell_index && src.node_index) ? src.cell_index[b] : deepsea_obs_index(*reinterpret_cast<const uint32_t*>(st), env.size);
    else trow = sq_task_row(st[34]);
  }

  if (warp == 9) {
    // ================= weight-copy warp =================
    if (lane == 0) {
      for (int t = 0; t < total; ++t) {
        const int s = t % kStages, ph = (t / kStages) & 1;
        const int layer = t < n1 ? 0 : (t < n1 + kLayerChunks ? 1 : 2);
        const int c = layer == 0 ? t : (layer == 1 ? t - n1 : t - n1 - kLayerChunks);
        const uint32_t bbytes = (uint32_t)(2 * (layer == 2 ? np3 : kH) * kCK * 2);
        mbar_wait(&sh->empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&sh->full_b[s], bbytes);
        bulk_g2s(sB + s * kBStageMax, tw.img[head][layer] + (size_t)c * (bbytes / 4), bbytes, &sh->full_b[s]);
      }
    }
    __syncwarp();
  } else if (warp == 8) {
    // ================= MMA-issue warp =================
    // The whole warp runs the loop on warp-uniform values and one ELECTED lane issues (umma.cuh: elect_one): a chunk costs two
    // mbarrier waits, 4-6 tcgen05.mma with uniform-register descriptors and one commit.
    {
      const uint32_t desc_hi = (uint32_t)(kSBO >> 4) | (1u << 14);              // SBO [32,46) + version=1 [46,48)
      const uint32_t lbo_bits = (uint32_t)(kCoreBytes >> 4) << 16;              // LBO [16,30)
      uint32_t a_lo32[kStages], b_lo32[kStages];
#pragma unroll
      for (int s = 0; s < kStages; ++s) {
        a_lo32[s] = ((smem_u32(sA + s * kAStage) & 0x3FFFFu) >> 4) | lbo_bits;
        b_lo32[s] = ((smem_u32(sB + s * kBStageMax) & 0x3FFFFu) >> 4) | lbo_bits;
      }
      auto mk = [&](uint32_t lo) { return ((uint64_t)desc_hi << 32) | (uint64_t)lo; };
      int t = 0;
#pragma unroll 1
      for (int layer = 0; layer < 3; ++layer) {
        const int nchunks = layer == 0 ? n1 : ((layer == 2 && simt3) ? 0 : kLayerChunks);
        const int npad = layer == 2 ? np3 : kH;
        const uint32_t idesc = idesc_f16(kTM, npad);
        const uint32_t d = tmem + (layer == 1 ? 256u : 0u);
        const uint32_t blo_off = (uint32_t)(npad * kCK * 2) >> 4, alo_off = (uint32_t)kAHalf >> 4;
#pragma unroll 1
        for (int c = 0; c < nchunks; ++c, ++t) {
          const int s = t % kStages, ph = (t / kStages) & 1;
          mbar_wait(&sh->full_b[s], ph);
          mbar_wait(&sh->full_a[s], ph);
          tc_fence_after();
          if (tr && lane == 0) trace[8 + t] = clock64();
          uint32_t al = 0, bl = 0;
#pragma unroll
          for (int q = 0; q < kStages; ++q)
            if (q == s) { al = a_lo32[q]; bl = b_lo32[q]; }
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < kCK / 16; ++j) {
              const uint32_t o = (uint32_t)(j * kKStepBytes) >> 4;
              mma_f16(d, mk(al + o), mk(bl + o), idesc, (c | j) != 0);
              mma_f16(d, mk(al + o), mk(bl + blo_off + o), idesc, 1);
              if (layer != 0) mma_f16(d, mk(al + alo_off + o), mk(bl + o), idesc, 1);  // 0/1 inputs are exact in FP16: no lo part
            }
            mma_commit(&sh->empty[s]);
            if (c == nchunks - 1) mma_commit(&sh->acc_done[layer]);
          }
          __syncwarp();
          if (tr && lane == 0) trace[520 + t] = clock64();
        }
      }
    }
  } else {
    // ================= worker warps: A producer (thread == row) + epilogue =================
    // Stage this thread's compact state record in shared memory first (the A stages are still unused): the observation
    // builders below read it byte by byte, and from global memory every one of those ~50 dependent byte loads cost an L2
    // round trip (16.7k of the kernel's 47.7k cycles at C3, profiles/trace_mlp_tensor.py).  Vector loads, issued together.
    bool state_staged = false;
    if (cached_bits && st && env.kind == EAZ_ENV_SUBLEQ && env.compact_bytes <= 64) {
      uint2 piece[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) piece[i] = 8 * i < env.compact_bytes ? __ldg(reinterpret_cast<const uint2*>(st) + i) : make_uint2(0u, 0u);
      uint8_t* slot = sA + (size_t)threadIdx.x * 64;  // 256 worker threads x 64 B inside A stage 0
#pragma unroll
      for (int i = 0; i < 8; ++i) reinterpret_cast<uint2*>(slot)[i] = piece[i];
      st = slot;
      trow = sq_task_row(st[34]);
      state_staged = true;
    }
    if (tr && threadIdx.x == 0) trace[6] = clock64();
    // ONE thread per row builds the row's bit-string (producer group 0; thread r and thread 128 + r serve the same row).  Both groups used
    // to build it redundantly, which is harmless for the paths that store whole words but RACED in the one-hot path below: it zeroes the
    // row's words and then ORs bits in, so the two threads' read-modify-writes of one word could interleave and drop a bit (A sets bit
    // r1, B zeroes the word, A reads 0, B sets r1, A writes r2: r1 is gone) -- rarely, and only for non-binary encodings: a few rows of
    // one head's outputs off by ~3e-2 in up to 15 % of the searches on some boxes (profiles/stress_tensor_subleq.py).
    if (cached_bits && warp < 4) {  // this row's whole observation as a bit-string in shared memory (Subleq._observe, subleq.py:679-707)
      const int w = env.obs_cols, ws = env.ws;
      if (state_staged && env.binary && ws == 16 && w == 5) {
        // subleq-16 (the reference's own experiment size): 48 rows x 5 bits, and for v in [0, 16] the 5-bit pattern of
        // subleq.py:88-97 IS v (low 4 bits, or bit 4 for the pad value 16) -- a straight-line pack of 48 fields, no branches
        unsigned long long acc[4] = {0ull, 0ull, 0ull, 0ull};
        const uint4 h0 = *reinterpret_cast<const uint4*>(st), h1 = *reinterpret_cast<const uint4*>(st + 16);   // in_after, out_after (u16 x 8 each)
        const uint2 m0 = *reinterpret_cast<const uint2*>(st + EAZ_SQ_HDR), m1 = *reinterpret_cast<const uint2*>(st + EAZ_SQ_HDR + 8);  // memory bytes 0..15
        const uint4 mw = make_uint4(m0.x, m0.y, m1.x, m1.y);
        const uint4 e0 = *reinterpret_cast<const uint4*>(sh->example[trow]), e1 = *reinterpret_cast<const uint4*>(sh->example[trow] + 8);
        const uint32_t mem4[4] = {mw.x, mw.y, mw.z, mw.w}, in4[4] = {h0.x, h0.y, h0.z, h0.w}, out4[4] = {h1.x, h1.y, h1.z, h1.w};
        const uint32_t ei4[4] = {e0.x, e0.y, e0.z, e0.w}, eo4[4] = {e1.x, e1.y, e1.z, e1.w};
#pragma unroll
        for (int r = 0; r < 48; ++r) {
          uint32_t v;
          if (r < 16) v = (mem4[r >> 2] >> (8 * (r & 3))) & 0xffu;
          else {
            const int i = (r - 16) & 7, part = (r - 16) >> 3;
            const uint32_t word = part == 0 ? ei4[i >> 1] : part == 1 ? in4[i >> 1] : part == 2 ? eo4[i >> 1] : out4[i >> 1];
            v = (word >> (16 * (i & 1))) & 0xffffu;
          }
          const int bit = 5 * r, q = bit >> 6, lo = bit & 63;
          acc[q] |= (unsigned long long)(v & 0x1fu) << lo;
          if (lo > 59) acc[q + 1] |= (unsigned long long)(v & 0x1fu) >> (64 - lo);
        }
#pragma unroll
        for (int wd = 0; wd < 8; ++wd)
          if (wd < bits_words) sbits[wd * kTM + row] = live ? (uint32_t)(acc[wd >> 1] >> (32 * (wd & 1))) : 0u;
        for (int wd = 8; wd < bits_words; ++wd) sbits[wd * kTM + row] = 0u;
      } else if (st && env.binary) {  // words are in [0, ws] for any reachable state: v % ws == (v == ws ? 0 : v), no division
        unsigned long long acc = 0ull;
        int fill = 0, wd = 0;
        for (int orow = 0; orow < ws + 32; ++orow) {
          const int v = sq_word_t(st, orow, ws, sh->example[trow]);
          const unsigned pat = (v == ws) ? (1u << (w - 1)) : ((unsigned)v & 0xffu);  // subleq.py:88-97
          if (live) acc |= (unsigned long long)pat << fill;
          fill += w;
          if (fill >= 32) {
            sbits[wd * kTM + row] = (uint32_t)acc;
            acc >>= 32;
            fill -= 32;
            ++wd;
          }
        }
        for (; wd < bits_words; ++wd) {
          sbits[wd * kTM + row] = (uint32_t)acc;
          acc >>= 32;
        }
      } else {
        for (int wd = 0; wd < bits_words; ++wd) sbits[wd * kTM + row] = 0u;
        if (live && st) {  // one-hot rows (subleq.py:51-55)
          for (int orow = 0, pos = 0; orow < ws + 32; ++orow, pos += w) {
            const int v = sq_word_t(st, orow, ws, sh->example[trow]);
            const int k = pos + (v == ws ? ws : floormod(v, ws));
            sbits[(k >> 5) * kTM + row] |= 1u << (k & 31);
          }
        } else if (live) {
          const uint8_t* o = src.dense + (size_t)b * net.D;
          for (int k = 0; k < net.D; ++k)
            if (o[k]) sbits[(k >> 5) * kTM + row] |= 1u << (k & 31);
        }
      }
    }
    if (tr && threadIdx.x == 0) trace[7] = clock64();
    // every worker is done with its staged record (A stage 0 may be overwritten) and group 0's bit-strings are visible to group 1
    if (state_staged || cached_bits) asm volatile("bar.sync 1, 256;" ::: "memory");
    if (tr && threadIdx.x == 0) trace[776] = clock64();
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int grp = warp >> 2;  // two producer groups: group g produces chunks t with t % 2 == g
    bool waited[2] = {false, false};
    auto layer_of = [&](int t) { return t < n1 ? 0 : (t < n1 + kLayerChunks ? 1 : 2); };
    auto chunk_of = [&](int t) { return t < n1 ? t : (t < n1 + kLayerChunks ? t - n1 : t - n1 - kLayerChunks); };
    // stage 1 of producing chunk t: issue the loads (global gather / TMEM read / bit-string word); raw values in r[]
    auto load_raw = [&](int t, uint32_t (&r)[kCK]) {
      const int layer = layer_of(t), c = chunk_of(t);
      if (layer == 0) {  // observation bits of K-chunk c
        const int k0 = c * kCK;
        if (cached_bits) {
          const uint32_t wd = sbits[(k0 >> 5) * kTM + row];
#pragma unroll
          for (int i = 0; i < kCK; ++i) r[i] = ((wd >> i) & 1u) ? 0x3F800000u : 0u;
        } else if (st) {
          const int w = env.obs_cols, ws = env.ws;
          int orow = k0 / w, oc = k0 - orow * w;
          int word = (live && k0 < net.D) ? sq_word_g(st, orow, ws, trow) : 0;
#pragma unroll
          for (int i = 0; i < kCK; ++i) {
            const bool in = live && (k0 + i) < net.D;
            r[i] = (in && sq_bit_g(word, oc, w, ws, env.binary)) ? 0x3F800000u : 0u;
            if (++oc == w) {
              oc = 0;
              ++orow;
              if (in && k0 + i + 1 < net.D) word = sq_word_g(st, orow, ws, trow);
            }
          }
        } else {
          const uint8_t* o = src.dense + (size_t)b * net.D + k0;
#pragma unroll
          for (int i = 0; i < kCK; ++i) r[i] = (live && k0 + i < net.D && o[i]) ? 0x3F800000u : 0u;
        }
      } else if (layer == 1 && gather) {  // W1[cell] row slice: one-hot observation
        const uint4* wrow = reinterpret_cast<const uint4*>(net.w[head][0] + (size_t)cell * kH + c * kCK);
#pragma unroll
        for (int q = 0; q < kCK / 4; ++q) {
          const uint4 x = live ? __ldg(wrow + q) : make_uint4(0u, 0u, 0u, 0u);
          r[4 * q + 0] = x.x; r[4 * q + 1] = x.y; r[4 * q + 2] = x.z; r[4 * q + 3] = x.w;
        }
      } else {  // previous layer's accumulator from TMEM
        if (!waited[layer - 1]) {
          warp_wait(&sh->acc_done[layer - 1], 0);
          tc_fence_after();
          waited[layer - 1] = true;
        }
        tmem_ld32(tmem + lane_base + (layer == 1 ? 0u : 256u) + c * kCK, r);
      }
    };
    // stage 2: bias + relu (layers 2, 3), split, store into the stage, signal the MMA warp
    auto finish_store = [&](int t, uint32_t (&r)[kCK]) {
      const int layer = layer_of(t), c = chunk_of(t), s = t % kStages, ph = (t / kStages) & 1;
      float v[kCK];
      if (layer == 0) {
#pragma unroll
        for (int i = 0; i < kCK; ++i) v[i] = __uint_as_float(r[i]);
      } else {
        const bool from_tmem = !(layer == 1 && gather);
        if (from_tmem) tmem_ld_wait();
        // accumulators carry the operand scales: layer 1 of bit inputs kWScale, everything else kActScale * kWScale
        const float un = (!from_tmem ? 1.0f : ((layer == 1) ? un_bits : un_l2)) * kActScale;
        const float4* bs = reinterpret_cast<const float4*>(&sh->bias_s[layer - 1][c * kCK]);
#pragma unroll
        for (int q = 0; q < kCK / 4; ++q) {
          const float4 bq = bs[q];
          v[4 * q + 0] = fmaxf(__fmaf_rn(__uint_as_float(r[4 * q + 0]), un, bq.x), 0.0f);
          v[4 * q + 1] = fmaxf(__fmaf_rn(__uint_as_float(r[4 * q + 1]), un, bq.y), 0.0f);
          v[4 * q + 2] = fmaxf(__fmaf_rn(__uint_as_float(r[4 * q + 2]), un, bq.z), 0.0f);
          v[4 * q + 3] = fmaxf(__fmaf_rn(__uint_as_float(r[4 * q + 3]), un, bq.w), 0.0f);
          // range guard: a hidden activation beyond the fp16 range of the scaled split is clamped AND reported (sticky flag)
          const float mx4 = fmaxf(fmaxf(v[4 * q + 0], v[4 * q + 1]), fmaxf(v[4 * q + 2], v[4 * q + 3]));
          if (!(mx4 <= 65504.0f)) {
            if (live) atomicOr(tw.num_flags, kNumActSaturated);
#pragma unroll
            for (int i = 0; i < 4; ++i) v[4 * q + i] = fminf(v[4 * q + i], 65504.0f);
          }
        }
        if (!live) {
#pragma unroll
          for (int i = 0; i < kCK; ++i) v[i] = 0.0f;
        }
      }
      warp_wait(&sh->empty[s], ph ^ 1);
      if (tr && threadIdx.x == 0 && t < 2) trace[778 + t] = clock64();
      store_a_chunk(sA + s * kAStage, row, v, layer != 0);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->full_a[s]);
      if (tr && (threadIdx.x & 127) == 0) trace[264 + t] = clock64();
    };
    // a chunk may be prefetched while the previous one is still being stored unless it needs an accumulator
    // that the MMA warp can only finish after that store (first TMEM-sourced chunks of a layer)
    auto can_prefetch = [&](int t) {
      const int layer = layer_of(t);
      if (layer == 0 || (layer == 1 && gather)) return true;
      return waited[layer - 1];
    };
    // Layer-1 chunks of cached observation bit-strings go straight from the 32-bit word to fp16 operand pieces (1.0 = 0x3C00, two
    // bits per 32-bit word: 0x3C00 * bit0 + 0x3C000000 * bit1); no lo tile (0 / 1 are exact in fp16, the MMA warp skips that product)
    auto bits_chunk = [&](int t) { return cached_bits && t < n1; };
    auto store_bits = [&](int t) {
      const int s = t % kStages, ph = (t / kStages) & 1;
      const uint32_t wd = live ? sbits[((t * kCK) >> 5) * kTM + row] : 0u;
      warp_wait(&sh->empty[s], ph ^ 1);
      uint8_t* stage = sA + s * kAStage;
#pragma unroll
      for (int q = 0; q < kCK / 8; ++q) {
        uint32_t h[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t two = (wd >> (8 * q + 2 * i)) & 3u;
          h[i] = (two & 1u) * 0x3C00u + (two >> 1) * 0x3C000000u;
        }
        *reinterpret_cast<uint4*>(stage + tile_offset_h32(row, q * 8)) = make_uint4(h[0], h[1], h[2], h[3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->full_a[s]);
      if (tr && (threadIdx.x & 127) == 0) trace[264 + t] = clock64();
    };
    uint32_t cur[kCK], nxt[kCK];
    int t = grp;
    if (t < total && !bits_chunk(t)) load_raw(t, cur);
    if (tr && threadIdx.x == 0) trace[777] = clock64();
    for (; t < total; t += 2) {
      const int tn = t + 2;
      if (bits_chunk(t)) {
        store_bits(t);
        if (tn < total && !bits_chunk(tn)) load_raw(tn, cur);
        continue;
      }
      const bool pre = tn < total && can_prefetch(tn);
      if (pre) load_raw(tn, nxt);
      finish_store(t, cur);
      if (tn < total && !pre) load_raw(tn, nxt);
#pragma unroll
      for (int i = 0; i < kCK; ++i) cur[i] = nxt[i];
    }

    // ---- novelty probe (UBE head): fully_connected.py:83-90
    int seen = 0;
    if (grp == 0 && head == EAZ_HEAD_UBE && live) {
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
        } else if (cached_bits) {
          for (int i = 0; i < L; ++i)
#pragma unroll
            for (int l = 0; l < 4; ++l) {
              const int k = kbeg0 + l * L + i;
              a[l] = xx_round(a[l], ((sbits[(k >> 5) * kTM + row] >> (k & 31)) & 1u) ? EAZ_XX_ONE : 0u);
            }
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

    // ---- layer 3 on the CUDA cores for narrow heads: y = relu(D2 + b2) @ W3, each producer group takes half of K
    float y3[kSimtOutMax] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (simt3) {
      if (!waited[1]) {
        warp_wait(&sh->acc_done[1], 0);
        tc_fence_after();
        waited[1] = true;
      }
      if (tr && threadIdx.x == 0) trace[3] = clock64();
      const int kbase = grp * (kH / 2);
#pragma unroll 2
      for (int kk = 0; kk < kH / 2; kk += 16) {
        uint32_t r[16];
        tmem_ld16(tmem + lane_base + 256u + kbase + kk, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int k = kbase + kk + i;
          const float h = fmaxf(__fmaf_rn(__uint_as_float(r[i]), un_l2, sh->bias[1][k]), 0.0f);
          if (nout == 1) {  // value / UBE heads
            y3[0] = __fmaf_rn(h, sh->w3[k][0], y3[0]);
          } else {
            const float4 w = *reinterpret_cast<const float4*>(sh->w3[k]);
            y3[0] = __fmaf_rn(h, w.x, y3[0]);
            y3[1] = __fmaf_rn(h, w.y, y3[1]);
            y3[2] = __fmaf_rn(h, w.z, y3[2]);
            y3[3] = __fmaf_rn(h, w.w, y3[3]);
          }
        }
      }
      if (grp == 1) *reinterpret_cast<float4*>(sh->part[row]) = make_float4(y3[0], y3[1], y3[2], y3[3]);
      if (tr && threadIdx.x == 0) trace[4] = clock64();
      asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 producer warps
      if (tr && threadIdx.x == 0) trace[5] = clock64();
      if (grp == 0) {
        const float4 o = *reinterpret_cast<const float4*>(sh->part[row]);
        y3[0] = __fadd_rn(y3[0], o.x); y3[1] = __fadd_rn(y3[1], o.y); y3[2] = __fadd_rn(y3[2], o.z); y3[3] = __fadd_rn(y3[3], o.w);
      }
    } else {
      warp_wait(&sh->acc_done[2], 0);
      tc_fence_after();
    }
    // ---- head epilogue
    float* logits = policy ? out.logits[head - EAZ_HEAD_EXPLOIT] : nullptr;
    for (int n0 = 0; grp == 0 && n0 < np3; n0 += 16) {  // group 0 (warps 0-3) owns the outputs
      uint32_t r[16];
      if (simt3) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = i < kSimtOutMax ? __float_as_uint(y3[i < kSimtOutMax ? i : 0]) : 0u;
      } else {
        tmem_ld16(tmem + lane_base + n0, r);
        tmem_ld_wait();
      }
      if (!live) continue;
      if (policy) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (n0 + i < nout) logits[(size_t)b * nout + n0 + i] = __fadd_rn(simt3 ? __uint_as_float(r[i]) : __fmul_rn(__uint_as_float(r[i]), un_l3), sh->bias[2][n0 + i]);
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
  if (warp == 8) tmem_dealloc(tmem, 512);
  if (tr && threadIdx.x == 0) trace[1] = clock64();
  if (tl_on) {
    const unsigned long long i = atomicAdd(tl, 1ull);
    if (i < 2000) { tl[8 + 4 * i] = tl_entry; tl[9 + 4 * i] = tl_wait; tl[10 + 4 * i] = globaltimer_ns(); tl[11 + 4 * i] = 1; }
  }
}

// ---------------------------------------------------------------- host side
static int k1pad_of(int D) { return (D + kCK - 1) / kCK * kCK; }

size_t tensor_weights_bytes(const NetDesc& net, const EnvDesc& env) {
  const int np3 = (net.A + 15) & ~15;
  const size_t l1 = env.kind == EAZ_ENV_DEEPSEA ? gather_table_bytes(net) : (size_t)k1pad_of(net.D) * 2 * kH * 2;
  const size_t l2 = (size_t)kH * 2 * kH * 2;
  const size_t per_head = l1 + l2 + (size_t)kH * 2 * np3 * 2 + (env.kind == EAZ_ENV_DEEPSEA ? l2 : 0);  // (+ the K = 16 W2 images)
  return 4 * ((per_head + 255) & ~(size_t)255) + 256;  // (+ the numeric status block: per-matrix scales, saturation flags)
}

int prepare_tensor_weights(const NetDesc& net, const EnvDesc& env, int heads_mask, void* buf, TensorWeights* tw, cudaStream_t st, bool fill) {
  if (net.H != kH) {
    set_error("tensor network path needs hidden size %d (got %d); use mlp_mode EXACT", kH, net.H);
    return EAZ_ERR_UNSUPPORTED;
  }
  const int k1pad = k1pad_of(net.D);
  const bool has_l1 = env.kind != EAZ_ENV_DEEPSEA;
  const size_t l1 = has_l1 ? (size_t)k1pad * 2 * kH * 2 : gather_table_bytes(net), l2 = (size_t)kH * 2 * kH * 2;
  const size_t per_head = (tensor_weights_bytes(net, env) - 256) / 4;
  NumStatus* ns = reinterpret_cast<NumStatus*>((uint8_t*)buf + 4 * per_head);
  tw->k1pad = k1pad;
  tw->wscale = &ns->wscale[0][0];
  tw->num_flags = &ns->flags;
  if (fill)
    if (int rc = launch_weight_scales(net, heads_mask, ns, st)) return rc;
  for (int h = 0; h < 4; ++h) {
    uint8_t* p = (uint8_t*)buf + (size_t)h * per_head;
    tw->img[h][0] = (const uint32_t*)p;
    tw->img[h][1] = (const uint32_t*)(p + l1);
    tw->img[h][2] = (const uint32_t*)(p + l1 + l2);
    tw->h1[h] = has_l1 ? nullptr : p;
    const size_t l3 = (size_t)kH * 2 * ((net.A + 15) & ~15) * 2;
    tw->w2_ck16[h] = has_l1 ? nullptr : p + l1 + l2 + l3;
    if (!fill || !(heads_mask & (1 << h))) continue;
    const int nout = h >= EAZ_HEAD_EXPLOIT ? net.A : 1;
    if (has_l1) {
      if (int rc = launch_tile_weights_f16(net.w[h][0], net.D, kH, k1pad, kH, tw->wscale + 3 * h + 0, p, st)) return rc;
    } else {
      if (int rc = prepare_gather_table(net, h, p, tw->num_flags, st)) return rc;
    }
    if (int rc = launch_tile_weights_f16(net.w[h][1], kH, kH, kH, kH, tw->wscale + 3 * h + 1, p + l1, st)) return rc;
    if (int rc = launch_tile_weights_f16(net.w[h][2], kH, nout, kH, (nout + 15) & ~15, tw->wscale + 3 * h + 2, p + l1 + l2, st)) return rc;
    if (!has_l1)
      if (int rc = launch_tile_weights_f16(net.w[h][1], kH, kH, kH, kH, tw->wscale + 3 * h + 1, p + l1 + l2 + l3, st, 16)) return rc;
  }
  return 0;
}

int launch_mlp_tensor(const NetDesc& net, const EnvDesc& env, const MlpSource& src, const TensorWeights& tw, int B, int heads_mask,
                      const MlpOutputs& out, cudaStream_t stream) {
  TensorHeads hl{0, {0, 0, 0, 0}};
  for (int h = 3; h >= 0; --h)  // policy heads first: their CTAs are the longest (layer 3 on the tensor core), so a partial last wave holds short ones
    if (heads_mask & (1 << h)) hl.head[hl.n++] = h;
  if (hl.n == 0 || B == 0) return 0;
  const bool gather = env.kind == EAZ_ENV_DEEPSEA && src.compact != nullptr;
  if (gather && net.A <= 4) return launch_mlp_gather(net, env, src, tw, B, heads_mask, out, stream);  // one-hot rows: mlp_gather.cu
  int bits_words = gather ? 0 : (tw.k1pad + 31) / 32;
  if (bits_words > kBitsWordsMax) bits_words = 0;  // too wide for shared memory: bits are derived per chunk instead
  const size_t smem = kTensorSmemFixed + (size_t)bits_words * kTM * 4;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(mlp_tensor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kTensorSmemFixed + kBitsWordsMax * kTM * 4));
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mlp_tensor_kernel)");
    attr_smem = kTensorSmemFixed + kBitsWordsMax * kTM * 4;
  }
  dim3 grid(ceil_div(B, kTM), hl.n);
  cudaError_t le = launch_pdl(mlp_tensor_kernel, grid, dim3(320), smem, stream, net, env, src, tw, B, hl, out, bits_words, g_mlp_trace, g_timeline);
  if (le != cudaSuccess) return cuda_fail(le, "mlp_tensor_kernel launch");
  return 0;
}

}  // namespace eaz

// Debug hook (not in the public header): device buffer of >= 1024 u64 receiving the timeline of CTA (0,0).
extern "C" void eaz_debug_set_mlp_trace(unsigned long long* device_buffer) { eaz::g_mlp_trace = device_buffer; }

// hash.cu -- XXHash.get_indices / BaseHash.__call__ / BaseHash.update
// (network/hashes.py:23-50,162-229) on float32 rows.  Four threads per row,
// one per xxhash lane (lane l consumes the contiguous quarter l of the row).
#include "common.cuh"

namespace eaz {

// mode 0: write indices; 1: lookup -> seen; 2: update (atomic OR)
__global__ void xxhash_rows_kernel(const uint32_t* __restrict__ x, int B, int D, int bits, int mode, uint32_t* __restrict__ indices,
                                   const uint8_t* __restrict__ set_r, uint8_t* __restrict__ seen, uint32_t* __restrict__ set_w) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = t >> 2, lane = t & 3;
  const int L = D >> 2;
  uint32_t acc = xx_init(lane);
  if (b < B) {
    const uint32_t* p = x + (size_t)b * D + (size_t)lane * L;
    int i = 0;
    for (; i + 4 <= L; i += 4) {  // 4 independent loads in flight per thread
      const uint32_t w0 = p[i], w1 = p[i + 1], w2 = p[i + 2], w3 = p[i + 3];
      acc = xx_round(xx_round(xx_round(xx_round(acc, w0), w1), w2), w3);
    }
    for (; i < L; ++i) acc = xx_round(acc, p[i]);
  }
  const unsigned base = (threadIdx.x & 31u) & ~3u;
  const uint32_t a0 = __shfl_sync(0xffffffffu, acc, base + 0), a1 = __shfl_sync(0xffffffffu, acc, base + 1),
                 a2 = __shfl_sync(0xffffffffu, acc, base + 2), a3 = __shfl_sync(0xffffffffu, acc, base + 3);
  if (b >= B || lane != 0) return;
  const uint32_t idx = xx_finish(a0, a1, a2, a3, L, bits);
  if (mode == 0) indices[b] = idx;
  else if (mode == 1) seen[b] = (uint8_t)((set_r[idx >> 3] >> (idx & 7u)) & 1u);
  else atomicOr(set_w + (idx >> 5), 1u << (idx & 31u));  // byte idx>>3, bit idx&7 == word idx>>5, bit idx&31 (little endian)
}

static int launch(const float* x, int B, int D, int bits, int mode, uint32_t* indices, const uint8_t* set_r, uint8_t* seen,
                  uint8_t* set_w, void* stream, const char* what) {
  EAZ_CHECK_ARG(bits > 0 && bits <= 32, "bits_per_hash %d violates 0 < bits <= 32 (hashes.py:154)", bits);
  EAZ_CHECK_ARG(D > 0 && D % 4 == 0, "hash input length %d is not a positive multiple of 4 (hashes.py:210)", D);
  EAZ_CHECK_ARG(x != nullptr && B >= 0, "%s: bad arguments", what);
  if (mode == 2) EAZ_CHECK_ARG(((uintptr_t)set_w & 3u) == 0, "binary_set must be 4-byte aligned for update");
  if (B == 0) return 0;
  const int threads = 128;
  xxhash_rows_kernel<<<ceil_div(B * 4, threads), threads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint32_t*>(x), B, D, bits, mode,
                                                                                     indices, set_r, seen,
                                                                                     reinterpret_cast<uint32_t*>(set_w));
  EAZ_CHECK_LAUNCH(what);
  return 0;
}

}  // namespace eaz

extern "C" {
int eaz_xxhash_indices(const float* x, int32_t B, int32_t D, int32_t bits, uint32_t* indices, void* stream) {
  EAZ_CHECK_ARG(indices != nullptr, "indices is NULL");
  return eaz::launch(x, B, D, bits, 0, indices, nullptr, nullptr, nullptr, stream, "eaz_xxhash_indices");
}
int eaz_hash_lookup(const float* x, int32_t B, int32_t D, int32_t bits, const uint8_t* binary_set, uint8_t* seen, void* stream) {
  EAZ_CHECK_ARG(binary_set && seen, "binary_set / seen is NULL");
  return eaz::launch(x, B, D, bits, 1, nullptr, binary_set, seen, nullptr, stream, "eaz_hash_lookup");
}
int eaz_hash_update(const float* x, int32_t B, int32_t D, int32_t bits, uint8_t* binary_set, void* stream) {
  EAZ_CHECK_ARG(binary_set != nullptr, "binary_set is NULL");
  return eaz::launch(x, B, D, bits, 2, nullptr, nullptr, nullptr, binary_set, stream, "eaz_hash_update");
}
}

// tile_weights.cu -- parameter-derived weight images of the tensor-core network kernels (built once per model):
// haiku [in,out] fp32 matrices -> split-precision (hi / lo), K-chunked images in the shared-memory operand layout of umma.cuh,
// so that a kernel fetches one chunk with ONE 1-D bulk async copy and hands it to tcgen05.mma unchanged.
#include "common.cuh"
#include "umma.cuh"

namespace eaz {
using namespace umma;

// W[K][N] (row-major, haiku layout) -> per K-chunk image [hi tile | lo tile], each tile [Npad x 32] in the
// canonical K-major layout of umma.cuh.  out must hold (Kpad/32) * 2 * Npad * 32 words.
template <int CK>
__global__ void tile_weights_kernel(const float* __restrict__ W, int K, int N, int Kpad, int Npad, uint32_t* __restrict__ out) {
  const int total = Kpad * Npad;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i / Npad, n = i % Npad;  // consecutive threads -> consecutive n (coalesced reads of W rows)
    const float w = (k < K && n < N) ? W[(size_t)k * N + n] : 0.0f;
    uint32_t hi, lo;
    split_tf32(w, hi, lo);
    const int c = k / CK, kk = k % CK;
    const size_t base = (size_t)c * 2 * Npad * CK;  // words
    const int off = tile_offset_ck<CK>(n, kk) >> 2;
    out[base + off] = hi;
    out[base + (size_t)Npad * CK + off] = lo;
  }
}

// fp16 variant for the network kernels: chunks of CK k (32: mlp_gather.cu / mlp_tensor.cu, 16: psearch.cuh), image per chunk =
// [hi tile | lo tile], each [Npad x CK] halves in the K-major core-matrix layout (8 rows x 16 bytes, CK / 8 core matrices per row
// group); values are scaled by `scale` (a power of two) and clamped to the fp16 range before the hi/lo split.
template <int CK>
__global__ void tile_weights_f16_kernel(const float* __restrict__ W, int K, int N, int Kpad, int Npad, float scale, __half* __restrict__ out) {
  const int total = Kpad * Npad;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i / Npad, n = i % Npad;
    float w = (k < K && n < N) ? W[(size_t)k * N + n] * scale : 0.0f;
    w = fminf(fmaxf(w, -65504.0f), 65504.0f);
    __half hi, lo;
    split_f16(w, hi, lo);
    const int c = k / CK, kk = k % CK;
    const size_t base = (size_t)c * 2 * Npad * CK;  // halves
    const int off = ((n >> 3) * (CK / 8) * kCoreBytes + (kk >> 3) * kCoreBytes + (n & 7) * 16 + (kk & 7) * 2) >> 1;
    out[base + off] = hi;
    out[base + (size_t)Npad * CK + off] = lo;
  }
}
int launch_tile_weights_f16(const float* W, int K, int N, int Kpad, int Npad, float scale, void* out, cudaStream_t st, int chunk_k) {
  if (chunk_k == 16) tile_weights_f16_kernel<16><<<ceil_div(Kpad * Npad, 256), 256, 0, st>>>(W, K, N, Kpad, Npad, scale, (__half*)out);
  else tile_weights_f16_kernel<32><<<ceil_div(Kpad * Npad, 256), 256, 0, st>>>(W, K, N, Kpad, Npad, scale, (__half*)out);
  EAZ_CHECK_LAUNCH("tile_weights_f16_kernel");
  return 0;
}

int launch_tile_weights(const float* W, int K, int N, int Kpad, int Npad, uint32_t* out, cudaStream_t st, int chunk_k) {
  if (chunk_k == 16) tile_weights_kernel<16><<<ceil_div(Kpad * Npad, 256), 256, 0, st>>>(W, K, N, Kpad, Npad, out);
  else tile_weights_kernel<32><<<ceil_div(Kpad * Npad, 256), 256, 0, st>>>(W, K, N, Kpad, Npad, out);
  EAZ_CHECK_LAUNCH("tile_weights_kernel");
  return 0;
}

}  // namespace eaz

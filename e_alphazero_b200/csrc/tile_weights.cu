// tile_weights.cu -- parameter-derived weight images of the tensor-core network kernels (built once per model):
// haiku [in,out] fp32 matrices -> split-precision (hi / lo), K-chunked images in the shared-memory operand layout of umma.cuh,
// so that a kernel fetches one chunk with ONE 1-D bulk async copy and hands it to tcgen05.mma unchanged.
#include "mlp.cuh"
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
__global__ void tile_weights_f16_kernel(const float* __restrict__ W, int K, int N, int Kpad, int Npad, const float* __restrict__ scale_ptr,
                                        __half* __restrict__ out) {
  const int total = Kpad * Npad;
  const float scale = *scale_ptr;  // power of two chosen from max |w| (weight_scales_kernel): nothing saturates below
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
int launch_tile_weights_f16(const float* W, int K, int N, int Kpad, int Npad, const float* scale, void* out, cudaStream_t st, int chunk_k) {
  if (chunk_k == 16) tile_weights_f16_kernel<16><<<ceil_div(Kpad * Npad, 256), 256, 0, st>>>(W, K, N, Kpad, Npad, scale, (__half*)out);
  else tile_weights_f16_kernel<32><<<ceil_div(Kpad * Npad, 256), 256, 0, st>>>(W, K, N, Kpad, Npad, scale, (__half*)out);
  EAZ_CHECK_LAUNCH("tile_weights_f16_kernel");
  return 0;
}

// ---- range guard of the scaled split: one power-of-two scale per weight matrix, from its own max |w|
struct WeightList {
  const float* w[4][3];
  int n[4][3];
};
__global__ void weight_max_kernel(WeightList wl, NumStatus* ns) {  // grid (blocks, 12): |w| as unsigned bits orders like the floats
  const int m = blockIdx.y, h = m / 3, l = m % 3;
  const float* w = wl.w[h][l];
  const int n = wl.n[h][l];
  if (!w || n <= 0) return;
  uint32_t mx = 0u;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) mx = max(mx, __float_as_uint(w[i]) & 0x7fffffffu);
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, s));
  if ((threadIdx.x & 31) == 0 && mx) atomicMax(&ns->wmax_bits[h][l], mx);
}
__global__ void weight_scales_kernel(NumStatus* ns) {
  const int m = threadIdx.x;
  if (m >= 12) return;
  const int h = m / 3, l = m % 3;
  const uint32_t bits = ns->wmax_bits[h][l];
  float scale = 256.0f;
  if (bits >= 0x7f800000u || __uint_as_float(bits) > 1048576.0f) {  // inf / nan / absurd: no scale keeps the split meaningful
    atomicOr(&ns->flags, kNumWeightsNonFinite);
    scale = 1.0f / 64.0f;
  } else if (bits) {
    // largest power of two with max|w| * scale <= 32768 (one binade below the fp16 limit), capped to [2^-6, 2^14]
    const int e = (int)(bits >> 23) - 127;  // max|w| in [2^e, 2^(e+1))
    int k = 14 - e;                         // 2^(e+1) * 2^k = 2^15
    k = k < -6 ? -6 : (k > 14 ? 14 : k);
    scale = __uint_as_float((uint32_t)(127 + k) << 23);
  }
  ns->wscale[h][l] = scale;
}
int launch_weight_scales(const NetDesc& net, int heads_mask, NumStatus* ns, cudaStream_t st) {
  if (cudaError_t e = cudaMemsetAsync(ns, 0, sizeof(NumStatus), st); e != cudaSuccess) return cuda_fail(e, "numeric status memset");
  WeightList wl{};
  for (int h = 0; h < 4; ++h)
    for (int l = 0; l < 3; ++l) {
      const bool on = (heads_mask >> h) & 1;
      const int nout = h >= EAZ_HEAD_EXPLOIT ? net.A : 1;
      wl.w[h][l] = on ? net.w[h][l] : nullptr;
      wl.n[h][l] = l == 0 ? net.D * net.H : (l == 1 ? net.H * net.H : net.H * nout);
    }
  weight_max_kernel<<<dim3(16, 12), 256, 0, st>>>(wl, ns);
  EAZ_CHECK_LAUNCH("weight_max_kernel");
  weight_scales_kernel<<<1, 32, 0, st>>>(ns);
  EAZ_CHECK_LAUNCH("weight_scales_kernel");
  return 0;
}

// the same for an arbitrary list of up to 12 matrices (convnet.cu): slot m -> ns->wscale[m / 3][m % 3]
int launch_weight_scales_list(const float* const* w, const int* n, int count, NumStatus* ns, cudaStream_t st) {
  if (cudaError_t e = cudaMemsetAsync(ns, 0, sizeof(NumStatus), st); e != cudaSuccess) return cuda_fail(e, "numeric status memset");
  WeightList wl{};
  for (int m = 0; m < 12; ++m) {
    wl.w[m / 3][m % 3] = m < count ? w[m] : nullptr;
    wl.n[m / 3][m % 3] = m < count ? n[m] : 0;
  }
  weight_max_kernel<<<dim3(16, 12), 256, 0, st>>>(wl, ns);
  EAZ_CHECK_LAUNCH("weight_max_kernel");
  weight_scales_kernel<<<1, 32, 0, st>>>(ns);
  EAZ_CHECK_LAUNCH("weight_scales_kernel");
  return 0;
}

int launch_tile_weights(const float* W, int K, int N, int Kpad, int Npad, uint32_t* out, cudaStream_t st, int chunk_k) {
  if (chunk_k == 16) tile_weights_kernel<16><<<ceil_div(Kpad * Npad, 256), 256, 0, st>>>(W, K, N, Kpad, Npad, out);
  else tile_weights_kernel<32><<<ceil_div(Kpad * Npad, 256), 256, 0, st>>>(W, K, N, Kpad, Npad, out);
  EAZ_CHECK_LAUNCH("tile_weights_kernel");
  return 0;
}

}  // namespace eaz

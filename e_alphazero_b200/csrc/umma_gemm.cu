// umma_gemm.cu -- bring-up / unit-test kernel for the tcgen05 building blocks of umma.cuh:
// D[128,N] = A[128,K] @ W[K,N] with 3xTF32 split precision on ONE CTA (4 producer/epilogue warps + 1 control
// warp, 2-stage mbarrier pipeline, A staged by the worker threads, B fetched with 1-D bulk copies from a
// pre-tiled weight image).  Exported as eaz_debug_umma_gemm (not part of the public header): tests compare it
// with an fp64 matmul.  The network kernel (mlp_tensor.cu) is built from the same pieces.
#include "common.cuh"
#include "umma.cuh"

namespace eaz {
using namespace umma;

int launch_tile_weights(const float* W, int K, int N, int Kpad, int Npad, uint32_t* out, cudaStream_t st, int chunk_k);  // tile_weights.cu

struct GemmSmem {
  uint64_t full_a[2], full_b[2], empty[2], done;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(160) umma_gemm_test_kernel(const float* __restrict__ A, const uint32_t* __restrict__ Wt, float* __restrict__ D,
                                                             int K, int N, int Npad) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // [A stage 0: hi 16K | lo 16K][A stage 1][B stage 0: hi | lo][B stage 1][barriers]
  const int a_stage = 2 * 128 * kChunkK * 4, b_stage = 2 * Npad * kChunkK * 4;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 2 * a_stage;
  GemmSmem* sh = reinterpret_cast<GemmSmem*>(sB + 2 * b_stage);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks = K / kChunkK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sh->full_a[s], 128);
      mbar_init(&sh->full_b[s], 1);
      mbar_init(&sh->empty[s], 1);
    }
    mbar_init(&sh->done, 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(&sh->tmem_base, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh->tmem_base;

  if (warp == 4) {
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(128, Npad);
      for (int c = 0; c < nchunks; ++c) {
        const int s = c & 1, ph = (c >> 1) & 1;
        mbar_wait(&sh->empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&sh->full_b[s], (uint32_t)b_stage);
        bulk_g2s(sB + s * b_stage, Wt + (size_t)c * (b_stage / 4), (uint32_t)b_stage, &sh->full_b[s]);
        mbar_wait(&sh->full_b[s], ph);
        mbar_wait(&sh->full_a[s], ph);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(sA + s * a_stage), a_lo = a_hi + 128 * kChunkK * 4;
        const uint32_t b_hi = smem_u32(sB + s * b_stage), b_lo = b_hi + Npad * kChunkK * 4;
#pragma unroll
        for (int j = 0; j < kKSteps; ++j) {
          const uint32_t o = j * kKStepBytes;
          mma_tf32(tmem, smem_desc(a_hi + o), smem_desc(b_hi + o), idesc, (c | j) != 0);
          mma_tf32(tmem, smem_desc(a_hi + o), smem_desc(b_lo + o), idesc, 1);
          mma_tf32(tmem, smem_desc(a_lo + o), smem_desc(b_hi + o), idesc, 1);
        }
        mma_commit(&sh->empty[s]);
      }
      mma_commit(&sh->done);
    }
    __syncwarp();
  } else {
    const int row = threadIdx.x;  // 0..127
    for (int c = 0; c < nchunks; ++c) {
      const int s = c & 1, ph = (c >> 1) & 1;
      mbar_wait(&sh->empty[s], ph ^ 1);
      uint8_t* hi = sA + s * a_stage;
      uint8_t* lo = hi + 128 * kChunkK * 4;
      const float4* src = reinterpret_cast<const float4*>(A + (size_t)row * K + c * kChunkK);
#pragma unroll
      for (int q = 0; q < kChunkK / 4; ++q) {
        const float4 v = src[q];
        uint4 h, l;
        split_tf32(v.x, h.x, l.x);
        split_tf32(v.y, h.y, l.y);
        split_tf32(v.z, h.z, l.z);
        split_tf32(v.w, h.w, l.w);
        const int off = tile_offset(row, q * 4);
        *reinterpret_cast<uint4*>(hi + off) = h;
        *reinterpret_cast<uint4*>(lo + off) = l;
      }
      fence_proxy_async();
      mbar_arrive(&sh->full_a[s]);
    }
    mbar_wait(&sh->done, 0);
    tc_fence_after();
    for (int n0 = 0; n0 < Npad; n0 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + n0, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (n0 + i < N) D[(size_t)row * N + n0 + i] = __uint_as_float(r[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 256);
}

}  // namespace eaz

using namespace eaz;

// Debug entry (not in include/eaz_b200.h): A [128,K], W [K,N], D [128,N]; K % 32 == 0, N <= 256;
// wtiles: device scratch of K * 2 * roundup16(N) * 4 bytes.
extern "C" int eaz_debug_umma_gemm(const float* A, const float* W, float* D, int32_t K, int32_t N, void* wtiles, void* stream) {
  EAZ_CHECK_ARG(A && W && D && wtiles && K > 0 && K % umma::kChunkK == 0 && N >= 1 && N <= 256, "eaz_debug_umma_gemm: bad arguments");
  const int Npad = (N + 15) / 16 * 16;
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = launch_tile_weights(W, K, N, K, Npad, (uint32_t*)wtiles, st, 32)) return rc;
  const size_t smem = 2 * (2 * 128 * umma::kChunkK * 4) + 2 * (2 * Npad * umma::kChunkK * 4) + sizeof(GemmSmem) + 1024;
  cudaError_t e = cudaFuncSetAttribute(umma_gemm_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(umma_gemm_test_kernel)");
  umma_gemm_test_kernel<<<1, 160, smem, st>>>(A, (const uint32_t*)wtiles, D, K, N, Npad);
  EAZ_CHECK_LAUNCH("umma_gemm_test_kernel");
  return 0;
}

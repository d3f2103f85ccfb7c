// Micro-measurement: does griddepcontrol.wait (ACQBULK) invalidate L1 lines fetched before it?  Does a global store
// invalidate or update the L1 line?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pdl_l1 pdl_l1.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void primary(unsigned* sink, int iters, unsigned* data) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  unsigned x = threadIdx.x;
  for (int i = 0; i < iters; ++i) x = x * 1664525u + 1013904223u;
  if (x == 12345u) sink[0] = x;
  if (blockIdx.x < 64 && threadIdx.x < 32) data[1024 * blockIdx.x + threadIdx.x] = 777u + (x & 0);  // written late, after the secondary's pre-wait read
}
__device__ __forceinline__ unsigned ld_ca(const unsigned* p) {
  unsigned v;
  asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ long long clk() {
  long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
  return t;
}
// load + an ordered consumer of the result, so that the following clock read waits for the data
__device__ __forceinline__ unsigned ld_use(const unsigned* p) {
  __shared__ volatile unsigned sbuf[32];
  unsigned v;
  asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  sbuf[threadIdx.x] = v;  // the store needs the data
  __syncwarp();
  return sbuf[threadIdx.x ^ 1] * 0 + v;
}
__global__ void secondary(unsigned* data, long long* out, int do_wait) {
  const unsigned* p = data + 1024 * blockIdx.x + threadIdx.x;  // one 128-byte line per warp
  long long t0 = clk();
  unsigned a = ld_use(p);
  long long t1 = clk();   // first touch: L2 (or DRAM)
  unsigned b = ld_use(p + (a & 0));
  long long t2 = clk();   // second touch: L1 hit
  if (do_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
  long long t3 = clk();
  unsigned c = ld_use(p + (b & 0));
  long long t4 = clk();   // after the PDL wait
  data[1024 * blockIdx.x + threadIdx.x] = c + 1;   // store to the line
  __syncwarp();
  long long t5 = clk();
  unsigned d = ld_use(p + (c & 0));
  long long t6 = clk();   // after own store
  unsigned e = ld_use(p + 32 + (d & 0));  // neighbouring line never touched: L2 again
  long long t7 = clk();
  if (threadIdx.x == 0) {
    long long* o = out + 8 * blockIdx.x;
    o[0] = t1 - t0; o[1] = t2 - t1; o[2] = t4 - t3; o[3] = t6 - t5; o[4] = t7 - t6; o[5] = t3 - t2; o[6] = a; o[7] = c;
  }
}
int main() {
  unsigned *data, *sink;
  long long* out;
  cudaMalloc(&data, 1 << 22);
  cudaMemset(data, 0, 1 << 22);
  cudaMalloc(&sink, 64);
  cudaMalloc(&out, 8 * 64 * sizeof(long long));
  for (int mode = 0; mode < 3; ++mode) {  // 0: no PDL, no wait; 1: PDL launch + wait; 2: plain launch + wait instruction
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(data, 0, 1 << 22);
      primary<<<148, 128>>>(sink, 200000, data);
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(64); cfg.blockDim = dim3(32);
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr; cfg.numAttrs = mode == 1 ? 1 : 0;
      cudaLaunchKernelEx(&cfg, secondary, data, out, mode != 0 ? 1 : 0);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    }
    long long h[8 * 64];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double m[6] = {0, 0, 0, 0, 0, 0};
    for (int b = 0; b < 64; ++b) for (int k = 0; k < 6; ++k) m[k] += h[8 * b + k] / 64.0;
    printf("values: pre-wait %lld, post-wait %lld\n", h[6], h[7]);
    printf("mode %d (%s): first touch %.0f, L1 re-read %.0f, after wait %.0f, after own store %.0f, untouched line %.0f cycles; wait itself %.0f\n", mode,
           mode == 0 ? "plain, no wait" : mode == 1 ? "PDL launch + griddepcontrol.wait" : "plain launch + wait instr", m[0], m[1], m[2], m[3], m[4], m[5]);
  }
  return 0;
}

// Tensor memory as a lane-private scratchpad: throughput/latency of tcgen05.ld/st.32x32b against ld/st.shared for the
// access pattern of the simplex row update (read 16 B per lane, two fp64 ops per cell, write 16 B back).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tmem_stream tmem_stream.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_ld4(unsigned taddr, unsigned &a, unsigned &b, unsigned &c, unsigned &d) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(taddr));
}
__device__ __forceinline__ void tmem_st4(unsigned taddr, unsigned a, unsigned b, unsigned c, unsigned d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// rows: fp64 pairs per lane held in TMEM (4 columns each); COLS = allocation (power of two >= 32)
template <int COLS>
__global__ void k_tmem(int rows, int iters, double coef, double *sink, long long *cyc) {
  __shared__ unsigned base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"l"((unsigned long long)__cvta_generic_to_shared(&base_s)), "n"(COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const unsigned base = base_s + ((unsigned)(warp & 3) * 32u << 16);
  for (int r = 0; r < rows; r++) tmem_st4(base + 4 * r, 0u, 0x3ff00000u, 0u, 0x40000000u);  // (1.0, 2.0)
  tmem_wait_st();
  const double p0 = 1e-9 * (threadIdx.x + 1), p1 = 2e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    for (int r = 0; r < rows; r += 4) {
      unsigned v[4][4];
#pragma unroll
      for (int i = 0; i < 4; i++) tmem_ld4(base + 4 * (r + i), v[i][0], v[i][1], v[i][2], v[i][3]);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 4; i++) {
        double x0 = __hiloint2double(v[i][1], v[i][0]), x1 = __hiloint2double(v[i][3], v[i][2]);
        x0 = __dsub_rn(x0, __dmul_rn(coef, p0));
        x1 = __dsub_rn(x1, __dmul_rn(coef, p1));
        tmem_st4(base + 4 * (r + i), __double2loint(x0), __double2hiint(x0), __double2loint(x1), __double2hiint(x1));
      }
    }
  }
  tmem_wait_st();
  long long t1 = clock64();
  unsigned a, b, c, d;
  tmem_ld4(base, a, b, c, d);
  tmem_wait_ld();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = __hiloint2double(b, a);
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base_s), "n"(COLS));
}

__global__ void k_smem(int rows, int iters, double coef, double *sink, long long *cyc) {
  extern __shared__ __align__(16) unsigned char raw[];
  double2 *buf = reinterpret_cast<double2 *>(raw) + (threadIdx.x >> 5) * rows * 32 + (threadIdx.x & 31);
  for (int r = 0; r < rows; r++) buf[r * 32] = make_double2(1.0, 2.0);
  const double p0 = 1e-9 * (threadIdx.x + 1), p1 = 2e-9;
  __syncwarp();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    for (int r = 0; r < rows; r += 4) {
      double2 v[4];
#pragma unroll
      for (int i = 0; i < 4; i++) v[i] = buf[(r + i) * 32];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        v[i].x = __dsub_rn(v[i].x, __dmul_rn(coef, p0));
        v[i].y = __dsub_rn(v[i].y, __dmul_rn(coef, p1));
        buf[(r + i) * 32] = v[i];
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = buf[0].x;
}

int main() {
  double *sink; long long *cyc;
  cudaMalloc(&sink, 148 * 32 * 1024 * 8); cudaMallocManaged(&cyc, 64);
  const int iters = 2000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto report = [&](const char *what, int ctas_per_sm, int warps, int rows, float ms) {
    const double bytes = 148.0 * ctas_per_sm * warps * 32 * rows * 32.0 * iters;  // 16 B read + 16 B written per lane-row
    printf("%-34s ctas/SM=%2d warps/CTA=%d rows=%2d: %8.3f ms  %7.1f GB/s/SM (r+w)  %6.1f cycles per 4-row block (thread 0)\n", what,
           ctas_per_sm, warps, rows, ms, bytes / (ms * 1e-3) / 1e9 / 148.0, (double)cyc[0] / iters / (rows / 4));
  };
  for (int rep = 0; rep < 2; rep++) {
    // TMEM, one-warp CTAs with 32-column allocations (8 rows of 2 fp64 per lane), 16 CTAs per SM
    cudaEventRecord(e0); k_tmem<32><<<148 * 16, 32>>>(8, iters, 0.5, sink, cyc); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) report("tmem 32x32b.x4, 1-warp CTAs", 16, 1, 8, ms);
    printf(rep ? "" : "%s", cudaGetErrorString(cudaGetLastError())); if (!rep) printf("\n");
    // TMEM, four-warp CTAs with 128-column allocations (32 rows per lane), 4 CTAs per SM
    cudaEventRecord(e0); k_tmem<128><<<148 * 4, 128>>>(32, iters, 0.5, sink, cyc); cudaEventRecord(e1); cudaDeviceSynchronize();
    cudaEventElapsedTime(&ms, e0, e1); if (rep) report("tmem 32x32b.x4, 4-warp CTAs", 4, 4, 32, ms);
    // TMEM, one 4-warp CTA per SM (latency view)
    cudaEventRecord(e0); k_tmem<128><<<148, 128>>>(32, iters, 0.5, sink, cyc); cudaEventRecord(e1); cudaDeviceSynchronize();
    cudaEventElapsedTime(&ms, e0, e1); if (rep) report("tmem 32x32b.x4, 1 CTA/SM", 1, 4, 32, ms);
    cudaEventRecord(e0); k_tmem<32><<<148, 32>>>(8, iters, 0.5, sink, cyc); cudaEventRecord(e1); cudaDeviceSynchronize();
    cudaEventElapsedTime(&ms, e0, e1); if (rep) report("tmem 32x32b.x4, 1 warp/SM", 1, 1, 8, ms);
    // shared memory, same pattern
    cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaEventRecord(e0); k_smem<<<148 * 16, 32, 8 * 32 * 16>>>(8, iters, 0.5, sink, cyc); cudaEventRecord(e1); cudaDeviceSynchronize();
    cudaEventElapsedTime(&ms, e0, e1); if (rep) report("smem ld/st.v2.f64, 1-warp CTAs", 16, 1, 8, ms);
    cudaEventRecord(e0); k_smem<<<148 * 12, 32, 32 * 32 * 16>>>(32, iters, 0.5, sink, cyc); cudaEventRecord(e1); cudaDeviceSynchronize();
    cudaEventElapsedTime(&ms, e0, e1); if (rep) report("smem ld/st.v2.f64, 12 1-warp CTAs", 12, 1, 32, ms);
    cudaEventRecord(e0); k_smem<<<148, 32, 8 * 32 * 16>>>(8, iters, 0.5, sink, cyc); cudaEventRecord(e1); cudaDeviceSynchronize();
    cudaEventElapsedTime(&ms, e0, e1); if (rep) report("smem ld/st.v2.f64, 1 warp/SM", 1, 1, 8, ms);
  }
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

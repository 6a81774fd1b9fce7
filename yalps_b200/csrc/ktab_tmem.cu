// K1t: tensor-memory resident kernel (tmem_kernel.cuh), its own translation unit.
#include "tmem_kernel.cuh"
#include "tmem_launch.h"

namespace yalps {

bool tmem_kernel_fits(int Hcap, int Wcap) { return Hcap >= 1 && Wcap >= 1 && Hcap <= kTmemMaxRows && Wcap <= kTmemMaxCols; }
int tmem_kernel_warps() { return kTmemWarps; }
// shape by tableau height: HR = 1 (up to 33 rows, 16 LPs per SM) or HR = 2 (up to 65 rows, 8 LPs per SM)
static bool tall(int Hcap) { return Hcap > TmemShape<1>::kMaxRows; }
bool tmem_kernel_is_tall(int Hcap) { return tall(Hcap); }
int tmem_kernel_ctas_per_sm(int Hcap) { return tall(Hcap) ? TmemShape<2>::kCtasPerSm : TmemShape<1>::kCtasPerSm; }

// A CTA beyond 512 / kColumns per SM would block in tcgen05.alloc until a resident (persistent) CTA exits:
// a dynamic shared-memory request that only fits kCtasPerSm times keeps such CTAs from becoming resident.
size_t tmem_kernel_dynamic_smem(int Hcap) { return (size_t)(227 * 1024) / (tmem_kernel_ctas_per_sm(Hcap) + 1) + 1024; }
size_t tmem_stream_dynamic_smem() { return tmem_kernel_dynamic_smem(1); }

const void *tmem_kernel_fn(int Hcap, bool count_rows, bool cycles) {
  if (cycles) {
    if (count_rows) return tall(Hcap) ? (const void *)k_simplex_tmem<2, true, true> : (const void *)k_simplex_tmem<1, true, true>;
    return tall(Hcap) ? (const void *)k_simplex_tmem<2, false, true> : (const void *)k_simplex_tmem<1, false, true>;
  }
  if (count_rows) return tall(Hcap) ? (const void *)k_simplex_tmem<2, true> : (const void *)k_simplex_tmem<1, true>;
  return tall(Hcap) ? (const void *)k_simplex_tmem<2> : (const void *)k_simplex_tmem<1>;
}

// Tensor-memory stream: every lane reads and rewrites its eight-row blocks (tcgen05.ld/st.32x32b.x32) with the
// multiply-subtract of the rank-1 update in between.  bytes = grid * 128 lanes * iters * 128 columns * 4 B * 2.
__global__ void __launch_bounds__(kTmemWarps * 32, kTmemCtasPerSm) k_tmem_stream(int iters, double coef, double *sink) {
  __shared__ unsigned s_base;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&s_base)),
                 "n"(kTmemColumns)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tbase = s_base + ((unsigned)(warp & 3) * 32u << 16);
  unsigned v[32];
#pragma unroll
  for (int i = 0; i < 32; i++) v[i] = (i & 1) ? 0x3ff00000u : 0u;  // 1.0
  for (int blk = 0; blk < kTmemColumns / 32; blk++) tm_st32(tbase + 32u * blk, v);
  tm_wait_st();
  const double p0 = 1e-9 * (threadIdx.x + 1), p1 = 2e-9;
  double acc = 0.0;
  for (int it = 0; it < iters; it++) {
    for (int blk = 0; blk < kTmemColumns / 32; blk++) {
      tm_ld32(tbase + 32u * blk, v);
      tm_wait_ld(v);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        double x0 = __hiloint2double((int)v[2 * i + 1], (int)v[2 * i]), x1 = __hiloint2double((int)v[17 + 2 * i], (int)v[16 + 2 * i]);
        x0 = __dsub_rn(x0, __dmul_rn(coef, p0));
        x1 = __dsub_rn(x1, __dmul_rn(coef, p1));
        v[2 * i] = (unsigned)__double2loint(x0);
        v[2 * i + 1] = (unsigned)__double2hiint(x0);
        v[16 + 2 * i] = (unsigned)__double2loint(x1);
        v[17 + 2 * i] = (unsigned)__double2hiint(x1);
      }
      tm_st32(tbase + 32u * blk, v);
    }
    tm_wait_st();
  }
  tm_ld32(tbase, v);
  tm_wait_ld(v);
  acc = __hiloint2double((int)v[1], (int)v[0]);
  if (acc == -1.0) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_base), "n"(kTmemColumns) : "memory");
}

const void *tmem_stream_fn() { return (const void *)k_tmem_stream; }
int tmem_stream_ctas_per_sm() { return kTmemCtasPerSm; }

cudaError_t launch_tmem_stream(int grid, int iters, double *sink, cudaStream_t stream, double *bytes) {
  k_tmem_stream<<<grid, kTmemWarps * 32, tmem_stream_dynamic_smem(), stream>>>(iters, 0.5, sink);
  *bytes = (double)grid * kTmemWarps * 32 * iters * kTmemColumns * 4.0 * 2.0;
  return cudaGetLastError();
}

cudaError_t launch_simplex_tmem(const BatchArgs &args, int grid, cudaStream_t stream) {
  const size_t smem = tmem_kernel_dynamic_smem(args.Hcap);
  if (args.check_cycles) {  // the instantiations with the cycle history
    const bool t = tall(args.Hcap);
    if (args.rows_out) {
      if (t)
        k_simplex_tmem<2, true, true><<<grid, kTmemWarps * 32, smem, stream>>>(args);
      else
        k_simplex_tmem<1, true, true><<<grid, kTmemWarps * 32, smem, stream>>>(args);
    } else if (t) {
      k_simplex_tmem<2, false, true><<<grid, kTmemWarps * 32, smem, stream>>>(args);
    } else {
      k_simplex_tmem<1, false, true><<<grid, kTmemWarps * 32, smem, stream>>>(args);
    }
  } else if (args.rows_out) {  // diagnostics instantiation: same code plus the rewritten-row counter
    if (tall(args.Hcap))
      k_simplex_tmem<2, true><<<grid, kTmemWarps * 32, smem, stream>>>(args);
    else
      k_simplex_tmem<1, true><<<grid, kTmemWarps * 32, smem, stream>>>(args);
  } else if (tall(args.Hcap)) {
    k_simplex_tmem<2><<<grid, kTmemWarps * 32, smem, stream>>>(args);
  } else {
    k_simplex_tmem<1><<<grid, kTmemWarps * 32, smem, stream>>>(args);
  }
  return cudaGetLastError();
}

}  // namespace yalps

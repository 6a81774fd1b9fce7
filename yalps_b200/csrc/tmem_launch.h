// tmem_launch.h -- host-side interface of the tensor-memory kernel K1t (tmem_kernel.cuh, compiled in ktab_tmem.cu).
#pragma once

#include <cuda_runtime.h>

#include "kernels.cuh"

namespace yalps {

bool tmem_kernel_fits(int Hcap, int Wcap);
int tmem_kernel_warps();         // LPs in flight per CTA (one per warp)
bool tmem_kernel_is_tall(int Hcap);       // more than 33 rows: the 256-column shape (8 LPs per SM instead of 16)
int tmem_kernel_ctas_per_sm(int Hcap);    // 512 TMEM columns / allocation per CTA
size_t tmem_kernel_dynamic_smem(int Hcap);  // padding request that keeps residency at tmem_kernel_ctas_per_sm()
const void *tmem_kernel_fn(int Hcap, bool count_rows = false, bool cycles = false);  // for cudaFuncSetAttribute (dynamic shared-memory limit, per device)
const void *tmem_stream_fn();
int tmem_stream_ctas_per_sm();
size_t tmem_stream_dynamic_smem();
cudaError_t launch_tmem_stream(int grid, int iters, double *sink, cudaStream_t stream, double *bytes);  // k_tmem_stream
cudaError_t launch_simplex_tmem(const BatchArgs &args, int grid, cudaStream_t stream);

}  // namespace yalps

// tmem_launch.h -- host-side interface of the tensor-memory kernel K1t (tmem_kernel.cuh, compiled in ktab_tmem.cu).
#pragma once

#include <cuda_runtime.h>

#include "kernels.cuh"

namespace yalps {

bool tmem_kernel_fits(int Hcap, int Wcap);
int tmem_kernel_warps();         // LPs in flight per CTA (one per warp)
int tmem_kernel_ctas_per_sm();   // 512 TMEM columns / allocation per CTA
size_t tmem_kernel_dynamic_smem();  // padding request that keeps residency at tmem_kernel_ctas_per_sm()
const void *tmem_kernel_fn();        // for cudaFuncSetAttribute (dynamic shared-memory limit, per device)
const void *tmem_stream_fn();
cudaError_t launch_tmem_stream(int grid, int iters, double *sink, cudaStream_t stream, double *bytes);  // k_tmem_stream
cudaError_t launch_simplex_tmem(const BatchArgs &args, int grid, cudaStream_t stream);

}  // namespace yalps

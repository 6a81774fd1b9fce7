// K1t: tensor-memory resident kernel (tmem_kernel.cuh), its own translation unit.
#include "tmem_kernel.cuh"
#include "tmem_launch.h"

namespace yalps {

bool tmem_kernel_fits(int Hcap, int Wcap) { return Hcap >= 1 && Wcap >= 1 && Hcap <= kTmemMaxRows && Wcap <= kTmemMaxCols; }
int tmem_kernel_warps() { return kTmemWarps; }
int tmem_kernel_ctas_per_sm() { return kTmemCtasPerSm; }

// A CTA beyond 512 / kTmemColumns per SM would block in tcgen05.alloc until a resident (persistent) CTA exits:
// a dynamic shared-memory request that only fits kTmemCtasPerSm times keeps such CTAs from becoming resident.
size_t tmem_kernel_dynamic_smem() { return (size_t)(227 * 1024) / (kTmemCtasPerSm + 1) + 1024; }

const void *tmem_kernel_fn() { return (const void *)k_simplex_tmem; }

cudaError_t launch_simplex_tmem(const BatchArgs &args, int grid, cudaStream_t stream) {
  k_simplex_tmem<<<grid, kTmemWarps * 32, tmem_kernel_dynamic_smem(), stream>>>(args);
  return cudaGetLastError();
}

}  // namespace yalps

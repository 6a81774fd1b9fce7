// bnb_launch.h -- host-side interface of the device-resident branch-and-cut kernel KB (bnb_kernel.cuh, ktab_bnb.cu).
#pragma once

#include <cuda_runtime.h>

#include "bnb_kernel.cuh"

namespace yalps {

struct BnbConfig {
  int nwc, kc, nwr;     // CTA shape of the workers (row-split simplex: nwc column warps x nwr row groups)
  const void *fn;
};

// Narrowest instantiated worker shape whose column split covers a tableau of width W (nullptr: none).
const BnbConfig *bnb_config_for(int W);
cudaError_t launch_bnb(const BnbConfig *cfg, const BnbArgs &args, int grid, size_t smem, cudaStream_t stream);

}  // namespace yalps

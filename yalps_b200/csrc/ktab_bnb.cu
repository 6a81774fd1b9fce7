// KB: device-resident branch and cut (bnb_kernel.cuh), its own translation unit.
#include "bnb_launch.h"

namespace yalps {

#define BENTRY(NWC, KC, NWR) {NWC, KC, NWR, (const void *)k_bnb<NWC, KC, NWR>}
static const BnbConfig kConfigs[] = {
    BENTRY(2, 1, 4),  // 256 threads, W <= 129
    BENTRY(4, 1, 4),  // 512 threads, W <= 257
    BENTRY(4, 2, 4),  // 512 threads, W <= 513
};
#undef BENTRY

const BnbConfig *bnb_config_for(int W) {
  for (const BnbConfig &c : kConfigs)
    if (32 * c.nwc * c.kc * 2 >= W - 1) return &c;
  return nullptr;
}

cudaError_t launch_bnb(const BnbConfig *cfg, const BnbArgs &args, int grid, size_t smem, cudaStream_t stream) {
  BnbArgs a = args;
  void *params[] = {&a};
  return cudaLaunchCooperativeKernel(cfg->fn, dim3(grid), dim3(cfg->nwc * cfg->nwr * 32), params, smem, stream);
}

}  // namespace yalps

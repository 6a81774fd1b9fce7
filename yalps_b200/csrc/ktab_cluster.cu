// KC: cluster-resident kernels (cluster_kernel.cuh), 256 threads per CTA (no register spills: the cluster barrier
// invalidates L1, so every spill reload after it would be an L2 round trip; measured for KG as well: 512-thread
// instantiations with 128 registers spill ~450 bytes and run 25-65 % slower).
#include "cluster_kernel.cuh"
#include "kernel_table.h"

namespace yalps {
// `resident` = KC (one LP per cluster), `global` = KG (the same kernel with the whole cooperative grid as its cluster)
#define KENTRY(NW, KC, NWR) {NW, KC, NWR, k_simplex_cluster<NW, KC, NWR, false>, k_simplex_cluster<NW, KC, NWR, true>}
static const KernelEntry kTable[] = {
    KENTRY(1, 1, 8), KENTRY(1, 2, 8), KENTRY(2, 2, 4), KENTRY(4, 2, 2), KENTRY(8, 2, 1), KENTRY(8, 4, 1),
};
#undef KENTRY
const KernelEntry *kernel_table_cluster(int *count) {
  *count = (int)(sizeof(kTable) / sizeof(kTable[0]));
  return kTable;
}
}  // namespace yalps

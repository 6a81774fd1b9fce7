// Row-split kernels, two column warps.
#include "kernel_table.h"

namespace yalps {
#define KENTRY(NW, KC, NWR) {NW, KC, NWR, k_simplex<NW, KC, true, NWR>, k_simplex<NW, KC, false, NWR>}
static const KernelEntry kTable[] = {
    KENTRY(2, 1, 2), KENTRY(2, 1, 4), KENTRY(2, 1, 8),
    KENTRY(2, 2, 2), KENTRY(2, 2, 4), KENTRY(2, 2, 8),
    KENTRY(2, 4, 2), KENTRY(2, 4, 4), KENTRY(2, 4, 8),
};
#undef KENTRY
const KernelEntry *kernel_table_split_b(int *count) {
  *count = (int)(sizeof(kTable) / sizeof(kTable[0]));
  return kTable;
}
}  // namespace yalps

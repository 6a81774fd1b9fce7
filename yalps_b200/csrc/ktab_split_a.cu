// Row-split kernels, one column warp (tableaus up to 64*KC coefficient columns in shared memory).
#include "kernel_table.h"

namespace yalps {
#define KENTRY(NW, KC, NWR) {NW, KC, NWR, k_simplex<NW, KC, true, NWR>, k_simplex<NW, KC, false, NWR>}
static const KernelEntry kTable[] = {
    KENTRY(1, 1, 2), KENTRY(1, 1, 4), KENTRY(1, 1, 8), KENTRY(1, 1, 16),
    KENTRY(1, 2, 2), KENTRY(1, 2, 4), KENTRY(1, 2, 8), KENTRY(1, 2, 16),
    KENTRY(1, 4, 2), KENTRY(1, 4, 4), KENTRY(1, 4, 8), KENTRY(1, 4, 16),
};
#undef KENTRY
const KernelEntry *kernel_table_split_a(int *count) {
  *count = (int)(sizeof(kTable) / sizeof(kTable[0]));
  return kTable;
}
}  // namespace yalps

// Row-split kernels, four and eight column warps (wide tableaus).
#include "kernel_table.h"

namespace yalps {
#define KENTRY(NW, KC, NWR) {NW, KC, NWR, k_simplex<NW, KC, true, NWR>, k_simplex<NW, KC, false, NWR>}
static const KernelEntry kTable[] = {
    KENTRY(4, 1, 2), KENTRY(4, 1, 4), KENTRY(4, 2, 2), KENTRY(4, 2, 4), KENTRY(4, 4, 2), KENTRY(4, 4, 4),
    KENTRY(8, 2, 2), KENTRY(8, 4, 2), KENTRY(8, 8, 2),
};
#undef KENTRY
const KernelEntry *kernel_table_split_c(int *count) {
  *count = (int)(sizeof(kTable) / sizeof(kTable[0]));
  return kTable;
}
}  // namespace yalps

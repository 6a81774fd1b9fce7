// One row group per CTA (throughput kernels): thread t owns vector-columns t, t+NT, ... and walks all rows.
#include "kernel_table.h"

namespace yalps {
#define KENTRY(NW, KC) {NW, KC, 1, k_simplex<NW, KC, true, 1>, k_simplex<NW, KC, false, 1>}
static const KernelEntry kTable[] = {
    KENTRY(1, 1), KENTRY(1, 2),  KENTRY(1, 3),  KENTRY(1, 4),  KENTRY(2, 1),  KENTRY(2, 2),  KENTRY(2, 3),
    KENTRY(2, 4), KENTRY(4, 1),  KENTRY(4, 2),  KENTRY(4, 4),  KENTRY(8, 1),  KENTRY(8, 2),  KENTRY(8, 4),
    KENTRY(8, 8), KENTRY(16, 1), KENTRY(16, 2), KENTRY(16, 4), KENTRY(32, 1), KENTRY(32, 2), KENTRY(32, 4),
    KENTRY(32, 8),
};
#undef KENTRY
const KernelEntry *kernel_table_base(int *count) {
  *count = (int)(sizeof(kTable) / sizeof(kTable[0]));
  return kTable;
}
}  // namespace yalps

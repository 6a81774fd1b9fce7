// kernel_table.h -- the instantiated k_simplex variants, spread over several translation units so that the
// library builds in parallel (ktab_base.cu: one row group; ktab_split_*.cu: row-split latency kernels).
#pragma once

#include "kernels.cuh"

namespace yalps {

typedef void (*SimplexKernel)(const BatchArgs);

struct KernelEntry {
  int nw;   // column warps (NWC)
  int kc;   // vector-columns per thread
  int nwr;  // row groups; CTA = nw * nwr warps
  SimplexKernel resident, global;
};

const KernelEntry *kernel_table_base(int *count);
const KernelEntry *kernel_table_split_a(int *count);
const KernelEntry *kernel_table_split_b(int *count);
const KernelEntry *kernel_table_split_c(int *count);
const KernelEntry *kernel_table_cluster(int *count);  // cluster_kernel.cuh: `resident` = KC, `global` = KG (grid-resident)

}  // namespace yalps

// fwd_extra.cu — histogram-selected forward families: eight gathers in flight for long-row graphs
// (fp32 rows, one register tile) and the row-parallel layout for short rows x narrow dense rows.
#include "fwd_launch.cuh"

namespace ofspmm {
int launch_family_unroll8(const FwdParams& p, int idx_dtype, int dense_dtype, int val_dtype, bool aligned,
                          const FwdLaunch& L, cudaStream_t stream) {
  return launch_family<false, kTaskItems, 2>(p, idx_dtype, dense_dtype, val_dtype, aligned, L, stream);
}
int launch_family_rowpar(const FwdParams& p, int idx_dtype, int dense_dtype, int val_dtype, bool aligned,
                         const FwdLaunch& L, cudaStream_t stream) {
  return launch_family<true, kTaskItems, 1>(p, idx_dtype, dense_dtype, val_dtype, aligned, L, stream);
}
}  // namespace ofspmm

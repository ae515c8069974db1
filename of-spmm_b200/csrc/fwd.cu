// fwd.cu — variant resolution and dispatch of the merge-path SpMM forward to its kernel family
// (fwd_launch.cuh), plus the task partitioner launch.
#include "internal.h"
#include "spmm_kernels.cuh"

namespace ofspmm {

int launch_family_base(const FwdParams&, int, int, int, bool, const FwdLaunch&, cudaStream_t);
int launch_family_small(const FwdParams&, int, int, int, bool, const FwdLaunch&, cudaStream_t);
int launch_family_rowpar(const FwdParams&, int, int, int, bool, const FwdLaunch&, cudaStream_t);
int launch_family_rows(const FwdParams&, int, int, cudaStream_t);

namespace {

int64_t vec_width(int dense_dtype) { return dense_dtype == OFSPMM_DTYPE_BFLOAT16 ? 8 : 4; }

// Which dense widths the row-parallel family has kernels for (16-byte aligned rows assumed; fwd
// falls back to the base family otherwise): the sub-warp layouts, 8 or 16 lanes per dense row.
bool rowpar_supported(int64_t n, int dense_dtype) {
  const int64_t v = vec_width(dense_dtype);
  return n % v == 0 && n / v <= 16;
}

// Dense widths the whole-row family has kernels for: 16-byte vectors, at most 32 of them.
bool rows_supported(int64_t n, int dense_dtype) {
  const int64_t v = vec_width(dense_dtype);
  return n > 0 && n % v == 0 && n / v <= 32;
}

}  // namespace

// AUTO (no histogram available): decided from the host-known sizes only — fewer 256-item tasks
// than resident warps -> 64-item tasks.  With the row-length histogram on the host,
// ofspmm_choose_variant() may also pick the row-parallel layout (api.cu).
FwdVariant resolve_variant(int variant, int64_t rows, int64_t nnz, int64_t n, int dense_dtype) {
  FwdVariant v{kTaskItems, false, false};
  if (variant & OFSPMM_VARIANT_EXPLICIT) {
    if (variant & OFSPMM_VARIANT_ROWS) {
      // sized like the 64-item family, which is also what runs when a launch cannot use the
      // whole-row kernel (int64 indices, unaligned or strided-odd rows)
      v.items = kSmallTaskItems;
      v.whole_rows = rows_supported(n, dense_dtype);
    } else if (variant & OFSPMM_VARIANT_ITEMS64) v.items = kSmallTaskItems;
    else if ((variant & OFSPMM_VARIANT_ROWPAR) && rowpar_supported(n, dense_dtype)) v.row_parallel = true;
    return v;
  }
  if (num_tasks(rows, nnz, kTaskItems) < kSmallProblemTasks) v.items = kSmallTaskItems;
  return v;
}

int encode_variant(const FwdVariant& v) {
  int code = OFSPMM_VARIANT_EXPLICIT;
  if (v.whole_rows) code |= OFSPMM_VARIANT_ROWS;
  else if (v.items == kSmallTaskItems) code |= OFSPMM_VARIANT_ITEMS64;
  if (v.row_parallel) code |= OFSPMM_VARIANT_ROWPAR;
  return code;
}

int launch_task_partition(const void* crow, int idx_dtype, int64_t rows, int64_t nnz, int64_t P,
                          int items, void* part, cudaStream_t stream) {
  const int threads = 256;
  const unsigned blocks = static_cast<unsigned>((P + 1 + threads - 1) / threads);
  if (idx_dtype == OFSPMM_DTYPE_INT32) {
    task_partition_kernel<int32_t><<<blocks, threads, 0, stream>>>(
        static_cast<const int32_t*>(crow), static_cast<int>(rows), static_cast<int>(nnz), items,
        static_cast<int>(P), static_cast<int2*>(part));
  } else {
    task_partition_kernel<int64_t><<<blocks, threads, 0, stream>>>(
        static_cast<const int64_t*>(crow), static_cast<int>(rows), static_cast<int>(nnz), items,
        static_cast<int>(P), static_cast<int2*>(part));
  }
  count_launch();
  OFSPMM_CUDA_OK(cudaGetLastError());
  return OFSPMM_OK;
}

// The whole-row kernel needs what resolve_variant cannot see: the pointers' alignment, the
// strides and the index dtype of THIS call.
bool rows_kernel_applies(const ofspmm_csr* A, const void* B, int64_t ldb, const void* C, int64_t ldc, int64_t n,
                         int dense_dtype, const FwdLaunch& L) {
  if (!L.variant.whole_rows || A->idx_dtype != OFSPMM_DTYPE_INT32 || !rows_supported(n, dense_dtype)) return false;
  const int64_t v = vec_width(dense_dtype);
  const uintptr_t bits = reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C) |
                         ((L.flags & kFwdBias) ? reinterpret_cast<uintptr_t>(L.bias) : 0);
  return (bits & 15) == 0 && ldb % v == 0 && ldc % v == 0;
}

int launch_fwd(const ofspmm_csr* A, const void* B, int64_t ldb, void* C, int64_t ldc, int64_t n,
               int dense_dtype, const void* part, float* carry, float* head, void* counter, int64_t P,
               const FwdLaunch& L, cudaStream_t stream) {
  FwdParams p;
  p.ldb = ldb;
  p.ldc = ldc;
  p.crow = A->crow;
  p.col = A->col;
  p.val = A->val;
  p.B = B;
  p.C = C;
  p.part = static_cast<const int2*>(part);
  p.carry = carry;
  p.head = head;
  p.cols = A->cols;
  p.rows = static_cast<int>(A->rows);
  p.nnz = static_cast<int>(A->nnz);
  p.n = static_cast<int>(n);
  p.P = static_cast<int>(P);
  p.panels = 1;
  p.bias = L.bias;
  p.flags = L.flags;
  p.acc32 = L.acc32;
  p.counter = L.dynamic ? static_cast<unsigned int*>(counter) : nullptr;
  const int64_t v = vec_width(dense_dtype);
  const bool aligned = ((reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C) |
                         ((L.flags & kFwdBias) ? reinterpret_cast<uintptr_t>(L.bias) : 0)) & 15) == 0;
  const bool vec_rows = aligned && n % v == 0 && ldb % v == 0 && ldc % v == 0;
  if (rows_kernel_applies(A, B, ldb, C, ldc, n, dense_dtype, L)) {
    p.counter = nullptr;
    return launch_family_rows(p, dense_dtype, A->val_dtype, stream);
  }
  if (L.variant.items == kSmallTaskItems)
    return launch_family_small(p, A->idx_dtype, dense_dtype, A->val_dtype, aligned, L, stream);
  if (L.variant.row_parallel && vec_rows && rowpar_supported(n, dense_dtype))
    return launch_family_rowpar(p, A->idx_dtype, dense_dtype, A->val_dtype, aligned, L, stream);
  return launch_family_base(p, A->idx_dtype, dense_dtype, A->val_dtype, aligned, L, stream);
}

const char* fwd_variant_name(int64_t n, int dense_dtype, bool aligned) {
  const int vecw = dense_dtype == OFSPMM_DTYPE_BFLOAT16 ? 8 : 4;
  if (aligned && n % vecw == 0) {
    const int64_t nvec = n / vecw;
    if (nvec <= 8) return "vector-per-row(8 lanes x 16B, 4 nnz per step)";
    if (nvec <= 16) return "vector-per-row(16 lanes x 16B, 2 nnz per step)";
    if (nvec <= 32) return "warp-per-row(32 lanes x 16B)";
    if (nvec <= 64) return "warp-per-row(32 lanes x 2 x 16B)";
    return "warp-per-row(32 lanes x 4 x 16B, column panels)";
  }
  return "warp-per-row(scalar lanes, unaligned n)";
}

}  // namespace ofspmm

// fwd.cu — variant selection and launch of the merge-path SpMM kernels (spmm_kernels.cuh).
#include <stdlib.h>

#include "internal.h"
#include "spmm_kernels.cuh"

namespace ofspmm {

namespace {

template <typename DT, typename ValT, typename IdxT, int VEC, int LPR, int CH, bool kFull, bool kRowPar = false>
int launch_full(FwdParams p, int panels, cudaStream_t stream) {
  p.panels = panels;
  constexpr int ITEMS = kTaskItems;
  constexpr int WARPS = kWarpsPerCta;
  auto kern = spmm_merge_kernel<DT, ValT, IdxT, VEC, LPR, CH, kFull, kRowPar, ITEMS, WARPS>;
  const size_t smem = sizeof(TaskStage<IdxT, ValT, ITEMS>) * WARPS + sizeof(uint64_t) * WARPS;
  OFSPMM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  DevInfo dev;
  if (int rc = get_dev_info(&dev)) return rc;
  int occ = 0;
  OFSPMM_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem));
  if (occ < 1) return OFSPMM_ERR_CUDA;
  // persistent grid: a whole number of CTAs per SM (148 SMs on B200), never more than the tasks
  const int64_t ctas_needed = (static_cast<int64_t>(p.P) + WARPS - 1) / WARPS;
  const int64_t ctas_all = (static_cast<int64_t>(p.P) * panels + WARPS - 1) / WARPS;
  const int64_t resident = static_cast<int64_t>(dev.sms) * occ;
  int64_t gx64 = ctas_all < resident ? ctas_all : resident;
  // OFSPMM_TASKS_PER_WARP=k (env, tuning): non-persistent grid whose CTAs retire after ~k tasks
  // per warp, so kernels of a higher-priority stream (NCCL collectives of the next column panel)
  // can get SMs while this kernel is still running.  Default: persistent.
  if (const char* env = getenv("OFSPMM_TASKS_PER_WARP")) {
    const long k = strtol(env, nullptr, 10);
    if (k > 0) {
      const int64_t want = (ctas_all + k - 1) / k;
      if (want > gx64) gx64 = want;
    }
  }
  const int gx = static_cast<int>(gx64);
  kern<<<gx, WARPS * 32, smem, stream>>>(p);
  count_launch();
  OFSPMM_CUDA_OK(cudaGetLastError());
  // stitch rows that span several tasks
  auto fix = spmm_fixup_kernel<DT, IdxT, VEC, WARPS>;
  fix<<<static_cast<unsigned>((ctas_needed + 31) / 32), WARPS * 32, 0, stream>>>(p);  // one lane per task
  count_launch();
  OFSPMM_CUDA_OK(cudaGetLastError());
  return OFSPMM_OK;
}

// kFull (no masked lanes, immediate chunk offsets) when n is a whole number of register tiles.
template <typename DT, typename ValT, typename IdxT, int VEC, int LPR, int CH>
int launch_one(const FwdParams& p, int panels, cudaStream_t stream) {
  if (p.n % (LPR * VEC * CH) == 0) return launch_full<DT, ValT, IdxT, VEC, LPR, CH, true>(p, panels, stream);
  return launch_full<DT, ValT, IdxT, VEC, LPR, CH, false>(p, panels, stream);
}

template <typename DT, typename ValT, typename IdxT>
int launch_typed(const FwdParams& p, bool aligned, cudaStream_t stream) {
  constexpr int VECW = 16 / sizeof(DT);
  const int n = p.n;
  if (aligned && n % VECW == 0 && p.ldb % VECW == 0 && p.ldc % VECW == 0) {
    const int nvec = n / VECW;
    if (nvec <= 8) return launch_one<DT, ValT, IdxT, VECW, 8, 1>(p, 1, stream);
    if (nvec <= 16) return launch_one<DT, ValT, IdxT, VECW, 16, 1>(p, 1, stream);
#ifdef OFSPMM_FORCE_PANEL16
    // tuning experiment: half-width column panels (B panel = half the bytes in L2)
    if (nvec % 16 == 0) return launch_one<DT, ValT, IdxT, VECW, 16, 1>(p, nvec / 16, stream);
#endif
#ifdef OFSPMM_FORCE_ROWPAR
    // tuning experiment: row-parallel layout (8 lanes x 4 chunks per row, 4 rows in flight per warp)
    if (nvec == 32) return launch_full<DT, ValT, IdxT, VECW, 8, 4, true, true>(p, 1, stream);
#endif
    if (nvec <= 32) return launch_one<DT, ValT, IdxT, VECW, 32, 1>(p, 1, stream);
    if (nvec <= 64) return launch_one<DT, ValT, IdxT, VECW, 32, 2>(p, 1, stream);
    return launch_one<DT, ValT, IdxT, VECW, 32, 4>(p, (nvec + 127) / 128, stream);
  }
  if (n <= 32) return launch_one<DT, ValT, IdxT, 1, 32, 1>(p, 1, stream);
  if (n <= 64) return launch_one<DT, ValT, IdxT, 1, 32, 2>(p, 1, stream);
  if (n <= 128) return launch_one<DT, ValT, IdxT, 1, 32, 4>(p, 1, stream);
  return launch_one<DT, ValT, IdxT, 1, 32, 8>(p, (n + 255) / 256, stream);
}

template <typename IdxT>
int launch_idx(const FwdParams& p, int dense_dtype, int val_dtype, bool aligned, cudaStream_t stream) {
  if (dense_dtype == OFSPMM_DTYPE_FLOAT && val_dtype == OFSPMM_DTYPE_FLOAT)
    return launch_typed<float, float, IdxT>(p, aligned, stream);
  if (dense_dtype == OFSPMM_DTYPE_BFLOAT16 && val_dtype == OFSPMM_DTYPE_FLOAT)
    return launch_typed<__nv_bfloat16, float, IdxT>(p, aligned, stream);
  if (dense_dtype == OFSPMM_DTYPE_BFLOAT16 && val_dtype == OFSPMM_DTYPE_BFLOAT16)
    return launch_typed<__nv_bfloat16, __nv_bfloat16, IdxT>(p, aligned, stream);
  return OFSPMM_ERR_UNSUPPORTED_DTYPE;
}

}  // namespace

int launch_task_partition(const void* crow, int idx_dtype, int64_t rows, int64_t nnz, int64_t P,
                          void* part, cudaStream_t stream) {
  const int threads = 256;
  const unsigned blocks = static_cast<unsigned>((P + 1 + threads - 1) / threads);
  if (idx_dtype == OFSPMM_DTYPE_INT32) {
    task_partition_kernel<int32_t><<<blocks, threads, 0, stream>>>(
        static_cast<const int32_t*>(crow), static_cast<int>(rows), static_cast<int>(nnz), kTaskItems,
        static_cast<int>(P), static_cast<int2*>(part));
  } else {
    task_partition_kernel<int64_t><<<blocks, threads, 0, stream>>>(
        static_cast<const int64_t*>(crow), static_cast<int>(rows), static_cast<int>(nnz), kTaskItems,
        static_cast<int>(P), static_cast<int2*>(part));
  }
  count_launch();
  OFSPMM_CUDA_OK(cudaGetLastError());
  return OFSPMM_OK;
}

int launch_fwd(const ofspmm_csr* A, const void* B, int64_t ldb, void* C, int64_t ldc, int64_t n,
               int dense_dtype, const void* part, float* carry, float* head, int64_t P,
               cudaStream_t stream) {
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
  const bool aligned = ((reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C)) & 15) == 0;
  if (A->idx_dtype == OFSPMM_DTYPE_INT32) return launch_idx<int32_t>(p, dense_dtype, A->val_dtype, aligned, stream);
  if (A->idx_dtype == OFSPMM_DTYPE_INT64) return launch_idx<int64_t>(p, dense_dtype, A->val_dtype, aligned, stream);
  return OFSPMM_ERR_UNSUPPORTED_DTYPE;
}

const char* fwd_variant_name(int64_t n, int dense_dtype, bool aligned) {
  const int vecw = dense_dtype == OFSPMM_DTYPE_BFLOAT16 ? 8 : 4;
  if (aligned && n % vecw == 0) {
    const int64_t nvec = n / vecw;
    if (nvec <= 8) return "merge_path/vector-per-row(8 lanes x 16B, 4 nnz per step)";
    if (nvec <= 16) return "merge_path/vector-per-row(16 lanes x 16B, 2 nnz per step)";
    if (nvec <= 32) return "merge_path/warp-per-row(32 lanes x 16B)";
    if (nvec <= 64) return "merge_path/warp-per-row(32 lanes x 2 x 16B)";
    return "merge_path/warp-per-row(32 lanes x 4 x 16B, column panels)";
  }
  return "merge_path/warp-per-row(scalar lanes, unaligned n)";
}

}  // namespace ofspmm

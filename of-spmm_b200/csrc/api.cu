// api.cu — the extern "C" boundary declared in include/ofspmm.h: argument validation, workspace
// carving and dispatch to the per-op launchers.  Nothing here allocates device memory or
// synchronises; every entry point is re-entrant.
#include <string.h>

#include "common.cuh"
#include "internal.h"

namespace ofspmm {

std::atomic<uint64_t> g_launches{0};

int get_dev_info(DevInfo* out) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return OFSPMM_ERR_NO_DEVICE;
  int sms = 0, major = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return OFSPMM_ERR_CUDA;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return OFSPMM_ERR_CUDA;
  if (major != 10) return OFSPMM_ERR_NO_DEVICE;  // the library carries sm_100a code only
  out->sms = sms;
  out->cc_major = major;
  return OFSPMM_OK;
}

FwdWorkspace fwd_workspace_layout(int64_t rows, int64_t nnz, int64_t n, int dense_dtype) {
  FwdWorkspace L;
  const size_t P = static_cast<size_t>(num_tasks(rows, nnz));
  L.part_off = 0;
  L.carry_off = align_up((P + 1) * sizeof(int2), 256);
  const size_t rowbuf = align_up(P * static_cast<size_t>(n) * sizeof(float), 256);
  L.head_off = L.carry_off + rowbuf;
  L.total = L.head_off + (dense_dtype == OFSPMM_DTYPE_FLOAT ? 0 : rowbuf);
  if (L.total == 0) L.total = 256;
  return L;
}

namespace {

bool dense_ok(int d) { return d == OFSPMM_DTYPE_FLOAT || d == OFSPMM_DTYPE_BFLOAT16; }
bool idx_ok(int d) { return d == OFSPMM_DTYPE_INT32 || d == OFSPMM_DTYPE_INT64; }
size_t dense_size(int d) { return d == OFSPMM_DTYPE_FLOAT ? 4 : 2; }
size_t idx_size(int d) { return d == OFSPMM_DTYPE_INT64 ? 8 : 4; }

int check_csr(const ofspmm_csr* A, bool need_val) {
  if (A == nullptr) return OFSPMM_ERR_INVALID_ARG;
  if (A->rows < 0 || A->cols < 0 || A->nnz < 0) return OFSPMM_ERR_INVALID_ARG;
  if (!idx_ok(A->idx_dtype)) return OFSPMM_ERR_UNSUPPORTED_DTYPE;
  if (A->val_dtype != OFSPMM_DTYPE_FLOAT && A->val_dtype != OFSPMM_DTYPE_BFLOAT16) return OFSPMM_ERR_UNSUPPORTED_DTYPE;
  if (A->rows >= (int64_t{1} << 31) - 1 || A->nnz >= (int64_t{1} << 31) - 1) return OFSPMM_ERR_TOO_LARGE;
  if (A->crow == nullptr) return OFSPMM_ERR_INVALID_ARG;
  if (A->nnz > 0 && (A->col == nullptr || (need_val && A->val == nullptr))) return OFSPMM_ERR_INVALID_ARG;
  return OFSPMM_OK;
}

int check_dtypes(const ofspmm_csr* A, int dense_dtype) {
  if (!dense_ok(dense_dtype)) return OFSPMM_ERR_UNSUPPORTED_DTYPE;
  if (A->val_dtype == OFSPMM_DTYPE_BFLOAT16 && dense_dtype != OFSPMM_DTYPE_BFLOAT16) return OFSPMM_ERR_UNSUPPORTED_DTYPE;
  return OFSPMM_OK;
}

int check_ws(const void* ws, size_t have, size_t need) {
  if (need == 0) return OFSPMM_OK;
  if (ws == nullptr || have < need) return OFSPMM_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(ws) & 15) return OFSPMM_ERR_WORKSPACE;
  return OFSPMM_OK;
}

// C = A·B through the merge-path kernels; shared by ofspmm_fwd and route (1) of ofspmm_bwd_b.
int run_fwd(const ofspmm_csr* A, const void* B, int64_t ldb, void* C, int64_t ldc, int64_t n,
            int dense_dtype, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (A->rows == 0 || n == 0) return OFSPMM_OK;
  if (ldb < n || ldc < n || ldb >= (int64_t{1} << 30)) return OFSPMM_ERR_INVALID_ARG;
  if (A->nnz == 0 || A->cols == 0) {
    OFSPMM_CUDA_OK(cudaMemset2DAsync(C, static_cast<size_t>(ldc) * dense_size(dense_dtype), 0,
                                     static_cast<size_t>(n) * dense_size(dense_dtype),
                                     static_cast<size_t>(A->rows), stream));
    return OFSPMM_OK;
  }
  const FwdWorkspace L = fwd_workspace_layout(A->rows, A->nnz, n, dense_dtype);
  if (int rc = check_ws(ws, ws_bytes, L.total)) return rc;
  unsigned char* w = static_cast<unsigned char*>(ws);
  const int64_t P = num_tasks(A->rows, A->nnz);
  if (int rc = launch_task_partition(A->crow, A->idx_dtype, A->rows, A->nnz, P, w + L.part_off, stream)) return rc;
  return launch_fwd(A, B, ldb, C, ldc, n, dense_dtype, w + L.part_off, reinterpret_cast<float*>(w + L.carry_off),
                    reinterpret_cast<float*>(w + L.head_off), P, stream);
}

}  // namespace
}  // namespace ofspmm

using namespace ofspmm;

extern "C" {

size_t ofspmm_fwd_workspace_bytes(int64_t rows, int64_t cols, int64_t nnz, int64_t n, int dense_dtype) {
  (void)cols;
  if (rows < 0 || nnz < 0 || n < 0) return 0;
  return fwd_workspace_layout(rows, nnz, n, dense_dtype).total;
}

int ofspmm_fwd(const ofspmm_csr* A, const void* B, void* C, int64_t n, int dense_dtype, void* workspace,
               size_t workspace_bytes, ofspmm_stream_t stream) {
  if (int rc = check_csr(A, true)) return rc;
  if (int rc = check_dtypes(A, dense_dtype)) return rc;
  if (n < 0 || n >= (int64_t{1} << 31)) return OFSPMM_ERR_INVALID_ARG;
  if (A->rows > 0 && n > 0 && (C == nullptr || (A->nnz > 0 && A->cols > 0 && B == nullptr))) return OFSPMM_ERR_INVALID_ARG;
  return run_fwd(A, B, n, C, n, n, dense_dtype, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}

int ofspmm_fwd_strided(const ofspmm_csr* A, const void* B, int64_t ldb, void* C, int64_t ldc, int64_t n,
                       int dense_dtype, void* workspace, size_t workspace_bytes, ofspmm_stream_t stream) {
  if (int rc = check_csr(A, true)) return rc;
  if (int rc = check_dtypes(A, dense_dtype)) return rc;
  if (n < 0 || n >= (int64_t{1} << 31)) return OFSPMM_ERR_INVALID_ARG;
  if (A->rows > 0 && n > 0 && (C == nullptr || (A->nnz > 0 && A->cols > 0 && B == nullptr))) return OFSPMM_ERR_INVALID_ARG;
  return run_fwd(A, B, ldb, C, ldc, n, dense_dtype, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}

size_t ofspmm_bwd_b_workspace_bytes(int64_t rows, int64_t cols, int64_t nnz, int64_t n, int dense_dtype,
                                    int have_transpose) {
  if (rows < 0 || cols < 0 || nnz < 0 || n < 0) return 0;
  if (have_transpose) return fwd_workspace_layout(cols, nnz, n, dense_dtype).total;
  size_t total = align_up((static_cast<size_t>(num_tasks(rows, nnz)) + 1) * sizeof(int2), 256);
  if (dense_dtype != OFSPMM_DTYPE_FLOAT) total += align_up(static_cast<size_t>(cols) * n * sizeof(float), 256);
  return total;
}

int ofspmm_bwd_b(const ofspmm_csr* A, const ofspmm_csr* At, const void* dY, void* dB, int64_t n,
                 int dense_dtype, void* workspace, size_t workspace_bytes, ofspmm_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = check_csr(A, true)) return rc;
  if (int rc = check_dtypes(A, dense_dtype)) return rc;
  if (n < 0 || n >= (int64_t{1} << 31)) return OFSPMM_ERR_INVALID_ARG;
  if (A->cols == 0 || n == 0) return OFSPMM_OK;
  if (dB == nullptr || (A->nnz > 0 && dY == nullptr)) return OFSPMM_ERR_INVALID_ARG;
  if (At != nullptr) {  // route (1): forward kernel on the cached transpose
    if (int rc = check_csr(At, true)) return rc;
    if (At->rows != A->cols || At->cols != A->rows || At->nnz != A->nnz) return OFSPMM_ERR_INVALID_ARG;
    if (int rc = check_dtypes(At, dense_dtype)) return rc;
    return run_fwd(At, dY, n, dB, n, n, dense_dtype, workspace, workspace_bytes, stream);
  }
  // route (2): vector-atomic scatter into an fp32 accumulator
  const size_t out_elems = static_cast<size_t>(A->cols) * static_cast<size_t>(n);
  if (A->nnz == 0 || A->rows == 0) {
    OFSPMM_CUDA_OK(cudaMemsetAsync(dB, 0, out_elems * dense_size(dense_dtype), stream));
    return OFSPMM_OK;
  }
  const size_t need = ofspmm_bwd_b_workspace_bytes(A->rows, A->cols, A->nnz, n, dense_dtype, 0);
  if (int rc = check_ws(workspace, workspace_bytes, need)) return rc;
  unsigned char* w = static_cast<unsigned char*>(workspace);
  const int64_t P = num_tasks(A->rows, A->nnz);
  const size_t part_bytes = align_up((static_cast<size_t>(P) + 1) * sizeof(int2), 256);
  if (int rc = launch_task_partition(A->crow, A->idx_dtype, A->rows, A->nnz, P, w, stream)) return rc;
  if (dense_dtype == OFSPMM_DTYPE_FLOAT)
    return launch_bwd_atomic(A, dY, static_cast<float*>(dB), nullptr, n, dense_dtype, w, P, stream);
  return launch_bwd_atomic(A, dY, reinterpret_cast<float*>(w + part_bytes), dB, n, dense_dtype, w, P, stream);
}

// Layout of the transient-transpose route: [t_crow | t_col | t_val | transpose scratch | fwd scratch]
namespace {
struct TransientLayout {
  size_t t_crow, t_col, t_val, tws, tws_bytes, fws, total;
};
TransientLayout transient_layout(int64_t rows, int64_t cols, int64_t nnz, int64_t n, int dense_dtype,
                                 int idx_dtype, int val_dtype) {
  TransientLayout L;
  const size_t is = idx_size(idx_dtype), vs = val_dtype == OFSPMM_DTYPE_FLOAT ? 4 : 2;
  size_t off = 0;
  L.t_crow = off; off += align_up(static_cast<size_t>(cols + 1) * is, 256);
  L.t_col = off;  off += align_up(static_cast<size_t>(nnz > 0 ? nnz : 1) * is, 256);
  L.t_val = off;  off += align_up(static_cast<size_t>(nnz > 0 ? nnz : 1) * vs, 256);
  L.tws = off;    L.tws_bytes = transpose_workspace_bytes(rows, cols, nnz, idx_dtype);
  off += align_up(L.tws_bytes, 256);
  L.fws = off;    off += fwd_workspace_layout(cols, nnz, n, dense_dtype).total;
  L.total = off;
  return L;
}
}  // namespace

size_t ofspmm_bwd_b_transient_workspace_bytes(int64_t rows, int64_t cols, int64_t nnz, int64_t n,
                                              int dense_dtype, int idx_dtype, int val_dtype) {
  if (rows < 0 || cols < 0 || nnz < 0 || n < 0 || !idx_ok(idx_dtype)) return 0;
  return transient_layout(rows, cols, nnz, n, dense_dtype, idx_dtype, val_dtype).total;
}

int ofspmm_bwd_b_transient(const ofspmm_csr* A, const void* dY, void* dB, int64_t n, int dense_dtype,
                           void* workspace, size_t workspace_bytes, ofspmm_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = check_csr(A, true)) return rc;
  if (int rc = check_dtypes(A, dense_dtype)) return rc;
  if (n < 0 || n >= (int64_t{1} << 31)) return OFSPMM_ERR_INVALID_ARG;
  if (A->cols == 0 || n == 0) return OFSPMM_OK;
  if (dB == nullptr || (A->nnz > 0 && dY == nullptr)) return OFSPMM_ERR_INVALID_ARG;
  if (A->nnz == 0 || A->rows == 0) {
    OFSPMM_CUDA_OK(cudaMemsetAsync(dB, 0, static_cast<size_t>(A->cols) * n * dense_size(dense_dtype), stream));
    return OFSPMM_OK;
  }
  const TransientLayout L = transient_layout(A->rows, A->cols, A->nnz, n, dense_dtype, A->idx_dtype, A->val_dtype);
  if (int rc = check_ws(workspace, workspace_bytes, L.total)) return rc;
  unsigned char* w = static_cast<unsigned char*>(workspace);
  if (int rc = launch_transpose(A, w + L.t_crow, w + L.t_col, w + L.t_val, nullptr, w + L.tws, L.tws_bytes, stream)) return rc;
  ofspmm_csr At = *A;
  At.rows = A->cols;
  At.cols = A->rows;
  At.crow = w + L.t_crow;
  At.col = w + L.t_col;
  At.val = w + L.t_val;
  return run_fwd(&At, dY, n, dB, n, n, dense_dtype, w + L.fws, workspace_bytes - L.fws, stream);
}

size_t ofspmm_sddmm_workspace_bytes(int64_t rows, int64_t cols, int64_t nnz, int64_t n, int dense_dtype) {
  (void)cols; (void)n; (void)dense_dtype;
  if (rows < 0 || nnz < 0) return 0;
  return align_up((static_cast<size_t>(num_tasks(rows, nnz)) + 1) * sizeof(int2), 256);
}

int ofspmm_sddmm(const ofspmm_csr* A, const void* dY, const void* B, void* dval, int64_t n, int dense_dtype,
                 void* workspace, size_t workspace_bytes, ofspmm_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = check_csr(A, false)) return rc;
  if (int rc = check_dtypes(A, dense_dtype)) return rc;
  if (n < 0 || n >= (int64_t{1} << 31)) return OFSPMM_ERR_INVALID_ARG;
  if (A->nnz == 0) return OFSPMM_OK;
  if (dval == nullptr) return OFSPMM_ERR_INVALID_ARG;
  const size_t val_size = A->val_dtype == OFSPMM_DTYPE_FLOAT ? 4 : 2;
  if (n == 0 || A->cols == 0) {
    OFSPMM_CUDA_OK(cudaMemsetAsync(dval, 0, static_cast<size_t>(A->nnz) * val_size, stream));
    return OFSPMM_OK;
  }
  if (dY == nullptr || B == nullptr) return OFSPMM_ERR_INVALID_ARG;
  const size_t need = ofspmm_sddmm_workspace_bytes(A->rows, A->cols, A->nnz, n, dense_dtype);
  if (int rc = check_ws(workspace, workspace_bytes, need)) return rc;
  const int64_t P = num_tasks(A->rows, A->nnz);
  if (int rc = launch_task_partition(A->crow, A->idx_dtype, A->rows, A->nnz, P, workspace, stream)) return rc;
  return launch_sddmm(A, dY, B, dval, n, dense_dtype, workspace, P, stream);
}

int ofspmm_partition(const void* crow, int idx_dtype, int64_t rows, int64_t nnz, int64_t parts,
                     int64_t* out_row, int64_t* out_nz, ofspmm_stream_t stream) {
  if (crow == nullptr || out_row == nullptr || out_nz == nullptr) return OFSPMM_ERR_INVALID_ARG;
  if (rows < 0 || nnz < 0 || parts < 1) return OFSPMM_ERR_INVALID_ARG;
  if (!idx_ok(idx_dtype)) return OFSPMM_ERR_UNSUPPORTED_DTYPE;
  return launch_partition_public(crow, idx_dtype, rows, nnz, parts, out_row, out_nz,
                                 reinterpret_cast<cudaStream_t>(stream));
}

int ofspmm_partition_host(const void* crow, int idx_dtype, int64_t rows, int64_t nnz, int64_t parts,
                          int64_t* out_row, int64_t* out_nz) {
  if (crow == nullptr || out_row == nullptr || out_nz == nullptr) return OFSPMM_ERR_INVALID_ARG;
  if (rows < 0 || nnz < 0 || parts < 1) return OFSPMM_ERR_INVALID_ARG;
  if (!idx_ok(idx_dtype)) return OFSPMM_ERR_UNSUPPORTED_DTYPE;
  const int64_t total = rows + nnz;
  const int64_t ipw = (total + parts - 1) / parts;
  for (int64_t k = 0; k <= parts; ++k) {
    int64_t d = k * ipw;
    if (d > total) d = total;
    const int64_t r = idx_dtype == OFSPMM_DTYPE_INT32
                          ? merge_path_search<int32_t>(static_cast<const int32_t*>(crow), rows, nnz, d)
                          : merge_path_search<int64_t>(static_cast<const int64_t*>(crow), rows, nnz, d);
    out_row[k] = r;
    out_nz[k] = d - r;
  }
  return OFSPMM_OK;
}

int ofspmm_row_hist(const void* crow, int idx_dtype, int64_t rows, int64_t* hist32, ofspmm_stream_t stream) {
  if (crow == nullptr || hist32 == nullptr || rows < 0) return OFSPMM_ERR_INVALID_ARG;
  if (!idx_ok(idx_dtype)) return OFSPMM_ERR_UNSUPPORTED_DTYPE;
  return launch_row_hist(crow, idx_dtype, rows, hist32, reinterpret_cast<cudaStream_t>(stream));
}

size_t ofspmm_csr_transpose_workspace_bytes(int64_t rows, int64_t cols, int64_t nnz, int idx_dtype) {
  if (rows < 0 || cols < 0 || nnz < 0 || !idx_ok(idx_dtype)) return 0;
  return transpose_workspace_bytes(rows, cols, nnz, idx_dtype);
}

int ofspmm_csr_transpose(const ofspmm_csr* A, void* t_crow, void* t_col, void* t_val, void* t_perm,
                         void* workspace, size_t workspace_bytes, ofspmm_stream_t stream) {
  if (int rc = check_csr(A, false)) return rc;
  if (t_crow == nullptr || (A->nnz > 0 && t_col == nullptr)) return OFSPMM_ERR_INVALID_ARG;
  return launch_transpose(A, t_crow, t_col, t_val, t_perm, workspace, workspace_bytes,
                          reinterpret_cast<cudaStream_t>(stream));
}

size_t ofspmm_fwd_host_workspace_bytes(int64_t rows, int64_t cols, int64_t nnz, int64_t n, int dense_dtype,
                                       int idx_dtype, int val_dtype) {
  if (rows < 0 || cols < 0 || nnz < 0 || n < 0) return 0;
  const size_t vs = val_dtype == OFSPMM_DTYPE_FLOAT ? 4 : 2;
  size_t total = 0;
  total += align_up(static_cast<size_t>(rows + 1) * idx_size(idx_dtype), 256);
  total += align_up(static_cast<size_t>(nnz) * idx_size(idx_dtype), 256);
  total += align_up(static_cast<size_t>(nnz) * vs, 256);
  total += align_up(static_cast<size_t>(cols) * n * dense_size(dense_dtype), 256);
  total += align_up(static_cast<size_t>(rows) * n * dense_size(dense_dtype), 256);
  total += fwd_workspace_layout(rows, nnz, n, dense_dtype).total;
  return total;
}

int ofspmm_fwd_host(const ofspmm_csr* Ah, const void* B_host, void* C_host, int64_t n, int dense_dtype,
                    void* workspace, size_t workspace_bytes, ofspmm_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = check_csr(Ah, true)) return rc;
  if (int rc = check_dtypes(Ah, dense_dtype)) return rc;
  if (n < 0 || n >= (int64_t{1} << 31)) return OFSPMM_ERR_INVALID_ARG;
  if (Ah->rows == 0 || n == 0) return OFSPMM_OK;
  if (C_host == nullptr || (Ah->nnz > 0 && Ah->cols > 0 && B_host == nullptr)) return OFSPMM_ERR_INVALID_ARG;
  const size_t need = ofspmm_fwd_host_workspace_bytes(Ah->rows, Ah->cols, Ah->nnz, n, dense_dtype,
                                                      Ah->idx_dtype, Ah->val_dtype);
  if (int rc = check_ws(workspace, workspace_bytes, need)) return rc;
  const size_t is = idx_size(Ah->idx_dtype), vs = Ah->val_dtype == OFSPMM_DTYPE_FLOAT ? 4 : 2;
  const size_t ds = dense_size(dense_dtype);
  unsigned char* w = static_cast<unsigned char*>(workspace);
  size_t off = 0;
  auto carve = [&](size_t bytes) { void* p = w + off; off += align_up(bytes, 256); return p; };
  void* d_crow = carve(static_cast<size_t>(Ah->rows + 1) * is);
  void* d_col = carve(static_cast<size_t>(Ah->nnz) * is);
  void* d_val = carve(static_cast<size_t>(Ah->nnz) * vs);
  void* d_B = carve(static_cast<size_t>(Ah->cols) * n * ds);
  void* d_C = carve(static_cast<size_t>(Ah->rows) * n * ds);
  OFSPMM_CUDA_OK(cudaMemcpyAsync(d_crow, Ah->crow, static_cast<size_t>(Ah->rows + 1) * is, cudaMemcpyHostToDevice, stream));
  if (Ah->nnz > 0) {
    OFSPMM_CUDA_OK(cudaMemcpyAsync(d_col, Ah->col, static_cast<size_t>(Ah->nnz) * is, cudaMemcpyHostToDevice, stream));
    OFSPMM_CUDA_OK(cudaMemcpyAsync(d_val, Ah->val, static_cast<size_t>(Ah->nnz) * vs, cudaMemcpyHostToDevice, stream));
  }
  if (Ah->cols > 0 && B_host != nullptr)
    OFSPMM_CUDA_OK(cudaMemcpyAsync(d_B, B_host, static_cast<size_t>(Ah->cols) * n * ds, cudaMemcpyHostToDevice, stream));
  ofspmm_csr Ad = *Ah;
  Ad.crow = d_crow;
  Ad.col = d_col;
  Ad.val = d_val;
  if (int rc = run_fwd(&Ad, d_B, n, d_C, n, n, dense_dtype, w + off, workspace_bytes - off, stream)) return rc;
  OFSPMM_CUDA_OK(cudaMemcpyAsync(C_host, d_C, static_cast<size_t>(Ah->rows) * n * ds, cudaMemcpyDeviceToHost, stream));
  return OFSPMM_OK;
}

const char* ofspmm_strerror(int status) {
  switch (status) {
    case OFSPMM_OK: return "ok";
    case OFSPMM_ERR_INVALID_ARG: return "invalid argument (null pointer, negative size or inconsistent shape)";
    case OFSPMM_ERR_UNSUPPORTED_DTYPE: return "unsupported dtype combination";
    case OFSPMM_ERR_WORKSPACE: return "workspace missing, misaligned or smaller than *_workspace_bytes()";
    case OFSPMM_ERR_CUDA: return "CUDA runtime call or kernel launch failed";
    case OFSPMM_ERR_TOO_LARGE: return "rows or nnz >= 2^31-1";
    case OFSPMM_ERR_NO_DEVICE: return "no sm_100 (B200) device current on this thread";
    default: return "unknown status";
  }
}

int ofspmm_version(void) { return 100; }

uint64_t ofspmm_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

const char* ofspmm_fwd_variant(int64_t rows, int64_t nnz, int64_t n, int dense_dtype) {
  (void)rows; (void)nnz;
  return fwd_variant_name(n, dense_dtype, true);
}

}  // extern "C"

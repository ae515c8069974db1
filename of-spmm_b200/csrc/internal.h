// internal.h — launcher interface between the C ABI (api.cu) and the per-op translation units.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <atomic>

#include "../../include/ofspmm.h"

namespace ofspmm {

#ifndef OFSPMM_ITEMS
#define OFSPMM_ITEMS 256
#endif
#ifndef OFSPMM_WARPS
#define OFSPMM_WARPS 4
#endif
constexpr int kTaskItems = OFSPMM_ITEMS;  // merge items (row-ends + non-zeros) per warp task
constexpr int kWarpsPerCta = OFSPMM_WARPS;

extern std::atomic<uint64_t> g_launches;
inline void count_launch(uint64_t k = 1) { g_launches.fetch_add(k, std::memory_order_relaxed); }

inline int64_t num_tasks(int64_t rows, int64_t nnz) {
  const int64_t total = rows + nnz;
  return (total + kTaskItems - 1) / kTaskItems;
}
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct DevInfo {
  int sms;
  int cc_major;
};
int get_dev_info(DevInfo* out);  // OFSPMM_OK / OFSPMM_ERR_CUDA

// Layout of the workspace shared by fwd / bwd(transpose) calls.
struct FwdWorkspace {
  size_t part_off, carry_off, head_off, total;
};
FwdWorkspace fwd_workspace_layout(int64_t rows, int64_t nnz, int64_t n, int dense_dtype);

// Each returns an OFSPMM_* status; all launches go to `stream`.
int launch_task_partition(const void* crow, int idx_dtype, int64_t rows, int64_t nnz, int64_t P,
                          void* part, cudaStream_t stream);
int launch_fwd(const ofspmm_csr* A, const void* B, int64_t ldb, void* C, int64_t ldc, int64_t n,
               int dense_dtype, const void* part, float* carry, float* head, int64_t P,
               cudaStream_t stream);
int launch_sddmm(const ofspmm_csr* A, const void* dY, const void* B, void* dval, int64_t n,
                 int dense_dtype, const void* part, int64_t P, cudaStream_t stream);
int launch_bwd_atomic(const ofspmm_csr* A, const void* dY, float* acc, void* dB_cast_out,
                      int64_t n, int dense_dtype, const void* part, int64_t P, cudaStream_t stream);
int launch_partition_public(const void* crow, int idx_dtype, int64_t rows, int64_t nnz,
                            int64_t parts, int64_t* out_row, int64_t* out_nz, cudaStream_t stream);
int launch_row_hist(const void* crow, int idx_dtype, int64_t rows, int64_t* hist32,
                    cudaStream_t stream);
size_t transpose_workspace_bytes(int64_t rows, int64_t cols, int64_t nnz, int idx_dtype);
int launch_transpose(const ofspmm_csr* A, void* t_crow, void* t_col, void* t_val, void* t_perm,
                     void* ws, size_t ws_bytes, cudaStream_t stream);
const char* fwd_variant_name(int64_t n, int dense_dtype, bool aligned);

#define OFSPMM_CUDA_OK(expr)                         \
  do {                                               \
    if ((expr) != cudaSuccess) return OFSPMM_ERR_CUDA; \
  } while (0)

}  // namespace ofspmm

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

constexpr int kSmallTaskItems = 64;       // task size of the small-problem family
// Below this many 256-item tasks the machine (148 SMs x 36 resident warps) is not filled by one
// task per warp: AUTO switches to 64-item tasks so 4x more warps work in parallel.
constexpr int64_t kSmallProblemTasks = 148 * 36;

inline int64_t num_tasks(int64_t rows, int64_t nnz, int items = kTaskItems) {
  const int64_t total = rows + nnz;
  return (total + items - 1) / items;
}

// Resolved kernel family of a forward launch (public encoding: OFSPMM_VARIANT_*).
struct FwdVariant {
  int items;          // merge items per warp task: kTaskItems or kSmallTaskItems
  bool row_parallel;  // sub-warp per row (short rows, narrow dense operand) instead of nnz-parallel
  bool whole_rows;    // spmm_rows_kernel: one launch, no merge path (items = kSmallTaskItems for sizing)
};
FwdVariant resolve_variant(int variant, int64_t rows, int64_t nnz, int64_t n, int dense_dtype);
struct FwdLaunch;
bool rows_kernel_applies(const ofspmm_csr* A, const void* B, int64_t ldb, const void* C, int64_t ldc, int64_t n,
                         int dense_dtype, const FwdLaunch& L);
int encode_variant(const FwdVariant& v);

// Per-launch options the C ABI passes down (ofspmm_opts).
struct FwdLaunch {
  FwdVariant variant;
  int tasks_per_warp;  // 0 = persistent grid
  int reserve_ctas;    // CTA slots per SM a persistent grid leaves free for overlapping exchange kernels
  bool dynamic;        // tasks drawn from a global counter (persistent grid only)
  unsigned flags;      // kFwd* epilogue bits
  const void* bias;
  float* acc32;        // fp32 accumulator of multi-pass 16-bit products
};
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct DevInfo {
  int sms;
  int cc_major;
  int ordinal;
};

// Launch constants of one kernel instantiation per device ordinal (0 = not queried yet); values
// are immutable once written, relaxed atomics make concurrent first calls benign.
struct KernelLaunchCache {
  static constexpr int kMaxDevices = 64;
  std::atomic<int> occ[kMaxDevices];
  KernelLaunchCache() { for (auto& o : occ) o.store(0, std::memory_order_relaxed); }
  int get(int dev) const { return dev >= 0 && dev < kMaxDevices ? occ[dev].load(std::memory_order_relaxed) : 0; }
  void set(int dev, int v) { if (dev >= 0 && dev < kMaxDevices) occ[dev].store(v, std::memory_order_relaxed); }
};
int get_dev_info(DevInfo* out);  // OFSPMM_OK / OFSPMM_ERR_CUDA

// Layout of the workspace shared by fwd / bwd(transpose) calls.
struct FwdWorkspace {
  size_t counter_off, part_off, carry_off, head_off, total;
};
FwdWorkspace fwd_workspace_layout(int64_t rows, int64_t nnz, int64_t n, int dense_dtype, int items = kTaskItems);

// Each returns an OFSPMM_* status; all launches go to `stream`.
int launch_task_partition(const void* crow, int idx_dtype, int64_t rows, int64_t nnz, int64_t P,
                          int items, void* part, cudaStream_t stream);
int launch_fwd(const ofspmm_csr* A, const void* B, int64_t ldb, void* C, int64_t ldc, int64_t n,
               int dense_dtype, const void* part, float* carry, float* head, void* counter, int64_t P,
               const FwdLaunch& L, cudaStream_t stream);
int launch_sddmm(const ofspmm_csr* A, const void* dY, const void* B, void* dval, int64_t n,
                 int dense_dtype, const void* part, void* counter, int64_t P, cudaStream_t stream);
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
size_t coo_to_csr_workspace_bytes(int64_t n, int64_t rows, int64_t cols);
int launch_coo_to_csr(const int64_t* row, const int64_t* col, const float* val, int64_t n, int64_t rows, int64_t cols,
                      int mode, int idx_dtype, void* crow, void* col_out, float* val_out, int64_t* counts, void* ws,
                      size_t ws_bytes, cudaStream_t stream);
int launch_expand_rows(const void* crow, int idx_dtype, int64_t rows, int64_t* out, cudaStream_t stream);
int launch_csr_normalize(const void* crow, const void* col, float* val, int idx_dtype, int64_t rows, int64_t cols, int mode,
                         float* dinv, cudaStream_t stream);
int launch_gather_vals(const void* val, int val_dtype, const void* perm, int idx_dtype, int64_t nnz,
                       void* out, cudaStream_t stream);
int launch_gather_rows(void* dst, int64_t ld_dst, const void* src, int64_t ld_src, const void* list,
                       int idx_dtype, int64_t idx_offset, int64_t count, int64_t n, int dense_dtype,
                       int max_ctas, cudaStream_t stream);
int launch_scatter_add_rows_f32(float* dst, int64_t ld_dst, const void* src, int64_t ld_src, const void* list,
                                int idx_dtype, int64_t idx_offset, int64_t count, int64_t n, int src_dtype,
                                int max_ctas, cudaStream_t stream);
int launch_signal_peers(void* const* slots, int n, unsigned long long epoch, cudaStream_t stream);
int launch_pull_rows_multi(void* dst, int64_t ld_dst, int64_t ld_src, const ofspmm_pull_seg* segs, int nseg,
                           unsigned long long epoch, int64_t n, int dense_dtype, int idx_dtype, int max_ctas,
                           cudaStream_t stream);
int launch_combine_rows_multi(float* acc, int64_t ld_acc, int64_t ld_src, const ofspmm_combine_seg* segs, int nseg,
                              unsigned long long epoch, int64_t rows, int64_t n, int src_dtype, int max_ctas,
                              cudaStream_t stream);
int launch_cast_from_f32(const float* src, void* dst, int64_t count, int dst_dtype, cudaStream_t stream);
int launch_scatter_add_rows(void* dst, int64_t ld_dst, const void* src, int64_t ld_src, const void* list,
                            int idx_dtype, int64_t idx_offset, int64_t count, int64_t n, int dense_dtype,
                            int max_ctas, cudaStream_t stream);

#define OFSPMM_CUDA_OK(expr)                         \
  do {                                               \
    if ((expr) != cudaSuccess) return OFSPMM_ERR_CUDA; \
  } while (0)

}  // namespace ofspmm

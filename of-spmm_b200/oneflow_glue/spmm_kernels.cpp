// oneflow/user/kernels/spmm_kernels.cpp — REGISTER_USER_KERNEL glue for spmm_csr,
// spmm_csr_grad_b and sddmm_csr.  Every Compute() body pulls raw pointers /
// shapes / attrs / the cudaStream_t out of the KernelComputeContext and calls the C ABI of
// libofspmm_b200.so (include/ofspmm.h).  CUDA only: exactly one registered kernel may match a
// context (oneflow/core/framework/user_op_registry_manager.cpp:93-117), and per the north star
// there is no CPU kernel on this path.
//
// Conventions followed: kernel class + registration macro over (dense dtype × index dtype) as in
// oneflow/user/kernels/gather_kernel.cpp:116-136; tmp_buffer via SetInferTmpSizeFn as in
// unsorted_segment_sum_kernel.cpp:191-202; per-op persistent data in OpKernelState
// (stateful_opkernel.cpp:919-928); fatal CHECK on a non-zero status like OF_CUDA_CHECK
// (oneflow/core/device/cuda_util.h:54-57); CudaGraphSupport (core/kernel/cuda_graph_support.h:28-42).
#ifdef WITH_CUDA
#include "oneflow/core/framework/framework.h"
#include "oneflow/core/kernel/cuda_graph_support.h"
#include "oneflow/core/ep/cuda/cuda_stream.h"
#include "ofspmm.h"

namespace oneflow {

namespace {

ofspmm_stream_t StreamOf(user_op::KernelComputeContext* ctx) {
  return reinterpret_cast<ofspmm_stream_t>(ctx->stream()->As<ep::CudaStream>()->cuda_stream());
}

// oneflow::DataType values are passed through unchanged (ofspmm.h reuses the numbering).
ofspmm_csr MakeCsr(user_op::KernelComputeContext* ctx, const user_op::Tensor* val, DataType val_dtype) {
  const user_op::Tensor* crow = ctx->Tensor4ArgNameAndIndex("a_crow", 0);
  const user_op::Tensor* col = ctx->Tensor4ArgNameAndIndex("a_col", 0);
  ofspmm_csr a;
  a.rows = ctx->Attr<int64_t>("a_rows");
  a.cols = ctx->Attr<int64_t>("a_cols");
  a.nnz = col->shape_view().elem_cnt();
  a.crow = crow->raw_dptr();
  a.col = col->raw_dptr();
  a.val = val != nullptr ? val->raw_dptr() : nullptr;
  a.idx_dtype = static_cast<int32_t>(crow->data_type());
  a.val_dtype = static_cast<int32_t>(val_dtype);
  return a;
}

#define OFSPMM_CHECK(expr)                                                              \
  do {                                                                                  \
    const int ofspmm_rc = (expr);                                                       \
    CHECK_EQ(ofspmm_rc, OFSPMM_OK) << "ofspmm: " << ofspmm_strerror(ofspmm_rc);         \
  } while (0)

// A^T kept across calls for the deterministic, atomic-free backward (route 1 of ofspmm_bwd_b).
// Keyed by the CSR device pointers; rebuilt when they change.  Device buffers come from the
// stream's device allocator once, outside CUDA-graph capture (first eager call / graph warm-up).
class SpmmTransposeState final : public user_op::OpKernelState {
 public:
  ~SpmmTransposeState() override { Release(); }
  const ofspmm_csr* GetOrBuild(user_op::KernelComputeContext* ctx, const ofspmm_csr& a);

 private:
  void Release();
  ep::Device* device_ = nullptr;
  const void *key_crow_ = nullptr, *key_col_ = nullptr, *key_val_ = nullptr;
  void *t_crow_ = nullptr, *t_col_ = nullptr, *t_val_ = nullptr, *ws_ = nullptr;
  ofspmm_csr at_{};
};

void SpmmTransposeState::Release() {
  if (device_ == nullptr) { return; }
  for (void* p : {t_crow_, t_col_, t_val_, ws_}) {
    if (p != nullptr) { device_->Free(ep::AllocationOptions{}, p); }
  }
  t_crow_ = t_col_ = t_val_ = ws_ = nullptr;
}

const ofspmm_csr* SpmmTransposeState::GetOrBuild(user_op::KernelComputeContext* ctx,
                                                 const ofspmm_csr& a) {
  if (a.crow == key_crow_ && a.col == key_col_ && a.val == key_val_) { return &at_; }
  Release();
  device_ = ctx->stream()->device();
  const size_t isz = a.idx_dtype == OFSPMM_DTYPE_INT64 ? 8 : 4;
  const size_t vsz = a.val_dtype == OFSPMM_DTYPE_FLOAT ? 4 : 2;
  const size_t ws_bytes = ofspmm_csr_transpose_workspace_bytes(a.rows, a.cols, a.nnz, a.idx_dtype);
  CHECK_JUST(device_->Alloc(ep::AllocationOptions{}, &t_crow_, (a.cols + 1) * isz));
  CHECK_JUST(device_->Alloc(ep::AllocationOptions{}, &t_col_, std::max<size_t>(a.nnz, 1) * isz));
  CHECK_JUST(device_->Alloc(ep::AllocationOptions{}, &t_val_, std::max<size_t>(a.nnz, 1) * vsz));
  CHECK_JUST(device_->Alloc(ep::AllocationOptions{}, &ws_, ws_bytes));
  OFSPMM_CHECK(ofspmm_csr_transpose(&a, t_crow_, t_col_, t_val_, nullptr, ws_, ws_bytes, StreamOf(ctx)));
  at_ = a;
  at_.rows = a.cols;
  at_.cols = a.rows;
  at_.crow = t_crow_;
  at_.col = t_col_;
  at_.val = t_val_;
  key_crow_ = a.crow;
  key_col_ = a.col;
  key_val_ = a.val;
  return &at_;
}

// ---------------------------------------------------------------- spmm_csr
class SpmmCsrKernel final : public user_op::OpKernel, public user_op::CudaGraphSupport {
 public:
  SpmmCsrKernel() = default;
  ~SpmmCsrKernel() override = default;

 private:
  void Compute(user_op::KernelComputeContext* ctx) const override {
    const user_op::Tensor* val = ctx->Tensor4ArgNameAndIndex("a_val", 0);
    const user_op::Tensor* b = ctx->Tensor4ArgNameAndIndex("b", 0);
    user_op::Tensor* out = ctx->Tensor4ArgNameAndIndex("out", 0);
    user_op::Tensor* tmp = ctx->Tensor4ArgNameAndIndex("tmp_buffer", 0);
    const ofspmm_csr a = MakeCsr(ctx, val, val->data_type());
    OFSPMM_CHECK(ofspmm_fwd(&a, b->raw_dptr(), out->mut_raw_dptr(), b->shape_view().At(1),
                            static_cast<int>(b->data_type()), tmp->mut_raw_dptr(),
                            tmp->shape_view().elem_cnt(), StreamOf(ctx)));
  }
  bool AlwaysComputeWhenAllOutputsEmpty() const override { return false; }
};

size_t InferSpmmTmpSize(user_op::InferContext* ctx) {
  const int64_t nnz = ctx->InputShape("a_col", 0).elem_cnt();
  return ofspmm_fwd_workspace_bytes(ctx->Attr<int64_t>("a_rows"), ctx->Attr<int64_t>("a_cols"), nnz,
                                    ctx->InputShape("b", 0).At(1),
                                    static_cast<int>(ctx->InputDType("b", 0)));
}

// ---------------------------------------------------------------- spmm_csr_grad_b
class SpmmCsrGradBKernel final : public user_op::OpKernel, public user_op::CudaGraphSupport {
 public:
  SpmmCsrGradBKernel() = default;
  ~SpmmCsrGradBKernel() override = default;

  std::shared_ptr<user_op::OpKernelState> CreateOpKernelState(
      user_op::KernelInitContext*) const override {
    return std::make_shared<SpmmTransposeState>();
  }

 private:
  using user_op::OpKernel::Compute;
  void Compute(user_op::KernelComputeContext* ctx, user_op::OpKernelState* state,
               const user_op::OpKernelCache*) const override {
    const user_op::Tensor* val = ctx->Tensor4ArgNameAndIndex("a_val", 0);
    const user_op::Tensor* dy = ctx->Tensor4ArgNameAndIndex("dy", 0);
    user_op::Tensor* db = ctx->Tensor4ArgNameAndIndex("db", 0);
    user_op::Tensor* tmp = ctx->Tensor4ArgNameAndIndex("tmp_buffer", 0);
    const ofspmm_csr a = MakeCsr(ctx, val, val->data_type());
    auto* tstate = dynamic_cast<SpmmTransposeState*>(state);
    CHECK_NOTNULL(tstate);
    const ofspmm_csr* at = tstate->GetOrBuild(ctx, a);
    OFSPMM_CHECK(ofspmm_bwd_b(&a, at, dy->raw_dptr(), db->mut_raw_dptr(), dy->shape_view().At(1),
                              static_cast<int>(dy->data_type()), tmp->mut_raw_dptr(),
                              tmp->shape_view().elem_cnt(), StreamOf(ctx)));
  }
  bool AlwaysComputeWhenAllOutputsEmpty() const override { return false; }
};

size_t InferGradBTmpSize(user_op::InferContext* ctx) {
  const int64_t nnz = ctx->InputShape("a_col", 0).elem_cnt();
  return ofspmm_bwd_b_workspace_bytes(ctx->Attr<int64_t>("a_rows"), ctx->Attr<int64_t>("a_cols"), nnz,
                                      ctx->InputShape("dy", 0).At(1),
                                      static_cast<int>(ctx->InputDType("dy", 0)), /*have_transpose=*/1);
}

// ---------------------------------------------------------------- sddmm_csr
class SddmmCsrKernel final : public user_op::OpKernel, public user_op::CudaGraphSupport {
 public:
  SddmmCsrKernel() = default;
  ~SddmmCsrKernel() override = default;

 private:
  void Compute(user_op::KernelComputeContext* ctx) const override {
    const user_op::Tensor* dy = ctx->Tensor4ArgNameAndIndex("dy", 0);
    const user_op::Tensor* b = ctx->Tensor4ArgNameAndIndex("b", 0);
    user_op::Tensor* dval = ctx->Tensor4ArgNameAndIndex("dval", 0);
    user_op::Tensor* tmp = ctx->Tensor4ArgNameAndIndex("tmp_buffer", 0);
    const ofspmm_csr a = MakeCsr(ctx, nullptr, dval->data_type());
    OFSPMM_CHECK(ofspmm_sddmm(&a, dy->raw_dptr(), b->raw_dptr(), dval->mut_raw_dptr(),
                              b->shape_view().At(1), static_cast<int>(b->data_type()),
                              tmp->mut_raw_dptr(), tmp->shape_view().elem_cnt(), StreamOf(ctx)));
  }
  bool AlwaysComputeWhenAllOutputsEmpty() const override { return false; }
};

size_t InferSddmmTmpSize(user_op::InferContext* ctx) {
  const int64_t nnz = ctx->InputShape("a_col", 0).elem_cnt();
  return ofspmm_sddmm_workspace_bytes(ctx->Attr<int64_t>("a_rows"), ctx->Attr<int64_t>("a_cols"), nnz,
                                      ctx->InputShape("b", 0).At(1),
                                      static_cast<int>(ctx->InputDType("b", 0)));
}

}  // namespace

#define REGISTER_SPMM_KERNELS(dense_dtype, index_dtype)                                        \
  REGISTER_USER_KERNEL("spmm_csr")                                                             \
      .SetCreateFn<SpmmCsrKernel>()                                                            \
      .SetIsMatchedHob((user_op::HobDeviceType() == DeviceType::kCUDA)                         \
                       && (user_op::HobDataType("b", 0) == dense_dtype)                        \
                       && (user_op::HobDataType("a_col", 0) == index_dtype))                   \
      .SetInferTmpSizeFn(InferSpmmTmpSize);                                                    \
  REGISTER_USER_KERNEL("spmm_csr_grad_b")                                                      \
      .SetCreateFn<SpmmCsrGradBKernel>()                                                       \
      .SetIsMatchedHob((user_op::HobDeviceType() == DeviceType::kCUDA)                         \
                       && (user_op::HobDataType("dy", 0) == dense_dtype)                       \
                       && (user_op::HobDataType("a_col", 0) == index_dtype))                   \
      .SetInferTmpSizeFn(InferGradBTmpSize);                                                   \
  REGISTER_USER_KERNEL("sddmm_csr")                                                            \
      .SetCreateFn<SddmmCsrKernel>()                                                           \
      .SetIsMatchedHob((user_op::HobDeviceType() == DeviceType::kCUDA)                         \
                       && (user_op::HobDataType("b", 0) == dense_dtype)                        \
                       && (user_op::HobDataType("a_col", 0) == index_dtype))                   \
      .SetInferTmpSizeFn(InferSddmmTmpSize);

REGISTER_SPMM_KERNELS(DataType::kFloat, DataType::kInt32)
REGISTER_SPMM_KERNELS(DataType::kFloat, DataType::kInt64)
REGISTER_SPMM_KERNELS(DataType::kBFloat16, DataType::kInt32)
REGISTER_SPMM_KERNELS(DataType::kBFloat16, DataType::kInt64)

}  // namespace oneflow
#endif  // WITH_CUDA

// oneflow/user/ops/spmm_op.cpp — shape / dtype / SBP inference of spmm_csr, spmm_csr_grad_b and
// sddmm_csr.  Written against the reference's user-op framework; conventions
// from oneflow/user/ops/matmul_op.cpp:23-138 and unsorted_segment_sum_op.cpp:66-78.
// The Python mirror with the same checks is of-spmm_b200/ops.py:infer_spmm_csr.
#include "oneflow/core/framework/framework.h"
#include "oneflow/core/framework/op_generated.h"

namespace oneflow {

namespace {

Maybe<void> CheckCsr(user_op::InferContext* ctx, int64_t* nnz) {
  const Shape& crow = ctx->InputShape("a_crow", 0);
  const Shape& col = ctx->InputShape("a_col", 0);
  const int64_t a_rows = ctx->Attr<int64_t>("a_rows");
  const int64_t a_cols = ctx->Attr<int64_t>("a_cols");
  CHECK_EQ_OR_RETURN(crow.NumAxes(), 1) << "a_crow must be 1-D";
  CHECK_EQ_OR_RETURN(col.NumAxes(), 1) << "a_col must be 1-D";
  CHECK_GE_OR_RETURN(a_rows, 0) << "a_rows must be non-negative";
  CHECK_GE_OR_RETURN(a_cols, 0) << "a_cols must be non-negative";
  CHECK_EQ_OR_RETURN(crow.At(0), a_rows + 1) << "a_crow must have a_rows+1 entries";
  *nnz = col.At(0);
  return Maybe<void>::Ok();
}

Maybe<void> CheckIndexTypes(user_op::InferContext* ctx) {
  CHECK_OR_RETURN(IsIndexDataType(ctx->InputDType("a_crow", 0))) << "a_crow must be an index dtype";
  CHECK_EQ_OR_RETURN(ctx->InputDType("a_col", 0), ctx->InputDType("a_crow", 0))
      << "a_col and a_crow must share one index dtype";
  return Maybe<void>::Ok();
}

Maybe<void> NoGradForIndices(const user_op::GetInputArgModifier& GetInputArgModifierFn) {
  for (const char* name : {"a_crow", "a_col"}) {
    user_op::InputArgModifier* m = GetInputArgModifierFn(name, 0);
    CHECK_NOTNULL_OR_RETURN(m);  // NOLINT(maybe-need-error-msg)
    m->set_requires_grad(false);
  }
  return Maybe<void>::Ok();
}

// A split of a_col / a_val is not a row split of A, so the CSR arrays are always broadcast; the
// dense side may be column-split (always legal for a row-wise linear map).  nnz-balanced row
// blocks live inside the library (of-spmm_b200/dist.py), outside SBP.
Maybe<void> DenseColumnSplitSbp(user_op::SbpContext* ctx, const char* dense_in, const char* out) {
  ctx->NewBuilder()
      .Broadcast(user_op::OpArg("a_crow", 0))
      .Broadcast(user_op::OpArg("a_col", 0))
      .Broadcast(user_op::OpArg("a_val", 0))
      .Split(user_op::OpArg(dense_in, 0), 1)
      .Split(user_op::OpArg(out, 0), 1)
      .Build();
  ctx->NewBuilder().Broadcast(ctx->inputs()).Broadcast(ctx->outputs()).Build();
  return Maybe<void>::Ok();
}

}  // namespace

// ---------------------------------------------------------------- spmm_csr
/*static*/ Maybe<void> SpmmCsrOp::InferLogicalTensorDesc(user_op::InferContext* ctx) {
  int64_t nnz = 0;
  JUST(CheckCsr(ctx, &nnz));
  const Shape& b = ctx->InputShape("b", 0);
  CHECK_EQ_OR_RETURN(b.NumAxes(), 2) << "b must be 2-D (a_cols x n)";
  CHECK_EQ_OR_RETURN(b.At(0), ctx->Attr<int64_t>("a_cols")) << "b rows must equal a_cols";
  CHECK_EQ_OR_RETURN(ctx->InputShape("a_val", 0).elem_cnt(), nnz) << "a_val and a_col must have nnz entries";
  ctx->SetOutputShape("out", 0, Shape({ctx->Attr<int64_t>("a_rows"), b.At(1)}));
  return Maybe<void>::Ok();
}
/*static*/ Maybe<void> SpmmCsrOp::InferPhysicalTensorDesc(user_op::InferContext* ctx) {
  return InferLogicalTensorDesc(ctx);
}
/*static*/ Maybe<void> SpmmCsrOp::InferDataType(user_op::InferContext* ctx) {
  JUST(CheckIndexTypes(ctx));
  const DataType dense = ctx->InputDType("b", 0);
  const DataType val = ctx->InputDType("a_val", 0);
  CHECK_OR_RETURN(val == dense || val == DataType::kFloat) << "a_val must be float32 or match b";
  ctx->SetOutputDType("out", 0, dense);
  return Maybe<void>::Ok();
}
/*static*/ Maybe<void> SpmmCsrOp::GetSbp(user_op::SbpContext* ctx) {
  return DenseColumnSplitSbp(ctx, "b", "out");
}
/*static*/ Maybe<void> SpmmCsrOp::ModifyInputArg(const GetInputArgModifier& fn,
                                                 const user_op::UserOpConfWrapper&) {
  return NoGradForIndices(fn);
}

// ---------------------------------------------------------------- spmm_csr_grad_b
/*static*/ Maybe<void> SpmmCsrGradBOp::InferLogicalTensorDesc(user_op::InferContext* ctx) {
  int64_t nnz = 0;
  JUST(CheckCsr(ctx, &nnz));
  const Shape& dy = ctx->InputShape("dy", 0);
  CHECK_EQ_OR_RETURN(dy.NumAxes(), 2) << "dy must be 2-D (a_rows x n)";
  CHECK_EQ_OR_RETURN(dy.At(0), ctx->Attr<int64_t>("a_rows")) << "dy rows must equal a_rows";
  ctx->SetOutputShape("db", 0, Shape({ctx->Attr<int64_t>("a_cols"), dy.At(1)}));
  return Maybe<void>::Ok();
}
/*static*/ Maybe<void> SpmmCsrGradBOp::InferPhysicalTensorDesc(user_op::InferContext* ctx) {
  return InferLogicalTensorDesc(ctx);
}
/*static*/ Maybe<void> SpmmCsrGradBOp::InferDataType(user_op::InferContext* ctx) {
  JUST(CheckIndexTypes(ctx));
  ctx->SetOutputDType("db", 0, ctx->InputDType("dy", 0));
  return Maybe<void>::Ok();
}
/*static*/ Maybe<void> SpmmCsrGradBOp::GetSbp(user_op::SbpContext* ctx) {
  return DenseColumnSplitSbp(ctx, "dy", "db");
}
/*static*/ Maybe<void> SpmmCsrGradBOp::ModifyInputArg(const GetInputArgModifier& fn,
                                                      const user_op::UserOpConfWrapper&) {
  return NoGradForIndices(fn);
}

// ---------------------------------------------------------------- sddmm_csr
/*static*/ Maybe<void> SddmmCsrOp::InferLogicalTensorDesc(user_op::InferContext* ctx) {
  int64_t nnz = 0;
  JUST(CheckCsr(ctx, &nnz));
  const Shape& dy = ctx->InputShape("dy", 0);
  const Shape& b = ctx->InputShape("b", 0);
  CHECK_EQ_OR_RETURN(dy.NumAxes(), 2);  // NOLINT(maybe-need-error-msg)
  CHECK_EQ_OR_RETURN(b.NumAxes(), 2);   // NOLINT(maybe-need-error-msg)
  CHECK_EQ_OR_RETURN(dy.At(0), ctx->Attr<int64_t>("a_rows")) << "dy rows must equal a_rows";
  CHECK_EQ_OR_RETURN(b.At(0), ctx->Attr<int64_t>("a_cols")) << "b rows must equal a_cols";
  CHECK_EQ_OR_RETURN(dy.At(1), b.At(1)) << "dy and b must have the same width";
  ctx->SetOutputShape("dval", 0, Shape({nnz}));
  return Maybe<void>::Ok();
}
/*static*/ Maybe<void> SddmmCsrOp::InferPhysicalTensorDesc(user_op::InferContext* ctx) {
  return InferLogicalTensorDesc(ctx);
}
/*static*/ Maybe<void> SddmmCsrOp::InferDataType(user_op::InferContext* ctx) {
  JUST(CheckIndexTypes(ctx));
  CHECK_EQ_OR_RETURN(ctx->InputDType("dy", 0), ctx->InputDType("b", 0)) << "dy and b must share a dtype";
  ctx->SetOutputDType("dval", 0, ctx->InputDType("b", 0));
  return Maybe<void>::Ok();
}
/*static*/ Maybe<void> SddmmCsrOp::GetSbp(user_op::SbpContext* ctx) {
  // a split of the dense width makes dval a partial sum
  ctx->NewBuilder()
      .Broadcast(user_op::OpArg("a_crow", 0))
      .Broadcast(user_op::OpArg("a_col", 0))
      .Split(user_op::OpArg("dy", 0), 1)
      .Split(user_op::OpArg("b", 0), 1)
      .PartialSum(user_op::OpArg("dval", 0))
      .Build();
  ctx->NewBuilder().Broadcast(ctx->inputs()).Broadcast(ctx->outputs()).Build();
  return Maybe<void>::Ok();
}
/*static*/ Maybe<void> SddmmCsrOp::ModifyInputArg(const GetInputArgModifier& fn,
                                                  const user_op::UserOpConfWrapper&) {
  return NoGradForIndices(fn);
}

}  // namespace oneflow

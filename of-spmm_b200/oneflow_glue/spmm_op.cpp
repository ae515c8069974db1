// oneflow/user/ops/spmm_op.cpp — shape / dtype / SBP inference of spmm_csr, fused_spmm_csr_bias_act,
// spmm_csr_grad_b, sddmm_csr and csr_transpose_structure.  Written against the reference's user-op framework; conventions
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
  for (const char* name : {"t_crow", "t_col", "t_perm"}) {   // optional cached structure of A^T
    user_op::InputArgModifier* m = GetInputArgModifierFn(name, 0);
    if (m != nullptr) { m->set_requires_grad(false); }
  }
  return Maybe<void>::Ok();
}

// A split of a_col / a_val is not a row split of A, so the CSR structure is always broadcast; the
// dense side may be column-split (always legal for a row-wise linear map).  Both products are
// bilinear in (a_val, dense operand), which gives the two partial-sum signatures — the analogue of
// matmul's P(a)·B(b) -> P and B(a)·P(b) -> P (oneflow/user/ops/matmul_op.cpp:112-136).
// nnz-balanced row blocks live inside the library (of-spmm_b200/dist.py), outside SBP.
Maybe<void> BilinearSbp(user_op::SbpContext* ctx, const char* dense_in, const char* out) {
  auto structure = [&](user_op::SbpSignatureBuilder& b) -> user_op::SbpSignatureBuilder& {
    b.Broadcast(user_op::OpArg("a_crow", 0)).Broadcast(user_op::OpArg("a_col", 0));
    for (const auto& in : ctx->inputs())   // optional cached structure of A^T: index tensors, broadcast
      if (in.first == "t_crow" || in.first == "t_col" || in.first == "t_perm") b.Broadcast(user_op::OpArg(in.first, in.second));
    return b;
  };
  {  // dense width split
    auto b = ctx->NewBuilder();
    structure(b).Broadcast(user_op::OpArg("a_val", 0)).Split(user_op::OpArg(dense_in, 0), 1).Split(user_op::OpArg(out, 0), 1).Build();
  }
  {  // partial sums of the edge values
    auto b = ctx->NewBuilder();
    structure(b).PartialSum(user_op::OpArg("a_val", 0)).Broadcast(user_op::OpArg(dense_in, 0)).PartialSum(user_op::OpArg(out, 0)).Build();
  }
  {  // partial sums of the dense operand
    auto b = ctx->NewBuilder();
    structure(b).Broadcast(user_op::OpArg("a_val", 0)).PartialSum(user_op::OpArg(dense_in, 0)).PartialSum(user_op::OpArg(out, 0)).Build();
  }
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
  return BilinearSbp(ctx, "b", "out");
}
/*static*/ Maybe<void> SpmmCsrOp::ModifyInputArg(const GetInputArgModifier& fn,
                                                 const user_op::UserOpConfWrapper&) {
  return NoGradForIndices(fn);
}

// ---------------------------------------------------------------- fused_spmm_csr_bias_act
// out = act(A·b + bias); same operand checks as spmm_csr plus the bias row.
/*static*/ Maybe<void> FusedSpmmCsrBiasActOp::InferLogicalTensorDesc(user_op::InferContext* ctx) {
  JUST(SpmmCsrOp::InferLogicalTensorDesc(ctx));
  const Shape& bias = ctx->InputShape("bias", 0);
  CHECK_EQ_OR_RETURN(bias.NumAxes(), 1) << "bias must be 1-D";
  CHECK_EQ_OR_RETURN(bias.At(0), ctx->InputShape("b", 0).At(1)) << "bias must have one entry per column of b";
  return Maybe<void>::Ok();
}
/*static*/ Maybe<void> FusedSpmmCsrBiasActOp::InferPhysicalTensorDesc(user_op::InferContext* ctx) {
  return InferLogicalTensorDesc(ctx);
}
/*static*/ Maybe<void> FusedSpmmCsrBiasActOp::InferDataType(user_op::InferContext* ctx) {
  JUST(SpmmCsrOp::InferDataType(ctx));
  CHECK_EQ_OR_RETURN(ctx->InputDType("bias", 0), ctx->InputDType("b", 0)) << "bias and b must share a dtype";
  return Maybe<void>::Ok();
}
/*static*/ Maybe<void> FusedSpmmCsrBiasActOp::GetSbp(user_op::SbpContext* ctx) {
  // bias add and ReLU are not linear, so no partial-sum signature survives: only the split of the
  // dense width (bias split along its only axis) and all-broadcast
  ctx->NewBuilder()
      .Broadcast(user_op::OpArg("a_crow", 0)).Broadcast(user_op::OpArg("a_col", 0)).Broadcast(user_op::OpArg("a_val", 0))
      .Split(user_op::OpArg("b", 0), 1).Split(user_op::OpArg("bias", 0), 0).Split(user_op::OpArg("out", 0), 1)
      .Build();
  ctx->NewBuilder().Broadcast(ctx->inputs()).Broadcast(ctx->outputs()).Build();
  return Maybe<void>::Ok();
}
/*static*/ Maybe<void> FusedSpmmCsrBiasActOp::ModifyInputArg(const GetInputArgModifier& fn,
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
  if (ctx->has_input("t_crow", 0)) {  // cached structure of A^T (csr_transpose_structure): all three or none
    CHECK_OR_RETURN(ctx->has_input("t_col", 0) && ctx->has_input("t_perm", 0)) << "t_crow, t_col and t_perm come together";
    CHECK_EQ_OR_RETURN(ctx->InputShape("t_crow", 0).elem_cnt(), ctx->Attr<int64_t>("a_cols") + 1)
        << "t_crow must have a_cols+1 entries";
    CHECK_EQ_OR_RETURN(ctx->InputShape("t_col", 0).elem_cnt(), nnz) << "t_col must have nnz entries";
    CHECK_EQ_OR_RETURN(ctx->InputShape("t_perm", 0).elem_cnt(), nnz) << "t_perm must have nnz entries";
  }
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
  return BilinearSbp(ctx, "dy", "db");
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
  // dval has the dtype of the values it is the gradient of (attr val_dtype, set by the grad
  // function from a_val) — with fp32 values and a bf16 dense operand that is fp32, not b's dtype
  const DataType val_dtype = ctx->Attr<DataType>("val_dtype");
  const DataType dense = ctx->InputDType("b", 0);
  CHECK_OR_RETURN(val_dtype == DataType::kInvalidDataType || val_dtype == DataType::kFloat || val_dtype == dense)
      << "val_dtype must be float32 or match b";
  ctx->SetOutputDType("dval", 0, val_dtype == DataType::kInvalidDataType ? dense : val_dtype);
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
  // linear in dy and in b: partial sums pass through
  ctx->NewBuilder().Broadcast(user_op::OpArg("a_crow", 0)).Broadcast(user_op::OpArg("a_col", 0))
      .PartialSum(user_op::OpArg("dy", 0)).Broadcast(user_op::OpArg("b", 0)).PartialSum(user_op::OpArg("dval", 0)).Build();
  ctx->NewBuilder().Broadcast(user_op::OpArg("a_crow", 0)).Broadcast(user_op::OpArg("a_col", 0))
      .Broadcast(user_op::OpArg("dy", 0)).PartialSum(user_op::OpArg("b", 0)).PartialSum(user_op::OpArg("dval", 0)).Build();
  ctx->NewBuilder().Broadcast(ctx->inputs()).Broadcast(ctx->outputs()).Build();
  return Maybe<void>::Ok();
}
/*static*/ Maybe<void> SddmmCsrOp::ModifyInputArg(const GetInputArgModifier& fn,
                                                  const user_op::UserOpConfWrapper&) {
  return NoGradForIndices(fn);
}

// ---------------------------------------------------------------- csr_transpose_structure
/*static*/ Maybe<void> CsrTransposeStructureOp::InferLogicalTensorDesc(user_op::InferContext* ctx) {
  int64_t nnz = 0;
  JUST(CheckCsr(ctx, &nnz));
  ctx->SetOutputShape("t_crow", 0, Shape({ctx->Attr<int64_t>("a_cols") + 1}));
  ctx->SetOutputShape("t_col", 0, Shape({nnz}));
  ctx->SetOutputShape("t_perm", 0, Shape({nnz}));
  return Maybe<void>::Ok();
}
/*static*/ Maybe<void> CsrTransposeStructureOp::InferPhysicalTensorDesc(user_op::InferContext* ctx) {
  return InferLogicalTensorDesc(ctx);
}
/*static*/ Maybe<void> CsrTransposeStructureOp::InferDataType(user_op::InferContext* ctx) {
  JUST(CheckIndexTypes(ctx));
  for (const char* out : {"t_crow", "t_col", "t_perm"}) { ctx->SetOutputDType(out, 0, ctx->InputDType("a_crow", 0)); }
  return Maybe<void>::Ok();
}
/*static*/ Maybe<void> CsrTransposeStructureOp::GetSbp(user_op::SbpContext* ctx) {
  ctx->NewBuilder().Broadcast(ctx->inputs()).Broadcast(ctx->outputs()).Build();   // index work: replicated
  return Maybe<void>::Ok();
}
/*static*/ Maybe<void> CsrTransposeStructureOp::ModifyInputArg(const GetInputArgModifier& fn,
                                                               const user_op::UserOpConfWrapper&) {
  return NoGradForIndices(fn);
}

}  // namespace oneflow

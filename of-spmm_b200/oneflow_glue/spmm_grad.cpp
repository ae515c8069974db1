// oneflow/core/autograd/gradient_funcs/spmm_csr.cpp — OpExprGradFunction of spmm_csr
// (SURVEY.md §8 a9).  Pattern: gradient_funcs/gather.cpp:29-72 and matmul.cpp:36-104; interface
// oneflow/core/framework/op_expr_grad_function.h:84-140.  Python mirror:
// of-spmm_b200/functional.py:_SpmmCsrFn.
#include "oneflow/core/framework/op_expr_grad_function.h"
#include "oneflow/core/functional/functional.h"

namespace oneflow {
namespace one {

struct SpmmCsrCaptureState : public AutoGradCaptureState {
  bool val_requires_grad = false;
  bool b_requires_grad = false;
  int64_t a_rows = 0;
  int64_t a_cols = 0;
  size_t crow_index = 0, col_index = 0, val_index = 0, b_index = 0;
  DataType val_dtype = DataType::kInvalidDataType;   // the SDDMM gradient takes the dtype of a_val
  bool has_at = false;                               // optional cached structure of A^T (inputs 4..6)
  size_t t_crow_index = 0, t_col_index = 0, t_perm_index = 0;
};

class SpmmCsr : public OpExprGradFunction<SpmmCsrCaptureState> {
 public:
  Maybe<void> Init(const OpExpr& op) override {
    const UserOpExpr* fw_op_expr = dynamic_cast<const UserOpExpr*>(&op);
    CHECK_NOTNULL_OR_RETURN(fw_op_expr);  // NOLINT(maybe-need-error-msg)
    base_attrs_ = MakeAttrMapFromUserOpConf(fw_op_expr->proto());
    return Maybe<void>::Ok();
  }

  // inputs: a_crow, a_col, a_val, b [, t_crow, t_col, t_perm] — index inputs never receive a gradient
  Maybe<void> Capture(SpmmCsrCaptureState* ctx, const TensorTuple& inputs, const TensorTuple& outputs,
                      const AttrMap& attrs) const override {
    ctx->val_requires_grad = inputs.at(2)->requires_grad();
    ctx->b_requires_grad = inputs.at(3)->requires_grad();
    if (!ctx->val_requires_grad && !ctx->b_requires_grad) { return Maybe<void>::Ok(); }
    ComposedAttrMap composed_attrs(attrs, base_attrs_);
    ctx->a_rows = JUST(composed_attrs.GetAttr<int64_t>("a_rows"));
    ctx->a_cols = JUST(composed_attrs.GetAttr<int64_t>("a_cols"));
    ctx->crow_index = ctx->SaveTensorForBackward(inputs.at(0));
    ctx->col_index = ctx->SaveTensorForBackward(inputs.at(1));
    ctx->val_dtype = inputs.at(2)->dtype();
    if (ctx->b_requires_grad) { ctx->val_index = ctx->SaveTensorForBackward(inputs.at(2)); }
    if (ctx->val_requires_grad) { ctx->b_index = ctx->SaveTensorForBackward(inputs.at(3)); }
    ctx->has_at = inputs.size() == 7 && ctx->b_requires_grad;
    if (ctx->has_at) {
      ctx->t_crow_index = ctx->SaveTensorForBackward(inputs.at(4));
      ctx->t_col_index = ctx->SaveTensorForBackward(inputs.at(5));
      ctx->t_perm_index = ctx->SaveTensorForBackward(inputs.at(6));
    }
    return Maybe<void>::Ok();
  }

  Maybe<void> Apply(const SpmmCsrCaptureState* ctx, const TensorTuple& out_grads,
                    TensorTuple* in_grads) const override {
    if (!ctx->val_requires_grad && !ctx->b_requires_grad) { return Maybe<void>::Ok(); }
    CHECK_EQ_OR_RETURN(out_grads.size(), 1);  // NOLINT(maybe-need-error-msg)
    in_grads->resize(ctx->has_at ? 7 : 4);
    const auto& crow = ctx->SavedTensors().at(ctx->crow_index);
    const auto& col = ctx->SavedTensors().at(ctx->col_index);
    if (ctx->val_requires_grad) {  // dval[p] = <dy[i,:], b[col[p],:]>
      const auto& b = ctx->SavedTensors().at(ctx->b_index);
      in_grads->at(2) = JUST(functional::SddmmCsr(crow, col, out_grads.at(0), b, ctx->a_rows, ctx->a_cols, ctx->val_dtype));
    }
    if (ctx->b_requires_grad) {    // db = A^T · dy
      const auto& val = ctx->SavedTensors().at(ctx->val_index);
      // with the cached structure of A^T: forward kernel on it, values re-gathered every call;
      // without: A^T built transiently inside the op's tmp_buffer.  Both deterministic.
      Optional<Tensor> t_crow, t_col, t_perm;
      if (ctx->has_at) {
        t_crow = ctx->SavedTensors().at(ctx->t_crow_index);
        t_col = ctx->SavedTensors().at(ctx->t_col_index);
        t_perm = ctx->SavedTensors().at(ctx->t_perm_index);
      }
      in_grads->at(3) = JUST(functional::SpmmCsrGradB(crow, col, val, out_grads.at(0), ctx->a_rows, ctx->a_cols, t_crow,
                                                      t_col, t_perm, /*atomic=*/false));
    }
    return Maybe<void>::Ok();
  }

 private:
  AttrMap base_attrs_;
};

REGISTER_OP_EXPR_GRAD_FUNCTION("spmm_csr", SpmmCsr);

// ---------------------------------------------------------------- fused_spmm_csr_bias_act
// out = act(A·b + bias).  With dz = dy ⊙ [out > 0] (ReLU; the mask comes from the saved OUTPUT, as in
// gradient_funcs/activation.cpp:205 — the pre-activation is never materialised) or dz = dy:
//   dbias = Σ_rows dz      (gradient_funcs/bias_add.cpp:62: ReduceSum over the non-bias axes)
//   dval  = sddmm_csr(dz, b)      db = A^T·dz
// Python mirror: of-spmm_b200/functional.py:_SpmmCsrEpilogueFn.
struct FusedSpmmCsrBiasActCaptureState : public AutoGradCaptureState {
  bool val_requires_grad = false, b_requires_grad = false, bias_requires_grad = false;
  bool relu = false;
  int64_t a_rows = 0, a_cols = 0;
  size_t crow_index = 0, col_index = 0, val_index = 0, b_index = 0, out_index = 0;
  DataType val_dtype = DataType::kInvalidDataType;
};

class FusedSpmmCsrBiasAct : public OpExprGradFunction<FusedSpmmCsrBiasActCaptureState> {
 public:
  Maybe<void> Init(const OpExpr& op) override {
    const UserOpExpr* fw_op_expr = dynamic_cast<const UserOpExpr*>(&op);
    CHECK_NOTNULL_OR_RETURN(fw_op_expr);  // NOLINT(maybe-need-error-msg)
    base_attrs_ = MakeAttrMapFromUserOpConf(fw_op_expr->proto());
    return Maybe<void>::Ok();
  }

  // inputs: a_crow, a_col, a_val, b, bias
  Maybe<void> Capture(FusedSpmmCsrBiasActCaptureState* ctx, const TensorTuple& inputs, const TensorTuple& outputs,
                      const AttrMap& attrs) const override {
    CHECK_EQ_OR_RETURN(inputs.size(), 5);  // NOLINT(maybe-need-error-msg)
    ctx->val_requires_grad = inputs.at(2)->requires_grad();
    ctx->b_requires_grad = inputs.at(3)->requires_grad();
    ctx->bias_requires_grad = inputs.at(4)->requires_grad();
    if (!ctx->val_requires_grad && !ctx->b_requires_grad && !ctx->bias_requires_grad) { return Maybe<void>::Ok(); }
    ComposedAttrMap composed_attrs(attrs, base_attrs_);
    ctx->a_rows = JUST(composed_attrs.GetAttr<int64_t>("a_rows"));
    ctx->a_cols = JUST(composed_attrs.GetAttr<int64_t>("a_cols"));
    ctx->relu = JUST(composed_attrs.GetAttr<bool>("relu"));
    ctx->val_dtype = inputs.at(2)->dtype();
    if (ctx->relu) { ctx->out_index = ctx->SaveTensorForBackward(outputs.at(0)); }
    if (ctx->val_requires_grad || ctx->b_requires_grad) {
      ctx->crow_index = ctx->SaveTensorForBackward(inputs.at(0));
      ctx->col_index = ctx->SaveTensorForBackward(inputs.at(1));
    }
    if (ctx->b_requires_grad) { ctx->val_index = ctx->SaveTensorForBackward(inputs.at(2)); }
    if (ctx->val_requires_grad) { ctx->b_index = ctx->SaveTensorForBackward(inputs.at(3)); }
    return Maybe<void>::Ok();
  }

  Maybe<void> Apply(const FusedSpmmCsrBiasActCaptureState* ctx, const TensorTuple& out_grads,
                    TensorTuple* in_grads) const override {
    if (!ctx->val_requires_grad && !ctx->b_requires_grad && !ctx->bias_requires_grad) { return Maybe<void>::Ok(); }
    CHECK_EQ_OR_RETURN(out_grads.size(), 1);  // NOLINT(maybe-need-error-msg)
    in_grads->resize(5);
    std::shared_ptr<Tensor> dz = out_grads.at(0);
    if (ctx->relu) { dz = JUST(functional::ReluGrad(out_grads.at(0), ctx->SavedTensors().at(ctx->out_index))); }
    if (ctx->bias_requires_grad) { in_grads->at(4) = JUST(functional::ReduceSum(dz, std::vector<int32_t>{0}, false)); }
    if (ctx->val_requires_grad) {
      in_grads->at(2) = JUST(functional::SddmmCsr(ctx->SavedTensors().at(ctx->crow_index), ctx->SavedTensors().at(ctx->col_index),
                                                  dz, ctx->SavedTensors().at(ctx->b_index), ctx->a_rows, ctx->a_cols,
                                                  ctx->val_dtype));
    }
    if (ctx->b_requires_grad) {
      in_grads->at(3) = JUST(functional::SpmmCsrGradB(ctx->SavedTensors().at(ctx->crow_index),
                                                      ctx->SavedTensors().at(ctx->col_index),
                                                      ctx->SavedTensors().at(ctx->val_index), dz, ctx->a_rows, ctx->a_cols,
                                                      Optional<Tensor>(), Optional<Tensor>(), Optional<Tensor>(),
                                                      /*atomic=*/false));
    }
    return Maybe<void>::Ok();
  }

 private:
  AttrMap base_attrs_;
};

REGISTER_OP_EXPR_GRAD_FUNCTION("fused_spmm_csr_bias_act", FusedSpmmCsrBiasAct);

}  // namespace one
}  // namespace oneflow

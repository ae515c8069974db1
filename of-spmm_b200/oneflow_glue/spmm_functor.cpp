// Functors for oneflow/core/functional/impl/nn_functor.cpp.  Pattern:
// GatherFunctor (oneflow/core/functional/impl/array_functor.cpp:919-933) and MatMulFunctor
// (impl/nn_functor.cpp:290-323): one OpExpr per input arity built once, attrs through the
// thread-cached map; optional tensor arguments as in the conv / batch-norm functors
// (impl/nn_functor.cpp:89,1041).  Add to functional_api.yaml (see functional_api.yaml.patch) and
// register below.
#include "oneflow/core/functional/function_library.h"
#include "oneflow/core/functional/impl/common.h"
#include "oneflow/core/framework/op_builder.h"
#include "oneflow/core/framework/op_interpreter/op_interpreter_util.h"

namespace oneflow {
namespace one {
namespace functional {
namespace impl {

// flow._C.spmm_csr(a_crow, a_col, a_val, b, a_rows, a_cols, t_crow=None, t_col=None, t_perm=None).
// The optional structure of A^T (from flow._C.csr_transpose_structure, computed once for a static
// graph) is not used by the forward; it rides along as extra inputs so that the grad function can
// hand it to spmm_csr_grad_b — ordinary tensors owned by the caller, no hidden cache.
class SpmmCsrFunctor {
 public:
  SpmmCsrFunctor() {
    op_ = CHECK_JUST(one::OpBuilder("spmm_csr").Input("a_crow").Input("a_col").Input("a_val").Input("b")
                         .Output("out").Build());
    op_with_at_ = CHECK_JUST(one::OpBuilder("spmm_csr").Input("a_crow").Input("a_col").Input("a_val").Input("b")
                                 .Input("t_crow").Input("t_col").Input("t_perm").Output("out").Build());
  }
  Maybe<Tensor> operator()(const std::shared_ptr<one::Tensor>& a_crow,
                           const std::shared_ptr<one::Tensor>& a_col,
                           const std::shared_ptr<one::Tensor>& a_val,
                           const std::shared_ptr<one::Tensor>& b, const int64_t& a_rows,
                           const int64_t& a_cols, const Optional<one::Tensor>& t_crow,
                           const Optional<one::Tensor>& t_col, const Optional<one::Tensor>& t_perm) const {
    CHECK_EQ_OR_RETURN(b->ndim(), 2) << Error::RuntimeError() << "b must be 2-D, got " << b->ndim() << "-D";
    CHECK_EQ_OR_RETURN(b->dim(0), a_cols) << Error::RuntimeError() << "b has " << b->dim(0)
                                          << " rows but a_cols = " << a_cols;
    auto& attrs = THREAD_CACHED_MUTABLE_ATTR_MAP("a_rows", "a_cols");
    attrs.SetAllAttrs(a_rows, a_cols);
    if (t_crow) {
      CHECK_OR_RETURN(t_col && t_perm) << Error::RuntimeError() << "t_crow, t_col and t_perm come together";
      return OpInterpUtil::Dispatch<Tensor>(*op_with_at_, {a_crow, a_col, a_val, b, JUST(t_crow), JUST(t_col), JUST(t_perm)}, attrs);
    }
    return OpInterpUtil::Dispatch<Tensor>(*op_, {a_crow, a_col, a_val, b}, attrs);
  }

 private:
  std::shared_ptr<OpExpr> op_, op_with_at_;
};

// flow._C.fused_spmm_csr_bias_act(a_crow, a_col, a_val, b, bias, a_rows, a_cols, relu=False)
class FusedSpmmCsrBiasActFunctor {
 public:
  FusedSpmmCsrBiasActFunctor() {
    op_ = CHECK_JUST(one::OpBuilder("fused_spmm_csr_bias_act").Input("a_crow").Input("a_col").Input("a_val").Input("b")
                         .Input("bias").Output("out").Build());
  }
  Maybe<Tensor> operator()(const std::shared_ptr<one::Tensor>& a_crow,
                           const std::shared_ptr<one::Tensor>& a_col,
                           const std::shared_ptr<one::Tensor>& a_val,
                           const std::shared_ptr<one::Tensor>& b,
                           const std::shared_ptr<one::Tensor>& bias, const int64_t& a_rows,
                           const int64_t& a_cols, const bool& relu) const {
    CHECK_EQ_OR_RETURN(b->ndim(), 2) << Error::RuntimeError() << "b must be 2-D, got " << b->ndim() << "-D";
    CHECK_EQ_OR_RETURN(b->dim(0), a_cols) << Error::RuntimeError() << "b has " << b->dim(0)
                                          << " rows but a_cols = " << a_cols;
    CHECK_EQ_OR_RETURN(bias->ndim(), 1) << Error::RuntimeError() << "bias must be 1-D, got " << bias->ndim() << "-D";
    CHECK_EQ_OR_RETURN(bias->dim(0), b->dim(1)) << Error::RuntimeError() << "bias has " << bias->dim(0)
                                                << " entries but b has " << b->dim(1) << " columns";
    auto& attrs = THREAD_CACHED_MUTABLE_ATTR_MAP("a_rows", "a_cols", "relu");
    attrs.SetAllAttrs(a_rows, a_cols, relu);
    return OpInterpUtil::Dispatch<Tensor>(*op_, {a_crow, a_col, a_val, b, bias}, attrs);
  }

 private:
  std::shared_ptr<OpExpr> op_;
};

class SpmmCsrGradBFunctor {
 public:
  SpmmCsrGradBFunctor() {
    op_ = CHECK_JUST(one::OpBuilder("spmm_csr_grad_b").Input("a_crow").Input("a_col").Input("a_val")
                         .Input("dy").Output("db").Build());
    op_with_at_ = CHECK_JUST(one::OpBuilder("spmm_csr_grad_b").Input("a_crow").Input("a_col").Input("a_val").Input("dy")
                                 .Input("t_crow").Input("t_col").Input("t_perm").Output("db").Build());
  }
  Maybe<Tensor> operator()(const std::shared_ptr<one::Tensor>& a_crow,
                           const std::shared_ptr<one::Tensor>& a_col,
                           const std::shared_ptr<one::Tensor>& a_val,
                           const std::shared_ptr<one::Tensor>& dy, const int64_t& a_rows,
                           const int64_t& a_cols, const Optional<one::Tensor>& t_crow,
                           const Optional<one::Tensor>& t_col, const Optional<one::Tensor>& t_perm,
                           const bool& atomic) const {
    auto& attrs = THREAD_CACHED_MUTABLE_ATTR_MAP("a_rows", "a_cols", "atomic");
    attrs.SetAllAttrs(a_rows, a_cols, atomic);
    if (t_crow) {
      CHECK_OR_RETURN(t_col && t_perm) << Error::RuntimeError() << "t_crow, t_col and t_perm come together";
      return OpInterpUtil::Dispatch<Tensor>(*op_with_at_, {a_crow, a_col, a_val, dy, JUST(t_crow), JUST(t_col), JUST(t_perm)}, attrs);
    }
    return OpInterpUtil::Dispatch<Tensor>(*op_, {a_crow, a_col, a_val, dy}, attrs);
  }

 private:
  std::shared_ptr<OpExpr> op_, op_with_at_;
};

class SddmmCsrFunctor {
 public:
  SddmmCsrFunctor() {
    op_ = CHECK_JUST(one::OpBuilder("sddmm_csr").Input("a_crow").Input("a_col").Input("dy").Input("b")
                         .Output("dval").Build());
  }
  // val_dtype: dtype of the values this is the gradient of (fp32 values with a bf16 dense operand
  // get an fp32 gradient); kInvalidDataType = the dense dtype
  Maybe<Tensor> operator()(const std::shared_ptr<one::Tensor>& a_crow,
                           const std::shared_ptr<one::Tensor>& a_col,
                           const std::shared_ptr<one::Tensor>& dy,
                           const std::shared_ptr<one::Tensor>& b, const int64_t& a_rows,
                           const int64_t& a_cols, const DataType& val_dtype) const {
    auto& attrs = THREAD_CACHED_MUTABLE_ATTR_MAP("a_rows", "a_cols", "val_dtype");
    attrs.SetAllAttrs(a_rows, a_cols, val_dtype);
    return OpInterpUtil::Dispatch<Tensor>(*op_, {a_crow, a_col, dy, b}, attrs);
  }

 private:
  std::shared_ptr<OpExpr> op_;
};

// flow._C.csr_transpose_structure(a_crow, a_col, a_rows, a_cols) -> (t_crow, t_col, t_perm)
class CsrTransposeStructureFunctor {
 public:
  CsrTransposeStructureFunctor() {
    op_ = CHECK_JUST(one::OpBuilder("csr_transpose_structure").Input("a_crow").Input("a_col")
                         .Output("t_crow").Output("t_col").Output("t_perm").Build());
  }
  Maybe<TensorTuple> operator()(const std::shared_ptr<one::Tensor>& a_crow,
                                const std::shared_ptr<one::Tensor>& a_col, const int64_t& a_rows,
                                const int64_t& a_cols) const {
    auto& attrs = THREAD_CACHED_MUTABLE_ATTR_MAP("a_rows", "a_cols");
    attrs.SetAllAttrs(a_rows, a_cols);
    return OpInterpUtil::Dispatch<TensorTuple>(*op_, {a_crow, a_col}, attrs);
  }

 private:
  std::shared_ptr<OpExpr> op_;
};

}  // namespace impl

ONEFLOW_FUNCTION_LIBRARY(m) {
  m.add_functor<impl::SpmmCsrFunctor>("SpmmCsr");
  m.add_functor<impl::FusedSpmmCsrBiasActFunctor>("FusedSpmmCsrBiasAct");
  m.add_functor<impl::SpmmCsrGradBFunctor>("SpmmCsrGradB");
  m.add_functor<impl::SddmmCsrFunctor>("SddmmCsr");
  m.add_functor<impl::CsrTransposeStructureFunctor>("CsrTransposeStructure");
}

}  // namespace functional
}  // namespace one
}  // namespace oneflow

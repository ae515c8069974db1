// Functors for oneflow/core/functional/impl/nn_functor.cpp.  Pattern:
// GatherFunctor (oneflow/core/functional/impl/array_functor.cpp:919-933) and MatMulFunctor
// (impl/nn_functor.cpp:290-323): one OpExpr built once, attrs through the thread-cached map.
// Add to functional_api.yaml (see functional_api.yaml.patch) and register below.
#include "oneflow/core/functional/function_library.h"
#include "oneflow/core/functional/impl/common.h"
#include "oneflow/core/framework/op_builder.h"
#include "oneflow/core/framework/op_interpreter/op_interpreter_util.h"

namespace oneflow {
namespace one {
namespace functional {
namespace impl {

class SpmmCsrFunctor {
 public:
  SpmmCsrFunctor() {
    op_ = CHECK_JUST(one::OpBuilder("spmm_csr").Input("a_crow").Input("a_col").Input("a_val").Input("b")
                         .Output("out").Build());
  }
  Maybe<Tensor> operator()(const std::shared_ptr<one::Tensor>& a_crow,
                           const std::shared_ptr<one::Tensor>& a_col,
                           const std::shared_ptr<one::Tensor>& a_val,
                           const std::shared_ptr<one::Tensor>& b, const int64_t& a_rows,
                           const int64_t& a_cols) const {
    CHECK_EQ_OR_RETURN(b->ndim(), 2) << Error::RuntimeError() << "b must be 2-D, got " << b->ndim() << "-D";
    CHECK_EQ_OR_RETURN(b->dim(0), a_cols) << Error::RuntimeError() << "b has " << b->dim(0)
                                          << " rows but a_cols = " << a_cols;
    auto& attrs = THREAD_CACHED_MUTABLE_ATTR_MAP("a_rows", "a_cols");
    attrs.SetAllAttrs(a_rows, a_cols);
    return OpInterpUtil::Dispatch<Tensor>(*op_, {a_crow, a_col, a_val, b}, attrs);
  }

 private:
  std::shared_ptr<OpExpr> op_;
};

class SpmmCsrGradBFunctor {
 public:
  SpmmCsrGradBFunctor() {
    op_ = CHECK_JUST(one::OpBuilder("spmm_csr_grad_b").Input("a_crow").Input("a_col").Input("a_val")
                         .Input("dy").Output("db").Build());
  }
  Maybe<Tensor> operator()(const std::shared_ptr<one::Tensor>& a_crow,
                           const std::shared_ptr<one::Tensor>& a_col,
                           const std::shared_ptr<one::Tensor>& a_val,
                           const std::shared_ptr<one::Tensor>& dy, const int64_t& a_rows,
                           const int64_t& a_cols) const {
    auto& attrs = THREAD_CACHED_MUTABLE_ATTR_MAP("a_rows", "a_cols");
    attrs.SetAllAttrs(a_rows, a_cols);
    return OpInterpUtil::Dispatch<Tensor>(*op_, {a_crow, a_col, a_val, dy}, attrs);
  }

 private:
  std::shared_ptr<OpExpr> op_;
};

class SddmmCsrFunctor {
 public:
  SddmmCsrFunctor() {
    op_ = CHECK_JUST(one::OpBuilder("sddmm_csr").Input("a_crow").Input("a_col").Input("dy").Input("b")
                         .Output("dval").Build());
  }
  Maybe<Tensor> operator()(const std::shared_ptr<one::Tensor>& a_crow,
                           const std::shared_ptr<one::Tensor>& a_col,
                           const std::shared_ptr<one::Tensor>& dy,
                           const std::shared_ptr<one::Tensor>& b, const int64_t& a_rows,
                           const int64_t& a_cols) const {
    auto& attrs = THREAD_CACHED_MUTABLE_ATTR_MAP("a_rows", "a_cols");
    attrs.SetAllAttrs(a_rows, a_cols);
    return OpInterpUtil::Dispatch<Tensor>(*op_, {a_crow, a_col, dy, b}, attrs);
  }

 private:
  std::shared_ptr<OpExpr> op_;
};

}  // namespace impl

ONEFLOW_FUNCTION_LIBRARY(m) {
  m.add_functor<impl::SpmmCsrFunctor>("SpmmCsr");
  m.add_functor<impl::SpmmCsrGradBFunctor>("SpmmCsrGradB");
  m.add_functor<impl::SddmmCsrFunctor>("SddmmCsr");
}

}  // namespace functional
}  // namespace one
}  // namespace oneflow

// Stand-ins for the functional / autograd slice used by oneflow_glue/spmm_functor.cpp and
// spmm_grad.cpp.  TEST INFRASTRUCTURE; each declaration cites the reference header it mirrors.
#pragma once
#include <any>
#include <tuple>

#include "oneflow/core/framework/framework.h"

namespace oneflow {

// oneflow/core/framework/attr_map.h, mutable_attr_map.h:60-80
class AttrMap {
 public:
  std::map<std::string, int64_t> ints;
};
class MutableAttrMap : public AttrMap {
 public:
  explicit MutableAttrMap(std::vector<std::string> names) : names_(std::move(names)) {}
  template <typename... Args> void SetAllAttrs(Args&&... args) {
    if (sizeof...(args) != names_.size()) mock::Fatal("SetAllAttrs: wrong number of attrs");
    size_t i = 0;
    (void)std::initializer_list<int>{(ints[names_[i++]] = static_cast<int64_t>(args), 0)...};
  }
 private:
  std::vector<std::string> names_;
};
#define THREAD_CACHED_MUTABLE_ATTR_MAP(...) \
  (*[] { static thread_local ::oneflow::MutableAttrMap m({__VA_ARGS__}); return &m; }())
class ComposedAttrMap {  // oneflow/core/framework/attr_map.h (prior attrs shadow base attrs)
 public:
  ComposedAttrMap(const AttrMap& prior, const AttrMap& base) : prior_(prior), base_(base) {}
  template <typename T> Maybe<T> GetAttr(const std::string& name) const {
    auto it = prior_.ints.find(name);
    if (it != prior_.ints.end()) return static_cast<T>(it->second);
    it = base_.ints.find(name);
    if (it != base_.ints.end()) return static_cast<T>(it->second);
    return mock::ErrorCarrier{"attr not found: " + name};
  }
 private:
  const AttrMap& prior_;
  const AttrMap& base_;
};

namespace one {

// oneflow/core/framework/tensor.h
// oneflow/core/common/optional.h (the slice the functors use): an optional tensor argument is an
// Optional<Tensor>; `if (opt)` tests it and JUST(opt) yields the std::shared_ptr<Tensor>.
template <typename T> class Optional {
 public:
  Optional() = default;
  Optional(std::shared_ptr<T> v) : v_(std::move(v)) {}  // NOLINT
  bool has_value() const { return v_ != nullptr; }
  explicit operator bool() const { return has_value(); }
  bool IsOk() const { return has_value(); }
  std::string msg() const { return "Optional has no value"; }
  std::shared_ptr<T> Value() const { return v_; }
 private:
  std::shared_ptr<T> v_;
};

class Tensor {
 public:
  Tensor(std::string name, std::vector<int64_t> dims, bool requires_grad, DataType dtype = kFloat)
      : name_(std::move(name)), dims_(std::move(dims)), requires_grad_(requires_grad), dtype_(dtype) {}
  bool requires_grad() const { return requires_grad_; }
  DataType dtype() const { return dtype_; }
  int64_t ndim() const { return static_cast<int64_t>(dims_.size()); }
  int64_t dim(int64_t i) const { return dims_.at(i); }
  const std::string& name() const { return name_; }
 private:
  std::string name_;
  std::vector<int64_t> dims_;
  bool requires_grad_;
  DataType dtype_;
};
using TensorTuple = std::vector<std::shared_ptr<Tensor>>;

// oneflow/core/framework/op_expr.h, op_builder.h:26-60
struct UserOpConf { AttrMap attrs; };
class OpExpr {
 public:
  virtual ~OpExpr() = default;
  std::string op_type_name;
  std::vector<std::string> inputs, outputs;
};
class UserOpExpr : public OpExpr {
 public:
  const UserOpConf& proto() const { return proto_; }
  UserOpConf proto_;
};
inline AttrMap MakeAttrMapFromUserOpConf(const UserOpConf& c) { return c.attrs; }
class OpBuilder {
 public:
  explicit OpBuilder(const std::string& op) { e_ = std::make_shared<UserOpExpr>(); e_->op_type_name = op; }
  OpBuilder& Input(const std::string& n) { e_->inputs.push_back(n); return *this; }
  OpBuilder& Output(const std::string& n) { e_->outputs.push_back(n); return *this; }
  Maybe<UserOpExpr> Build() { return e_; }
 private:
  std::shared_ptr<UserOpExpr> e_;
};

// oneflow/core/framework/op_interpreter/op_interpreter_util.h:146-154
struct DispatchRecord { std::string op; std::vector<std::string> inputs; std::map<std::string, int64_t> attrs; };
std::vector<DispatchRecord>& DispatchLog();
struct OpInterpUtil {
  template <typename T>
  static Maybe<T> Dispatch(const OpExpr& op, const TensorTuple& inputs, const AttrMap& attrs) {
    DispatchRecord r;
    r.op = op.op_type_name;
    for (const auto& t : inputs) r.inputs.push_back(t->name());
    r.attrs = attrs.ints;
    DispatchLog().push_back(r);
    if constexpr (std::is_same<T, TensorTuple>::value) {
      auto outs = std::make_shared<TensorTuple>();
      for (const auto& o : op.outputs) outs->push_back(std::make_shared<Tensor>(op.op_type_name + ":" + o, std::vector<int64_t>{}, false));
      return outs;
    } else {
      return std::make_shared<Tensor>(op.op_type_name + ":" + op.outputs.at(0), std::vector<int64_t>{}, false);
    }
  }
};

// oneflow/core/framework/op_expr_grad_function.h:33-60,84-150,245-246
class AutoGradCaptureState {
 public:
  virtual ~AutoGradCaptureState() = default;
  const TensorTuple& SavedTensors() const { return saved_; }
  size_t SaveTensorForBackward(const std::shared_ptr<Tensor>& t) { saved_.push_back(t); return saved_.size() - 1; }
 private:
  TensorTuple saved_;
};
class OpExprGradFunctionIf {
 public:
  virtual ~OpExprGradFunctionIf() = default;
  virtual std::shared_ptr<AutoGradCaptureState> MakeCustomState() const = 0;
  virtual Maybe<void> Init(const OpExpr& op) = 0;
  virtual Maybe<void> CaptureIf(AutoGradCaptureState*, const TensorTuple&, const TensorTuple&, const AttrMap&) const = 0;
  virtual Maybe<void> ApplyIf(const AutoGradCaptureState*, const TensorTuple&, TensorTuple*) const = 0;
};
template <typename StateT>
class OpExprGradFunction : public OpExprGradFunctionIf {
 public:
  std::shared_ptr<AutoGradCaptureState> MakeCustomState() const override { return std::make_shared<StateT>(); }
  Maybe<void> CaptureIf(AutoGradCaptureState* ctx, const TensorTuple& in, const TensorTuple& out,
                        const AttrMap& attrs) const override {
    return Capture(dynamic_cast<StateT*>(ctx), in, out, attrs);
  }
  Maybe<void> ApplyIf(const AutoGradCaptureState* ctx, const TensorTuple& og, TensorTuple* ig) const override {
    return Apply(dynamic_cast<const StateT*>(ctx), og, ig);
  }
 protected:
  virtual Maybe<void> Capture(StateT*, const TensorTuple&, const TensorTuple&, const AttrMap&) const = 0;
  virtual Maybe<void> Apply(const StateT*, const TensorTuple&, TensorTuple*) const = 0;
};
std::map<std::string, std::function<OpExprGradFunctionIf*()>>& GradFunctionRegistry();
struct GradRegisterTrigger {
  GradRegisterTrigger(const std::string& op, std::function<OpExprGradFunctionIf*()> f) { GradFunctionRegistry()[op] = std::move(f); }
};
#define REGISTER_OP_EXPR_GRAD_FUNCTION(op_type, op_grad) \
  static ::oneflow::one::GradRegisterTrigger OF_MOCK_CAT(g_grad_trigger_, __COUNTER__)(op_type, [] { return static_cast<::oneflow::one::OpExprGradFunctionIf*>(new op_grad); })

// oneflow/core/functional/function_library.h + generated functional.h: functors are stored
// type-erased under their YAML names, each with the signature of its operator().
namespace functional {
template <typename T> struct FunctorSignature;
template <typename C, typename R, typename... A> struct FunctorSignature<R (C::*)(A...) const> {
  using type = std::function<R(A...)>;
  template <typename F> static type Wrap(std::shared_ptr<F> f) { return [f](A... a) { return (*f)(a...); }; }
};
std::map<std::string, std::any>& FunctionLibraryStore();
class FunctionLibrary {
 public:
  template <typename F> void add_functor(const std::string& name) {
    using Sig = FunctorSignature<decltype(&F::operator())>;
    FunctionLibraryStore()[name] = Sig::Wrap(std::make_shared<F>());
  }
};
struct FunctionLibraryTrigger { explicit FunctionLibraryTrigger(void (*fn)(FunctionLibrary&)) { FunctionLibrary m; fn(m); } };
#define ONEFLOW_FUNCTION_LIBRARY(m)                                                     \
  static void OF_MOCK_CAT(of_function_library_, __LINE__)(::oneflow::one::functional::FunctionLibrary&); \
  static ::oneflow::one::functional::FunctionLibraryTrigger OF_MOCK_CAT(g_fl_trigger_, __LINE__)(        \
      &OF_MOCK_CAT(of_function_library_, __LINE__));                                    \
  static void OF_MOCK_CAT(of_function_library_, __LINE__)(::oneflow::one::functional::FunctionLibrary& m)
// what functional_api.yaml.patch generates (tools/functional/*.py): free functions by YAML name,
// signatures exactly as in the YAML entries
using TensorPtr = std::shared_ptr<Tensor>;
template <typename R, typename... A> R CallFunctor(const char* name, A... a) {
  return std::any_cast<const std::function<R(A...)>&>(FunctionLibraryStore().at(name))(a...);
}
using OptTensor = Optional<Tensor>;
inline Maybe<Tensor> SpmmCsr(const TensorPtr& crow, const TensorPtr& col, const TensorPtr& val, const TensorPtr& b,
                             const int64_t& rows, const int64_t& cols, const OptTensor& t_crow = OptTensor(),
                             const OptTensor& t_col = OptTensor(), const OptTensor& t_perm = OptTensor()) {
  return CallFunctor<Maybe<Tensor>, const TensorPtr&, const TensorPtr&, const TensorPtr&, const TensorPtr&, const int64_t&,
                     const int64_t&, const OptTensor&, const OptTensor&, const OptTensor&>("SpmmCsr", crow, col, val, b, rows,
                                                                                          cols, t_crow, t_col, t_perm);
}
inline Maybe<Tensor> SpmmCsrGradB(const TensorPtr& crow, const TensorPtr& col, const TensorPtr& val, const TensorPtr& dy,
                                  const int64_t& rows, const int64_t& cols, const OptTensor& t_crow, const OptTensor& t_col,
                                  const OptTensor& t_perm, const bool& atomic) {
  return CallFunctor<Maybe<Tensor>, const TensorPtr&, const TensorPtr&, const TensorPtr&, const TensorPtr&, const int64_t&,
                     const int64_t&, const OptTensor&, const OptTensor&, const OptTensor&, const bool&>(
      "SpmmCsrGradB", crow, col, val, dy, rows, cols, t_crow, t_col, t_perm, atomic);
}
inline Maybe<Tensor> SddmmCsr(const TensorPtr& crow, const TensorPtr& col, const TensorPtr& dy, const TensorPtr& b,
                              const int64_t& rows, const int64_t& cols, const DataType& val_dtype) {
  return CallFunctor<Maybe<Tensor>, const TensorPtr&, const TensorPtr&, const TensorPtr&, const TensorPtr&, const int64_t&,
                     const int64_t&, const DataType&>("SddmmCsr", crow, col, dy, b, rows, cols, val_dtype);
}
inline Maybe<Tensor> FusedSpmmCsrBiasAct(const TensorPtr& crow, const TensorPtr& col, const TensorPtr& val, const TensorPtr& b,
                                         const TensorPtr& bias, const int64_t& rows, const int64_t& cols, const bool& relu = false) {
  return CallFunctor<Maybe<Tensor>, const TensorPtr&, const TensorPtr&, const TensorPtr&, const TensorPtr&, const TensorPtr&,
                     const int64_t&, const int64_t&, const bool&>("FusedSpmmCsrBiasAct", crow, col, val, b, bias, rows, cols, relu);
}
// two functions of the reference's own library the fused op's grad function calls
// (functional_api.yaml:295-298 "reduce_sum", :547-549 "relu_grad"); here they only log the dispatch
inline Maybe<Tensor> ReluGrad(const TensorPtr& dy, const TensorPtr& y) {
  DispatchLog().push_back(DispatchRecord{"relu_grad", {dy->name(), y->name()}, {}});
  return std::make_shared<Tensor>("relu_grad:dx", std::vector<int64_t>{}, false);
}
inline Maybe<Tensor> ReduceSum(const TensorPtr& x, const std::vector<int32_t>& axis, const bool& keepdims) {
  std::map<std::string, int64_t> attrs{{"keepdims", keepdims ? 1 : 0}};
  for (size_t i = 0; i < axis.size(); ++i) attrs["axis" + std::to_string(i)] = axis[i];
  DispatchLog().push_back(DispatchRecord{"reduce_sum", {x->name()}, attrs});
  return std::make_shared<Tensor>("reduce_sum:y", std::vector<int64_t>{}, false);
}
inline Maybe<TensorTuple> CsrTransposeStructure(const TensorPtr& crow, const TensorPtr& col, const int64_t& rows,
                                                const int64_t& cols) {
  return CallFunctor<Maybe<TensorTuple>, const TensorPtr&, const TensorPtr&, const int64_t&, const int64_t&>(
      "CsrTransposeStructure", crow, col, rows, cols);
}
}  // namespace functional
}  // namespace one
}  // namespace oneflow

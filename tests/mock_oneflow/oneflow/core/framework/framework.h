// Minimal stand-in for the slice of OneFlow's user-op framework that the glue sources in
// of-spmm_b200/oneflow_glue/ use.  TEST INFRASTRUCTURE: lets spmm_op.cpp / spmm_kernels.cpp be
// compiled and exercised here, where OneFlow itself cannot be built.  Every
// declaration mirrors the reference's signature; the citation says where the real one lives
// (paths relative to /root/reference).  Nothing here is copied: bodies are trivial containers.
#pragma once
#include <cstdint>
#include <functional>
#include <initializer_list>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace oneflow {

// oneflow/core/common/data_type.proto:4-17
enum DataType { kInvalidDataType = 0, kChar = 1, kFloat = 2, kDouble = 3, kInt8 = 4, kInt32 = 5, kInt64 = 6,
                kUInt8 = 7, kFloat16 = 9, kBFloat16 = 11 };
// oneflow/core/common/device_type.proto
enum class DeviceType { kInvalidDevice = 0, kCPU = 1, kCUDA = 2, kMockDevice = 3 };
inline bool IsIndexDataType(DataType t) { return t == kInt32 || t == kInt64; }  // data_type_seq.h:50-52

// ---- Maybe<T> / JUST / CHECK_*_OR_RETURN (oneflow/core/common/maybe.h:331, just.h:110-125).
// Like the reference: JUST(x) of a Maybe<class T> yields std::shared_ptr<T>, of a scalar yields T.
namespace mock {
struct ErrorCarrier { std::string msg; };
[[noreturn]] void Fatal(const std::string& msg);
}  // namespace mock
template <typename T, typename Enable = void> class Maybe;
template <> class Maybe<void, void> {
 public:
  Maybe() = default;
  Maybe(const mock::ErrorCarrier& e) : ok_(false), msg_(e.msg) {}  // NOLINT
  static Maybe Ok() { return Maybe(); }
  static Maybe Error(std::string msg) { return Maybe(mock::ErrorCarrier{std::move(msg)}); }
  bool IsOk() const { return ok_; }
  const std::string& msg() const { return msg_; }
  void Value() const {}
 private:
  bool ok_ = true;
  std::string msg_;
};
template <typename T> class Maybe<T, typename std::enable_if<std::is_scalar<T>::value>::type> {
 public:
  Maybe(T v) : v_(v) {}  // NOLINT
  Maybe(const mock::ErrorCarrier& e) : ok_(false), msg_(e.msg) {}  // NOLINT
  bool IsOk() const { return ok_; }
  const std::string& msg() const { return msg_; }
  T Value() const { return v_; }
 private:
  T v_{};
  bool ok_ = true;
  std::string msg_;
};
template <typename T> class Maybe<T, typename std::enable_if<std::is_class<T>::value>::type> {
 public:
  Maybe(std::shared_ptr<T> v) : v_(std::move(v)) {}  // NOLINT
  Maybe(const mock::ErrorCarrier& e) : ok_(false), msg_(e.msg) {}  // NOLINT
  bool IsOk() const { return ok_; }
  const std::string& msg() const { return msg_; }
  std::shared_ptr<T> Value() const { return v_; }
 private:
  std::shared_ptr<T> v_;
  bool ok_ = true;
  std::string msg_;
};
struct Error { static const char* RuntimeError() { return "RuntimeError: "; } };
namespace mock {
struct ErrorStream {  // collects `<< msg` and converts to a failed Maybe<T>
  explicit ErrorStream(const char* cond) { ss << "Check failed: " << cond << " "; }
  template <typename T> ErrorStream& operator<<(const T& v) { ss << v; return *this; }
  template <typename T> operator Maybe<T>() const { return Maybe<T>(ErrorCarrier{ss.str()}); }
  std::ostringstream ss;
};
struct FatalStream {  // glog-style fatal CHECK
  explicit FatalStream(const char* cond) { ss << "Check failed: " << cond << " "; }
  template <typename T> FatalStream& operator<<(const T& v) { ss << v; return *this; }
  ~FatalStream() noexcept(false) { Fatal(ss.str()); }
  std::ostringstream ss;
};
}  // namespace mock
#define CHECK_OR_RETURN(c) if (c) {} else return ::oneflow::mock::ErrorStream(#c)
#define CHECK_EQ_OR_RETURN(a, b) CHECK_OR_RETURN((a) == (b))
#define CHECK_GE_OR_RETURN(a, b) CHECK_OR_RETURN((a) >= (b))
#define CHECK_NOTNULL_OR_RETURN(p) CHECK_OR_RETURN((p) != nullptr)
// GNU statement expressions, as in the reference's just.h
#define JUST(expr) ({ auto&& just_m_ = (expr); if (!just_m_.IsOk()) return ::oneflow::mock::ErrorCarrier{just_m_.msg()}; just_m_.Value(); })
#define CHECK_JUST(expr) ({ auto&& cj_m_ = (expr); if (!cj_m_.IsOk()) ::oneflow::mock::Fatal(cj_m_.msg()); cj_m_.Value(); })
#define CHECK_EQ(a, b) if ((a) == (b)) {} else ::oneflow::mock::FatalStream(#a " == " #b)
#define CHECK_NOTNULL(p) if ((p) != nullptr) {} else ::oneflow::mock::FatalStream(#p " != nullptr")

// ---- Shape / ShapeView (oneflow/core/common/shape.h, shape_view.h)
class Shape {
 public:
  Shape() = default;
  Shape(std::initializer_list<int64_t> d) : dims_(d) {}
  explicit Shape(std::vector<int64_t> d) : dims_(std::move(d)) {}
  int64_t NumAxes() const { return static_cast<int64_t>(dims_.size()); }
  int64_t At(int64_t i) const { return dims_.at(i); }
  int64_t elem_cnt() const { int64_t n = 1; for (auto d : dims_) n *= d; return n; }
  bool operator==(const Shape& o) const { return dims_ == o.dims_; }
 private:
  std::vector<int64_t> dims_;
};
using ShapeView = Shape;

namespace ep {  // oneflow/core/ep/include/{stream.h:29-47,device.h:53-56,allocation_options.h}
struct AllocationOptions {};
class Device {
 public:
  virtual ~Device() = default;
  virtual Maybe<void> Alloc(const AllocationOptions&, void** ptr, size_t size) = 0;
  virtual void Free(const AllocationOptions&, void* ptr) = 0;
};
class Stream {
 public:
  virtual ~Stream() = default;
  virtual Device* device() const = 0;
  template <typename T> T* As() { return static_cast<T*>(this); }
};
}  // namespace ep

namespace user_op {

// oneflow/core/framework/user_op_tensor.h:31-72
class Tensor {
 public:
  virtual ShapeView shape_view() const = 0;
  virtual DataType data_type() const = 0;
  virtual const void* raw_dptr() const = 0;
  virtual void* mut_raw_dptr() = 0;
};

struct OpArg {  // oneflow/core/framework/user_op_conf.h
  OpArg(std::string n, int32_t i) : name(std::move(n)), index(i) {}
  std::string name;
  int32_t index;
};

// oneflow/core/framework/infer_util.h:52-64 (+ Attr<T>)
class InferContext {
 public:
  virtual ~InferContext() = default;
  virtual const Shape& InputShape(const std::string&, int32_t) const = 0;
  virtual void SetOutputShape(const std::string&, int32_t, const Shape&) = 0;
  virtual DataType InputDType(const std::string&, int32_t) const = 0;
  virtual void SetOutputDType(const std::string&, int32_t, DataType) = 0;
  virtual bool has_input(const std::string& arg_name, int32_t index) const = 0;   // infer_util.h:71
  template <typename T> const T& Attr(const std::string& name) const;
 protected:
  virtual const int64_t& AttrInt64(const std::string& name) const = 0;
  virtual const bool& AttrBool(const std::string& name) const = 0;
  virtual const DataType& AttrDataType(const std::string& name) const = 0;
};
template <> inline const int64_t& InferContext::Attr<int64_t>(const std::string& name) const { return AttrInt64(name); }
template <> inline const bool& InferContext::Attr<bool>(const std::string& name) const { return AttrBool(name); }
template <> inline const DataType& InferContext::Attr<DataType>(const std::string& name) const { return AttrDataType(name); }

// oneflow/core/framework/sbp_context.h (UserOpSbpSignatureBuilder)
class SbpSignatureBuilder {
 public:
  using Args = std::vector<std::pair<std::string, int32_t>>;
  explicit SbpSignatureBuilder(std::vector<std::string>* sink) : sink_(sink) {}
  SbpSignatureBuilder& Broadcast(const OpArg& a) { cur_ += a.name + ":B "; return *this; }
  SbpSignatureBuilder& Broadcast(const Args& as) { for (auto& a : as) cur_ += a.first + ":B "; return *this; }
  SbpSignatureBuilder& Split(const OpArg& a, int64_t axis) { cur_ += a.name + ":S(" + std::to_string(axis) + ") "; return *this; }
  SbpSignatureBuilder& PartialSum(const OpArg& a) { cur_ += a.name + ":P "; return *this; }
  void Build() { sink_->push_back(cur_); }
 private:
  std::vector<std::string>* sink_;
  std::string cur_;
};
class SbpContext {
 public:
  virtual ~SbpContext() = default;
  virtual const SbpSignatureBuilder::Args& inputs() const = 0;
  virtual const SbpSignatureBuilder::Args& outputs() const = 0;
  SbpSignatureBuilder NewBuilder() { return SbpSignatureBuilder(&signatures); }
  std::vector<std::string> signatures;
};

class InputArgModifier {  // oneflow/core/operator/arg_modifier_signature.proto
 public:
  void set_requires_grad(bool v) { requires_grad_ = v; }
  bool requires_grad() const { return requires_grad_; }
 private:
  bool requires_grad_ = true;
};
using GetInputArgModifier = std::function<InputArgModifier*(const std::string&, int32_t)>;
class UserOpConfWrapper {};

// oneflow/core/framework/op_kernel.h:249-334
class OpKernelState { public: virtual ~OpKernelState() = default; };
class OpKernelCache { public: virtual ~OpKernelCache() = default; };
class KernelInitContext {};
class KernelCacheContext {};
class KernelComputeContext {
 public:
  virtual ~KernelComputeContext() = default;
  virtual Tensor* Tensor4ArgNameAndIndex(const std::string&, int32_t) = 0;
  virtual ep::Stream* stream() = 0;
  virtual bool has_input(const std::string& arg_name, int32_t index) const = 0;   // op_kernel.h:120-122
  template <typename T> const T& Attr(const std::string& name) const;
 protected:
  virtual const int64_t& AttrInt64(const std::string& name) const = 0;
  virtual const bool& AttrBool(const std::string& name) const = 0;
};
template <> inline const int64_t& KernelComputeContext::Attr<int64_t>(const std::string& name) const { return AttrInt64(name); }
template <> inline const bool& KernelComputeContext::Attr<bool>(const std::string& name) const { return AttrBool(name); }

class OpKernel {
 public:
  virtual ~OpKernel() = default;
  virtual std::shared_ptr<OpKernelState> CreateOpKernelState(KernelInitContext*) const { return nullptr; }
  virtual void Compute(KernelComputeContext* ctx, OpKernelState*, const OpKernelCache*) const { Compute(ctx); }
  virtual void Compute(KernelComputeContext*) const {}
  virtual bool AlwaysComputeWhenAllOutputsEmpty() const = 0;
};

// ---- registration: REGISTER_USER_KERNEL(name).SetCreateFn<K>().SetIsMatchedHob(e).SetInferTmpSizeFn(f)
// (oneflow/core/framework/user_op_registry_manager.h:70-84, user_op_kernel_registry.h:77-94,
//  user_op_hob.h:31-71)
struct KernelMatchQuery { DeviceType device; std::map<std::string, DataType> dtypes; };
using Hob = std::function<bool(const KernelMatchQuery&)>;
inline Hob operator&&(Hob a, Hob b) { return [a, b](const KernelMatchQuery& q) { return a(q) && b(q); }; }
struct HobDeviceTypeT {};
inline HobDeviceTypeT HobDeviceType() { return {}; }
inline Hob operator==(HobDeviceTypeT, DeviceType d) { return [d](const KernelMatchQuery& q) { return q.device == d; }; }
struct HobDataTypeT { std::string arg; };
inline HobDataTypeT HobDataType(const std::string& arg, int32_t) { return {arg}; }
inline Hob operator==(HobDataTypeT h, DataType d) {
  return [h, d](const KernelMatchQuery& q) { auto it = q.dtypes.find(h.arg); return it != q.dtypes.end() && it->second == d; };
}
struct KernelRegistration {
  std::string op;
  std::function<std::unique_ptr<OpKernel>()> create;
  Hob matched;
  std::function<size_t(InferContext*)> infer_tmp_size;
};
std::vector<KernelRegistration>& KernelRegistry();
class KernelRegistryBuilder {
 public:
  explicit KernelRegistryBuilder(const std::string& op) { reg_.op = op; }
  template <typename K> KernelRegistryBuilder& SetCreateFn() { reg_.create = [] { return std::unique_ptr<OpKernel>(new K()); }; return *this; }
  KernelRegistryBuilder& SetIsMatchedHob(Hob h) { reg_.matched = std::move(h); return *this; }
  KernelRegistryBuilder& SetInferTmpSizeFn(std::function<size_t(InferContext*)> f) { reg_.infer_tmp_size = std::move(f); return *this; }
  KernelRegistration Finish() const { return reg_; }
 private:
  KernelRegistration reg_;
};
struct KernelRegisterTrigger { KernelRegisterTrigger(const KernelRegistryBuilder& b) { KernelRegistry().push_back(b.Finish()); } };

}  // namespace user_op

#define OF_MOCK_CAT_(a, b) a##b
#define OF_MOCK_CAT(a, b) OF_MOCK_CAT_(a, b)
#define REGISTER_USER_KERNEL(name) \
  static ::oneflow::user_op::KernelRegisterTrigger OF_MOCK_CAT(g_register_trigger_, __COUNTER__) = \
      ::oneflow::user_op::KernelRegistryBuilder(name)

}  // namespace oneflow

// glue_harness.cpp — compiles the OneFlow glue sources (of-spmm_b200/oneflow_glue/spmm_op.cpp,
// spmm_kernels.cpp) against tests/mock_oneflow and drives them the way the reference's framework
// would: op inference through InferContext / SbpContext, kernel lookup through the
// REGISTER_USER_KERNEL predicates, InferTmpSize, then OpKernel::Compute(KernelComputeContext*)
// on a CUDA stream.  TEST INFRASTRUCTURE (reads like a reference test: build inputs, run the op,
// compare with a plain loop).
//
//   glue_harness host   — everything that needs no GPU (inference, errors, SBP, registry)
//   glue_harness gpu    — additionally runs spmm_csr / spmm_csr_grad_b / sddmm_csr kernels
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

#include "oneflow/core/ep/cuda/cuda_stream.h"
#include "oneflow/core/framework/autograd_mock.h"
#include "oneflow/core/framework/framework.h"
#include "oneflow/core/framework/op_generated.h"
#include "ofspmm.h"

namespace oneflow {
namespace mock {
void Fatal(const std::string& msg) {
  std::fprintf(stderr, "FATAL: %s\n", msg.c_str());
  std::exit(70);
}
}  // namespace mock
namespace user_op {
std::vector<KernelRegistration>& KernelRegistry() {
  static std::vector<KernelRegistration> r;
  return r;
}
}  // namespace user_op
namespace one {
std::vector<DispatchRecord>& DispatchLog() {
  static std::vector<DispatchRecord> r;
  return r;
}
std::map<std::string, std::function<OpExprGradFunctionIf*()>>& GradFunctionRegistry() {
  static std::map<std::string, std::function<OpExprGradFunctionIf*()>> r;
  return r;
}
namespace functional {
std::map<std::string, std::any>& FunctionLibraryStore() {
  static std::map<std::string, std::any> r;
  return r;
}
}  // namespace functional
}  // namespace one
}  // namespace oneflow

using namespace oneflow;

#define EXPECT(cond)                                                        \
  do {                                                                      \
    if (!(cond)) {                                                          \
      std::fprintf(stderr, "%s:%d: EXPECT failed: %s\n", __FILE__, __LINE__, #cond); \
      std::exit(1);                                                         \
    }                                                                       \
  } while (0)

namespace {

struct TensorDescs {
  std::map<std::string, Shape> shapes;
  std::map<std::string, DataType> dtypes;
  std::map<std::string, int64_t> attrs;
  std::map<std::string, bool> battrs{{"atomic", false}, {"relu", false}};
  std::map<std::string, DataType> dattrs{{"val_dtype", kInvalidDataType}};
  std::vector<std::string> present;   // optional inputs that are bound
};

class MockInferContext final : public user_op::InferContext {
 public:
  explicit MockInferContext(TensorDescs* d) : d_(d) {}
  const Shape& InputShape(const std::string& n, int32_t) const override { return d_->shapes.at(n); }
  void SetOutputShape(const std::string& n, int32_t, const Shape& s) override { d_->shapes[n] = s; }
  DataType InputDType(const std::string& n, int32_t) const override { return d_->dtypes.at(n); }
  void SetOutputDType(const std::string& n, int32_t, DataType t) override { d_->dtypes[n] = t; }
  bool has_input(const std::string& n, int32_t) const override {
    return std::find(d_->present.begin(), d_->present.end(), n) != d_->present.end();
  }
 protected:
  const int64_t& AttrInt64(const std::string& n) const override { return d_->attrs.at(n); }
  const bool& AttrBool(const std::string& n) const override { return d_->battrs.at(n); }
  const DataType& AttrDataType(const std::string& n) const override { return d_->dattrs.at(n); }
 private:
  TensorDescs* d_;
};

class MockSbpContext final : public user_op::SbpContext {
 public:
  MockSbpContext(user_op::SbpSignatureBuilder::Args in, user_op::SbpSignatureBuilder::Args out)
      : in_(std::move(in)), out_(std::move(out)) {}
  const user_op::SbpSignatureBuilder::Args& inputs() const override { return in_; }
  const user_op::SbpSignatureBuilder::Args& outputs() const override { return out_; }
 private:
  user_op::SbpSignatureBuilder::Args in_, out_;
};

TensorDescs SpmmDescs(int64_t rows, int64_t cols, int64_t nnz, int64_t n, DataType dense, DataType idx) {
  TensorDescs d;
  d.shapes["a_crow"] = Shape({rows + 1});
  d.shapes["a_col"] = Shape({nnz});
  d.shapes["a_val"] = Shape({nnz});
  d.shapes["b"] = Shape({cols, n});
  d.shapes["dy"] = Shape({rows, n});
  d.dtypes["a_crow"] = idx;
  d.dtypes["a_col"] = idx;
  d.dtypes["a_val"] = kFloat;
  d.dtypes["b"] = dense;
  d.dtypes["dy"] = dense;
  d.attrs["a_rows"] = rows;
  d.attrs["a_cols"] = cols;
  return d;
}

const user_op::KernelRegistration* FindKernel(const std::string& op, const user_op::KernelMatchQuery& q, int* matches) {
  const user_op::KernelRegistration* found = nullptr;
  *matches = 0;
  for (const auto& r : user_op::KernelRegistry())
    if (r.op == op && r.matched(q)) { found = &r; ++*matches; }
  return found;
}

void HostChecks() {
  // ---- shape / dtype inference
  TensorDescs d = SpmmDescs(300, 200, 1234, 64, kFloat, kInt32);
  MockInferContext ic(&d);
  EXPECT(SpmmCsrOp::InferLogicalTensorDesc(&ic).IsOk());
  EXPECT(d.shapes.at("out") == Shape({300, 64}));
  EXPECT(SpmmCsrOp::InferDataType(&ic).IsOk());
  EXPECT(d.dtypes.at("out") == kFloat);
  EXPECT(SpmmCsrGradBOp::InferLogicalTensorDesc(&ic).IsOk());
  EXPECT(d.shapes.at("db") == Shape({200, 64}));
  EXPECT(SddmmCsrOp::InferLogicalTensorDesc(&ic).IsOk());
  EXPECT(d.shapes.at("dval") == Shape({1234}));
  // the SDDMM gradient takes the dtype of the values it differentiates (attr val_dtype), not b's
  TensorDescs db16 = SpmmDescs(300, 200, 1234, 64, kBFloat16, kInt32);
  db16.dattrs["val_dtype"] = kFloat;
  MockInferContext ic_b16(&db16);
  EXPECT(SddmmCsrOp::InferDataType(&ic_b16).IsOk() && db16.dtypes.at("dval") == kFloat);
  db16.dattrs["val_dtype"] = kInvalidDataType;
  EXPECT(SddmmCsrOp::InferDataType(&ic_b16).IsOk() && db16.dtypes.at("dval") == kBFloat16);
  // csr_transpose_structure: three index outputs
  EXPECT(CsrTransposeStructureOp::InferLogicalTensorDesc(&ic).IsOk() && CsrTransposeStructureOp::InferDataType(&ic).IsOk());
  EXPECT(d.shapes.at("t_crow") == Shape({201}) && d.shapes.at("t_col") == Shape({1234}) && d.shapes.at("t_perm") == Shape({1234}));
  EXPECT(d.dtypes.at("t_perm") == kInt32);
  // spmm_csr_grad_b with the cached structure bound: shapes are checked
  d.present = {"t_crow", "t_col", "t_perm"};
  EXPECT(SpmmCsrGradBOp::InferLogicalTensorDesc(&ic).IsOk());
  d.shapes["t_crow"] = Shape({7});
  Maybe<void> mt = SpmmCsrGradBOp::InferLogicalTensorDesc(&ic);
  EXPECT(!mt.IsOk() && mt.msg().find("a_cols+1") != std::string::npos);
  d.present.clear();
  // errors surface as failed Maybe<void> with the message of the failing CHECK
  TensorDescs bad = SpmmDescs(300, 200, 1234, 64, kFloat, kInt32);
  bad.attrs["a_cols"] = 7;
  MockInferContext ic_bad(&bad);
  Maybe<void> m = SpmmCsrOp::InferLogicalTensorDesc(&ic_bad);
  EXPECT(!m.IsOk() && m.msg().find("a_cols") != std::string::npos);
  bad = SpmmDescs(300, 200, 1234, 64, kFloat, kInt32);
  bad.shapes["a_crow"] = Shape({300});
  MockInferContext ic_bad2(&bad);
  m = SpmmCsrOp::InferLogicalTensorDesc(&ic_bad2);
  EXPECT(!m.IsOk() && m.msg().find("a_rows+1") != std::string::npos);
  bad = SpmmDescs(300, 200, 1234, 64, kFloat, kInt32);
  bad.dtypes["a_crow"] = kFloat;
  MockInferContext ic_bad3(&bad);
  m = SpmmCsrOp::InferDataType(&ic_bad3);
  EXPECT(!m.IsOk() && m.msg().find("index dtype") != std::string::npos);
  // ---- SBP: CSR arrays broadcast, dense side column-split; SDDMM column split -> partial sum
  MockSbpContext sc({{"a_crow", 0}, {"a_col", 0}, {"a_val", 0}, {"b", 0}}, {{"out", 0}});
  EXPECT(SpmmCsrOp::GetSbp(&sc).IsOk());
  EXPECT(sc.signatures.size() == 4);
  EXPECT(sc.signatures[0].find("b:S(1)") != std::string::npos && sc.signatures[0].find("out:S(1)") != std::string::npos);
  EXPECT(sc.signatures[0].find("a_col:B") != std::string::npos);
  // bilinear: partial sums of the values or of the dense operand give a partial-sum output
  EXPECT(sc.signatures[1].find("a_val:P") != std::string::npos && sc.signatures[1].find("b:B") != std::string::npos &&
         sc.signatures[1].find("out:P") != std::string::npos && sc.signatures[1].find("a_col:B") != std::string::npos);
  EXPECT(sc.signatures[2].find("a_val:B") != std::string::npos && sc.signatures[2].find("b:P") != std::string::npos &&
         sc.signatures[2].find("out:P") != std::string::npos);
  // spmm_csr_grad_b: same family (incl. the P-sum signatures), cached structure broadcast
  MockSbpContext sg({{"a_crow", 0}, {"a_col", 0}, {"a_val", 0}, {"dy", 0}, {"t_crow", 0}, {"t_col", 0}, {"t_perm", 0}}, {{"db", 0}});
  EXPECT(SpmmCsrGradBOp::GetSbp(&sg).IsOk() && sg.signatures.size() == 4);
  EXPECT(sg.signatures[2].find("dy:P") != std::string::npos && sg.signatures[2].find("db:P") != std::string::npos &&
         sg.signatures[2].find("t_perm:B") != std::string::npos);
  EXPECT(sg.signatures[1].find("a_val:P") != std::string::npos && sg.signatures[1].find("db:P") != std::string::npos);
  MockSbpContext sd({{"a_crow", 0}, {"a_col", 0}, {"dy", 0}, {"b", 0}}, {{"dval", 0}});
  EXPECT(SddmmCsrOp::GetSbp(&sd).IsOk());
  EXPECT(sd.signatures[0].find("dval:P") != std::string::npos);
  EXPECT(sd.signatures.size() == 4 && sd.signatures[1].find("dy:P") != std::string::npos &&
         sd.signatures[2].find("b:P") != std::string::npos);
  // ---- fused_spmm_csr_bias_act: spmm_csr's checks plus the bias row; no partial-sum signature
  TensorDescs df = SpmmDescs(300, 200, 1234, 64, kFloat, kInt32);
  df.shapes["bias"] = Shape({64});
  df.dtypes["bias"] = kFloat;
  MockInferContext icf(&df);
  EXPECT(FusedSpmmCsrBiasActOp::InferLogicalTensorDesc(&icf).IsOk() && df.shapes.at("out") == Shape({300, 64}));
  EXPECT(FusedSpmmCsrBiasActOp::InferDataType(&icf).IsOk() && df.dtypes.at("out") == kFloat);
  df.shapes["bias"] = Shape({63});
  Maybe<void> mf = FusedSpmmCsrBiasActOp::InferLogicalTensorDesc(&icf);
  EXPECT(!mf.IsOk() && mf.msg().find("one entry per column") != std::string::npos);
  df.shapes["bias"] = Shape({64});
  df.dtypes["bias"] = kBFloat16;
  mf = FusedSpmmCsrBiasActOp::InferDataType(&icf);
  EXPECT(!mf.IsOk() && mf.msg().find("share a dtype") != std::string::npos);
  MockSbpContext sf({{"a_crow", 0}, {"a_col", 0}, {"a_val", 0}, {"b", 0}, {"bias", 0}}, {{"out", 0}});
  EXPECT(FusedSpmmCsrBiasActOp::GetSbp(&sf).IsOk() && sf.signatures.size() == 2);
  EXPECT(sf.signatures[0].find("b:S(1)") != std::string::npos && sf.signatures[0].find("bias:S(0)") != std::string::npos &&
         sf.signatures[0].find("out:S(1)") != std::string::npos && sf.signatures[0].find("a_val:B") != std::string::npos);
  for (const auto& sig : sf.signatures) EXPECT(sig.find(":P") == std::string::npos);   // bias / ReLU are not linear
  // ---- index inputs never require grad (ModifyInputArg)
  std::map<std::string, user_op::InputArgModifier> mods;
  auto getter = [&](const std::string& n, int32_t) { return &mods[n]; };
  EXPECT(SpmmCsrOp::ModifyInputArg(getter, user_op::UserOpConfWrapper()).IsOk());
  EXPECT(!mods["a_crow"].requires_grad() && !mods["a_col"].requires_grad());
  // ---- registry: exactly one kernel per (CUDA, dense dtype, index dtype); none for CPU
  int matches = 0;
  for (const char* op : {"spmm_csr", "fused_spmm_csr_bias_act", "spmm_csr_grad_b", "sddmm_csr"}) {
    const char* dense_arg = std::string(op) == "spmm_csr_grad_b" ? "dy" : "b";
    for (DataType dense : {kFloat, kBFloat16})
      for (DataType idx : {kInt32, kInt64}) {
        user_op::KernelMatchQuery q{DeviceType::kCUDA, {{dense_arg, dense}, {"a_col", idx}}};
        EXPECT(FindKernel(op, q, &matches) != nullptr && matches == 1);
      }
    user_op::KernelMatchQuery cpu{DeviceType::kCPU, {{dense_arg, kFloat}, {"a_col", kInt32}}};
    FindKernel(op, cpu, &matches);
    EXPECT(matches == 0);  // no CPU kernel on this path
    user_op::KernelMatchQuery f16{DeviceType::kCUDA, {{dense_arg, kFloat16}, {"a_col", kInt32}}};
    FindKernel(op, f16, &matches);
    EXPECT(matches == 0);
  }
  // ---- tmp_buffer size comes from the library's workspace query
  user_op::KernelMatchQuery q{DeviceType::kCUDA, {{"b", kFloat}, {"a_col", kInt32}}};
  const auto* reg = FindKernel("spmm_csr", q, &matches);
  TensorDescs d2 = SpmmDescs(300, 200, 1234, 64, kFloat, kInt32);
  MockInferContext ic2(&d2);
  EXPECT(reg->infer_tmp_size(&ic2) == ofspmm_fwd_workspace_bytes(300, 200, 1234, 64, OFSPMM_DTYPE_FLOAT));
  EXPECT(FindKernel("fused_spmm_csr_bias_act", q, &matches)->infer_tmp_size(&ic2) ==
         ofspmm_fwd_ex_workspace_bytes(300, 200, 1234, 64, OFSPMM_DTYPE_FLOAT, OFSPMM_VARIANT_AUTO));
  // spmm_csr_grad_b: the route decides the tmp size — transient (default), atomic, cached structure
  user_op::KernelMatchQuery qg{DeviceType::kCUDA, {{"dy", kFloat}, {"a_col", kInt32}}};
  const auto* regg = FindKernel("spmm_csr_grad_b", qg, &matches);
  EXPECT(regg->infer_tmp_size(&ic2) ==
         ofspmm_bwd_b_transient_workspace_bytes(300, 200, 1234, 64, OFSPMM_DTYPE_FLOAT, OFSPMM_DTYPE_INT32, OFSPMM_DTYPE_FLOAT));
  d2.battrs["atomic"] = true;
  EXPECT(regg->infer_tmp_size(&ic2) == ofspmm_bwd_b_workspace_bytes(300, 200, 1234, 64, OFSPMM_DTYPE_FLOAT, 0));
  d2.battrs["atomic"] = false;
  d2.present = {"t_crow", "t_col", "t_perm"};
  EXPECT(regg->infer_tmp_size(&ic2) == ofspmm_bwd_b_cached_workspace_bytes(300, 200, 1234, 64, OFSPMM_DTYPE_FLOAT, OFSPMM_DTYPE_FLOAT));
  user_op::KernelMatchQuery qt{DeviceType::kCUDA, {{"a_col", kInt64}}};
  EXPECT(FindKernel("csr_transpose_structure", qt, &matches) != nullptr && matches == 1);
  std::printf("glue host checks ok: %zu kernel registrations\n", user_op::KernelRegistry().size());
}

// ------------------------------------------------------------------ functional layer + grad function
// (SURVEY.md §8 a8 / a9): functor -> Dispatch with the right op, input order and attrs; the grad
// function saves only what it needs and dispatches sddmm_csr / spmm_csr_grad_b iff required.
void AutogradChecks() {
  using one::Tensor;
  auto T = [](const char* name, std::vector<int64_t> dims, bool rg) { return std::make_shared<Tensor>(name, std::move(dims), rg); };
  auto& log = one::DispatchLog();
  auto crow = T("crow", {301}, false), col = T("col", {1234}, false);
  // ---- functor: flow._C.spmm_csr(...)
  log.clear();
  auto out = one::functional::SpmmCsr(crow, col, T("val", {1234}, true), T("b", {200, 64}, true), 300, 200);
  EXPECT(out.IsOk() && log.size() == 1 && log[0].op == "spmm_csr");
  EXPECT((log[0].inputs == std::vector<std::string>{"crow", "col", "val", "b"}));
  // ... with the cached structure of A^T riding along as optional inputs
  auto tcrow = T("t_crow", {201}, false), tcol = T("t_col", {1234}, false), tperm = T("t_perm", {1234}, false);
  log.clear();
  out = one::functional::SpmmCsr(crow, col, T("val", {1234}, true), T("b", {200, 64}, true), 300, 200, tcrow, tcol, tperm);
  EXPECT(out.IsOk() && log.size() == 1 && log[0].inputs.size() == 7 && log[0].inputs[6] == "t_perm");
  auto half = one::functional::SpmmCsr(crow, col, T("val", {1234}, true), T("b", {200, 64}, true), 300, 200, tcrow);
  EXPECT(!half.IsOk() && half.msg().find("come together") != std::string::npos);
  // flow._C.csr_transpose_structure -> three tensors
  log.clear();
  auto ts = one::functional::CsrTransposeStructure(crow, col, 300, 200);
  EXPECT(ts.IsOk() && ts.Value()->size() == 3 && log[0].op == "csr_transpose_structure");
  EXPECT(log[0].attrs.at("a_rows") == 300 && log[0].attrs.at("a_cols") == 200);
  auto bad = one::functional::SpmmCsr(crow, col, T("val", {1234}, true), T("b", {200}, true), 300, 200);
  EXPECT(!bad.IsOk() && bad.msg().find("RuntimeError") != std::string::npos && bad.msg().find("2-D") != std::string::npos);
  bad = one::functional::SpmmCsr(crow, col, T("val", {1234}, true), T("b", {7, 64}, true), 300, 200);
  EXPECT(!bad.IsOk() && bad.msg().find("a_cols") != std::string::npos);

  // ---- grad function registered for "spmm_csr"
  EXPECT(one::GradFunctionRegistry().count("spmm_csr") == 1);
  one::UserOpExpr fw;
  fw.op_type_name = "spmm_csr";
  fw.proto_.attrs.ints = {{"a_rows", 300}, {"a_cols", 200}};
  auto run_case = [&](bool val_rg, bool b_rg, one::TensorTuple* in_grads, bool with_at = false, DataType vdt = kFloat) {
    std::unique_ptr<one::OpExprGradFunctionIf> g(one::GradFunctionRegistry().at("spmm_csr")());
    EXPECT(g->Init(fw).IsOk());
    auto state = g->MakeCustomState();
    one::TensorTuple inputs = {crow, col, std::make_shared<Tensor>("val", std::vector<int64_t>{1234}, val_rg, vdt),
                               T("b", {200, 64}, b_rg)};
    if (with_at) { inputs.push_back(tcrow); inputs.push_back(tcol); inputs.push_back(tperm); }
    one::TensorTuple outputs = {T("out", {300, 64}, val_rg || b_rg)};
    EXPECT(g->CaptureIf(state.get(), inputs, outputs, AttrMap()).IsOk());
    log.clear();
    in_grads->assign(with_at ? 7 : 4, nullptr);
    EXPECT(g->ApplyIf(state.get(), {T("dy", {300, 64}, false)}, in_grads).IsOk());
    return state->SavedTensors().size();
  };
  one::TensorTuple ig;
  // both need grad: SDDMM for the values, A^T·dy for b; index inputs get nothing
  size_t saved = run_case(true, true, &ig);
  EXPECT(saved == 4 && log.size() == 2);
  EXPECT(log[0].op == "sddmm_csr" && (log[0].inputs == std::vector<std::string>{"crow", "col", "dy", "b"}));
  EXPECT(log[1].op == "spmm_csr_grad_b" && (log[1].inputs == std::vector<std::string>{"crow", "col", "val", "dy"}));
  EXPECT(log[1].attrs.at("a_rows") == 300 && log[1].attrs.at("a_cols") == 200 && log[1].attrs.at("atomic") == 0);
  EXPECT(log[0].attrs.at("val_dtype") == kFloat);   // SDDMM gradient in the dtype of a_val
  EXPECT(ig[0] == nullptr && ig[1] == nullptr && ig[2] != nullptr && ig[3] != nullptr);
  // cached structure of A^T given to the forward: saved and handed to spmm_csr_grad_b, no grads for it
  saved = run_case(true, true, &ig, /*with_at=*/true, kBFloat16);
  EXPECT(saved == 7 && log.size() == 2 && log[1].op == "spmm_csr_grad_b");
  EXPECT((log[1].inputs == std::vector<std::string>{"crow", "col", "val", "dy", "t_crow", "t_col", "t_perm"}));
  EXPECT(log[0].attrs.at("val_dtype") == kBFloat16);
  EXPECT(ig.size() == 7 && ig[4] == nullptr && ig[5] == nullptr && ig[6] == nullptr);
  // only b: no SDDMM, b itself is not saved
  saved = run_case(false, true, &ig);
  EXPECT(saved == 3 && log.size() == 1 && log[0].op == "spmm_csr_grad_b" && ig[2] == nullptr && ig[3] != nullptr);
  // only the values: no A^T·dy, the values themselves are not saved
  saved = run_case(true, false, &ig);
  EXPECT(saved == 3 && log.size() == 1 && log[0].op == "sddmm_csr" && ig[2] != nullptr && ig[3] == nullptr);
  // nothing requires grad: nothing saved, nothing dispatched
  saved = run_case(false, false, &ig);
  EXPECT(saved == 0 && log.empty());
  // ---- fused_spmm_csr_bias_act: functor and grad function
  auto bias = T("bias", {64}, true);
  log.clear();
  out = one::functional::FusedSpmmCsrBiasAct(crow, col, T("val", {1234}, true), T("b", {200, 64}, true), bias, 300, 200, true);
  EXPECT(out.IsOk() && log.size() == 1 && log[0].op == "fused_spmm_csr_bias_act");
  EXPECT((log[0].inputs == std::vector<std::string>{"crow", "col", "val", "b", "bias"}) && log[0].attrs.at("relu") == 1);
  bad = one::functional::FusedSpmmCsrBiasAct(crow, col, T("val", {1234}, true), T("b", {200, 64}, true), T("bias", {63}, true), 300, 200);
  EXPECT(!bad.IsOk() && bad.msg().find("63 entries") != std::string::npos);
  EXPECT(one::GradFunctionRegistry().count("fused_spmm_csr_bias_act") == 1);
  auto run_fused = [&](bool val_rg, bool b_rg, bool bias_rg, bool relu, one::TensorTuple* in_grads) {
    one::UserOpExpr ff;
    ff.op_type_name = "fused_spmm_csr_bias_act";
    ff.proto_.attrs.ints = {{"a_rows", 300}, {"a_cols", 200}, {"relu", relu ? 1 : 0}};
    std::unique_ptr<one::OpExprGradFunctionIf> g(one::GradFunctionRegistry().at("fused_spmm_csr_bias_act")());
    EXPECT(g->Init(ff).IsOk());
    auto state = g->MakeCustomState();
    one::TensorTuple inputs = {crow, col, T("val", {1234}, val_rg), T("b", {200, 64}, b_rg), T("bias", {64}, bias_rg)};
    one::TensorTuple outputs = {T("out", {300, 64}, val_rg || b_rg || bias_rg)};
    EXPECT(g->CaptureIf(state.get(), inputs, outputs, AttrMap()).IsOk());
    log.clear();
    in_grads->assign(5, nullptr);
    EXPECT(g->ApplyIf(state.get(), {T("dy", {300, 64}, false)}, in_grads).IsOk());
    return state->SavedTensors().size();
  };
  // everything needs grad, ReLU on: mask from the saved output, then bias / values / dense gradients of dz
  saved = run_fused(true, true, true, true, &ig);
  EXPECT(saved == 5 && log.size() == 4);
  EXPECT(log[0].op == "relu_grad" && (log[0].inputs == std::vector<std::string>{"dy", "out"}));
  EXPECT(log[1].op == "reduce_sum" && (log[1].inputs == std::vector<std::string>{"relu_grad:dx"}) && log[1].attrs.at("axis0") == 0 &&
         log[1].attrs.at("keepdims") == 0);
  EXPECT(log[2].op == "sddmm_csr" && (log[2].inputs == std::vector<std::string>{"crow", "col", "relu_grad:dx", "b"}));
  EXPECT(log[3].op == "spmm_csr_grad_b" && (log[3].inputs == std::vector<std::string>{"crow", "col", "val", "relu_grad:dx"}));
  EXPECT(ig[0] == nullptr && ig[1] == nullptr && ig[2] != nullptr && ig[3] != nullptr && ig[4] != nullptr);
  // no activation: dy goes straight through, the output is not saved
  saved = run_fused(true, true, true, false, &ig);
  EXPECT(saved == 4 && log.size() == 3 && log[0].op == "reduce_sum" && (log[0].inputs == std::vector<std::string>{"dy"}));
  EXPECT((log[2].inputs == std::vector<std::string>{"crow", "col", "val", "dy"}));
  // only the bias: one reduction, no sparse kernel, no CSR saved
  saved = run_fused(false, false, true, true, &ig);
  EXPECT(saved == 1 && log.size() == 2 && log[1].op == "reduce_sum" && ig[2] == nullptr && ig[3] == nullptr && ig[4] != nullptr);
  // nothing requires grad
  saved = run_fused(false, false, false, true, &ig);
  EXPECT(saved == 0 && log.empty());
  std::printf("glue autograd checks ok: functors + OpExprGradFunction dispatch as specified\n");
}

// ------------------------------------------------------------------ GPU part
#define CUDA_OK(expr)                                                                    \
  do {                                                                                   \
    cudaError_t e_ = (expr);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      std::fprintf(stderr, "%s:%d: %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));   \
      std::exit(2);                                                                      \
    }                                                                                    \
  } while (0)

class CudaMallocDevice final : public ep::Device {
 public:
  Maybe<void> Alloc(const ep::AllocationOptions&, void** ptr, size_t size) override {
    if (cudaMalloc(ptr, size ? size : 1) != cudaSuccess) return Maybe<void>::Error("cudaMalloc failed");
    ++live;
    return Maybe<void>::Ok();
  }
  void Free(const ep::AllocationOptions&, void* ptr) override {
    cudaFree(ptr);
    --live;
  }
  int live = 0;
};

class DevTensor final : public user_op::Tensor {
 public:
  DevTensor(Shape s, DataType t, const void* host, size_t bytes) : shape_(std::move(s)), dtype_(t), bytes_(bytes) {
    CUDA_OK(cudaMalloc(&ptr_, bytes ? bytes : 1));
    if (host != nullptr && bytes) CUDA_OK(cudaMemcpy(ptr_, host, bytes, cudaMemcpyHostToDevice));
  }
  ~DevTensor() { cudaFree(ptr_); }
  ShapeView shape_view() const override { return shape_; }
  DataType data_type() const override { return dtype_; }
  const void* raw_dptr() const override { return ptr_; }
  void* mut_raw_dptr() override { return ptr_; }
  void ToHost(void* dst) const { CUDA_OK(cudaMemcpy(dst, ptr_, bytes_, cudaMemcpyDeviceToHost)); }
 private:
  Shape shape_;
  DataType dtype_;
  size_t bytes_;
  void* ptr_ = nullptr;
};

class MockComputeContext final : public user_op::KernelComputeContext {
 public:
  MockComputeContext(ep::Stream* s, std::map<std::string, int64_t> attrs) : stream_(s), attrs_(std::move(attrs)) {}
  user_op::Tensor* Tensor4ArgNameAndIndex(const std::string& n, int32_t) override {
    auto it = tensors.find(n);
    return it == tensors.end() ? nullptr : it->second;
  }
  ep::Stream* stream() override { return stream_; }
  bool has_input(const std::string& n, int32_t) const override { return tensors.count(n) != 0; }
  std::map<std::string, user_op::Tensor*> tensors;
  bool atomic = false;
 protected:
  const int64_t& AttrInt64(const std::string& n) const override { return attrs_.at(n); }
  const bool& AttrBool(const std::string&) const override { return atomic; }
 private:
  ep::Stream* stream_;
  std::map<std::string, int64_t> attrs_;
};

void GpuChecks() {
  const int64_t M = 700, K = 500, N = 64;
  std::mt19937 rng(7);
  std::uniform_real_distribution<float> uni(-1.f, 1.f);
  std::vector<int32_t> crow(M + 1, 0), col;
  std::vector<float> val;
  for (int64_t i = 0; i < M; ++i) {
    const int len = i % 97 == 0 ? 400 : static_cast<int>(rng() % 12);  // a few long rows, some empty
    std::vector<int32_t> cs;
    for (int t = 0; t < len; ++t) cs.push_back(static_cast<int32_t>(rng() % K));
    std::sort(cs.begin(), cs.end());
    cs.erase(std::unique(cs.begin(), cs.end()), cs.end());
    for (int32_t c : cs) { col.push_back(c); val.push_back(uni(rng)); }
    crow[i + 1] = static_cast<int32_t>(col.size());
  }
  const int64_t nnz = static_cast<int64_t>(col.size());
  std::vector<float> B(K * N), dY(M * N);
  for (auto& x : B) x = uni(rng);
  for (auto& x : dY) x = 0.5f * (uni(rng) + 1.f);
  // plain-loop expectations (double accumulate)
  std::vector<double> C(M * N, 0.0), dB(K * N, 0.0), dval(nnz, 0.0);
  for (int64_t i = 0; i < M; ++i)
    for (int32_t p = crow[i]; p < crow[i + 1]; ++p)
      for (int64_t j = 0; j < N; ++j) {
        C[i * N + j] += static_cast<double>(val[p]) * B[col[p] * N + j];
        dB[col[p] * N + j] += static_cast<double>(val[p]) * dY[i * N + j];
        dval[p] += static_cast<double>(dY[i * N + j]) * B[col[p] * N + j];
      }

  cudaStream_t cs;
  CUDA_OK(cudaStreamCreate(&cs));
  CudaMallocDevice device;
  ep::CudaStream stream(cs, &device);
  DevTensor t_crow(Shape({M + 1}), kInt32, crow.data(), crow.size() * 4);
  DevTensor t_col(Shape({nnz}), kInt32, col.data(), col.size() * 4);
  DevTensor t_val(Shape({nnz}), kFloat, val.data(), val.size() * 4);
  DevTensor t_b(Shape({K, N}), kFloat, B.data(), B.size() * 4);
  DevTensor t_dy(Shape({M, N}), kFloat, dY.data(), dY.size() * 4);
  DevTensor t_out(Shape({M, N}), kFloat, nullptr, C.size() * 4);
  DevTensor t_db(Shape({K, N}), kFloat, nullptr, dB.size() * 4);
  DevTensor t_dval(Shape({nnz}), kFloat, nullptr, dval.size() * 4);
  TensorDescs d = SpmmDescs(M, K, nnz, N, kFloat, kInt32);
  MockInferContext ic(&d);
  int matches = 0;

  auto run = [&](const char* op, const char* dense_arg, std::map<std::string, user_op::Tensor*> tensors, int repeat,
                 bool atomic = false) {
    user_op::KernelMatchQuery q{DeviceType::kCUDA, {{dense_arg, std::string(dense_arg) == "a_col" ? kInt32 : kFloat}, {"a_col", kInt32}}};
    const auto* reg = FindKernel(op, q, &matches);
    EXPECT(reg != nullptr && matches == 1);
    d.battrs["atomic"] = atomic;
    d.present.clear();
    for (const char* opt : {"t_crow", "t_col", "t_perm"})
      if (tensors.count(opt) && std::string(op) != "csr_transpose_structure") d.present.push_back(opt);
    const size_t tmp = reg->infer_tmp_size(&ic);
    DevTensor t_tmp(Shape({static_cast<int64_t>(tmp)}), kChar, nullptr, tmp);
    tensors["tmp_buffer"] = &t_tmp;
    MockComputeContext ctx(&stream, {{"a_rows", M}, {"a_cols", K}});
    ctx.tensors = tensors;
    ctx.atomic = atomic;
    std::unique_ptr<user_op::OpKernel> kernel = reg->create();
    user_op::KernelInitContext init;
    std::shared_ptr<user_op::OpKernelState> state = kernel->CreateOpKernelState(&init);
    for (int r = 0; r < repeat; ++r) kernel->Compute(&ctx, state.get(), nullptr);  // state reused across calls
    CUDA_OK(cudaStreamSynchronize(cs));
  };
  auto max_err = [](const std::vector<float>& got, const std::vector<double>& want) {
    double e = 0;
    for (size_t i = 0; i < got.size(); ++i) e = std::max(e, std::fabs(got[i] - want[i]) / (1.0 + std::fabs(want[i])));
    return e;
  };

  run("spmm_csr", "b", {{"a_crow", &t_crow}, {"a_col", &t_col}, {"a_val", &t_val}, {"b", &t_b}, {"out", &t_out}}, 1);
  std::vector<float> got(C.size());
  t_out.ToHost(got.data());
  EXPECT(max_err(got, C) < 2e-5);

  // fused_spmm_csr_bias_act: out = relu(A·b + bias) in one pass (the mock context answers every bool
  // attr with its `atomic` member, so atomic = true here means relu = true)
  std::vector<float> bias(N);
  for (int64_t j = 0; j < N; ++j) bias[j] = 0.03f * static_cast<float>(j) - 1.f;
  DevTensor t_bias(Shape({N}), kFloat, bias.data(), bias.size() * 4);
  DevTensor t_fused(Shape({M, N}), kFloat, nullptr, C.size() * 4);
  run("fused_spmm_csr_bias_act", "b", {{"a_crow", &t_crow}, {"a_col", &t_col}, {"a_val", &t_val}, {"b", &t_b}, {"bias", &t_bias},
                                      {"out", &t_fused}}, 1, /*relu=*/true);
  std::vector<double> Cf(C.size());
  for (int64_t i = 0; i < M; ++i)
    for (int64_t j = 0; j < N; ++j) Cf[i * N + j] = std::max(0.0, C[i * N + j] + static_cast<double>(bias[j]));
  t_fused.ToHost(got.data());
  EXPECT(max_err(got, Cf) < 2e-5);

  run("spmm_csr_grad_b", "dy", {{"a_crow", &t_crow}, {"a_col", &t_col}, {"a_val", &t_val}, {"dy", &t_dy}, {"db", &t_db}}, 2);
  got.resize(dB.size());
  t_db.ToHost(got.data());
  EXPECT(max_err(got, dB) < 2e-5);
  EXPECT(device.live == 0);  // nothing is ever allocated through the stream's device: no hidden op state
  // atomic route (attr)
  run("spmm_csr_grad_b", "dy", {{"a_crow", &t_crow}, {"a_col", &t_col}, {"a_val", &t_val}, {"dy", &t_dy}, {"db", &t_db}}, 1, true);
  t_db.ToHost(got.data());
  EXPECT(max_err(got, dB) < 2e-5);
  // cached-structure route: csr_transpose_structure once, then spmm_csr_grad_b with the three optional
  // inputs; the values are re-gathered on every call, so an in-place update of a_val is seen
  DevTensor t_tcrow(Shape({K + 1}), kInt32, nullptr, (K + 1) * 4), t_tcol(Shape({nnz}), kInt32, nullptr, nnz * 4);
  DevTensor t_tperm(Shape({nnz}), kInt32, nullptr, nnz * 4);
  run("csr_transpose_structure", "a_col", {{"a_crow", &t_crow}, {"a_col", &t_col}, {"t_crow", &t_tcrow}, {"t_col", &t_tcol},
                                          {"t_perm", &t_tperm}}, 1);
  std::map<std::string, user_op::Tensor*> cached = {{"a_crow", &t_crow}, {"a_col", &t_col}, {"a_val", &t_val}, {"dy", &t_dy},
                                                    {"db", &t_db}, {"t_crow", &t_tcrow}, {"t_col", &t_tcol}, {"t_perm", &t_tperm}};
  run("spmm_csr_grad_b", "dy", cached, 1);
  t_db.ToHost(got.data());
  EXPECT(max_err(got, dB) < 2e-5);
  std::vector<float> val2(val);
  for (auto& v : val2) v = -2.f * v;
  CUDA_OK(cudaMemcpy(t_val.mut_raw_dptr(), val2.data(), val2.size() * 4, cudaMemcpyHostToDevice));   // same pointer, new values
  run("spmm_csr_grad_b", "dy", cached, 1);
  t_db.ToHost(got.data());
  std::vector<double> dB2(dB);
  for (auto& x : dB2) x *= -2.0;
  EXPECT(max_err(got, dB2) < 4e-5);
  CUDA_OK(cudaMemcpy(t_val.mut_raw_dptr(), val.data(), val.size() * 4, cudaMemcpyHostToDevice));

  run("sddmm_csr", "b", {{"a_crow", &t_crow}, {"a_col", &t_col}, {"dy", &t_dy}, {"b", &t_b}, {"dval", &t_dval}}, 1);
  got.resize(dval.size());
  t_dval.ToHost(got.data());
  EXPECT(max_err(got, dval) < 2e-5);
  CUDA_OK(cudaStreamDestroy(cs));
  std::printf("glue gpu checks ok: spmm_csr / fused_spmm_csr_bias_act / spmm_csr_grad_b / sddmm_csr through OpKernel::Compute, "
              "%llu library launches\n", static_cast<unsigned long long>(ofspmm_launch_count()));
}

}  // namespace

int main(int argc, char** argv) {
  const std::string mode = argc > 1 ? argv[1] : "host";
  HostChecks();
  AutogradChecks();
  if (mode == "gpu") GpuChecks();
  return 0;
}

// What oneflow_tblgen would emit from OneFlowUserOps.td.patch (cmake/op_schema.cmake:47-80): one
// class per op with the static inference hooks the .td flags ask for.
#pragma once
#include "oneflow/core/framework/framework.h"
namespace oneflow {
#define OF_MOCK_DECLARE_OP(Name)                                                                 \
  struct Name {                                                                                  \
    static Maybe<void> InferLogicalTensorDesc(user_op::InferContext* ctx);                       \
    static Maybe<void> InferPhysicalTensorDesc(user_op::InferContext* ctx);                      \
    static Maybe<void> InferDataType(user_op::InferContext* ctx);                                \
    static Maybe<void> GetSbp(user_op::SbpContext* ctx);                                         \
    using GetInputArgModifier = user_op::GetInputArgModifier;                                    \
    static Maybe<void> ModifyInputArg(const GetInputArgModifier&, const user_op::UserOpConfWrapper&); \
  };
OF_MOCK_DECLARE_OP(SpmmCsrOp)
OF_MOCK_DECLARE_OP(SpmmCsrGradBOp)
OF_MOCK_DECLARE_OP(SddmmCsrOp)
OF_MOCK_DECLARE_OP(CsrTransposeStructureOp)
OF_MOCK_DECLARE_OP(FusedSpmmCsrBiasActOp)
#undef OF_MOCK_DECLARE_OP
}  // namespace oneflow

// stand-in for oneflow/core/kernel/cuda_graph_support.h:28-42
#pragma once
#include "oneflow/core/framework/framework.h"
namespace oneflow { namespace user_op {
class CudaGraphSupport {
 public:
  virtual ~CudaGraphSupport() = default;
  virtual bool IsCudaGraphSupported(KernelInitContext*, OpKernelState*) const { return true; }
};
} }

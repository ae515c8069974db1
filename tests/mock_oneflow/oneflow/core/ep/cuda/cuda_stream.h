// stand-in for oneflow/core/ep/cuda/cuda_stream.h:86-99
#pragma once
#include "oneflow/core/framework/framework.h"
struct CUstream_st;
namespace oneflow { namespace ep {
class CudaStream : public Stream {
 public:
  CudaStream(CUstream_st* s, Device* d) : stream_(s), device_(d) {}
  CUstream_st* cuda_stream() const { return stream_; }
  Device* device() const override { return device_; }
 private:
  CUstream_st* stream_;
  Device* device_;
};
} }

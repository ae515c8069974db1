// stand-in: see oneflow/core/framework/autograd_mock.h
#pragma once
#include "oneflow/core/framework/autograd_mock.h"

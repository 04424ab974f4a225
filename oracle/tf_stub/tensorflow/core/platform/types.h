// TEST INFRASTRUCTURE ONLY -- stand-in for tensorflow/core/platform/types.h (see ../framework/op_kernel.h).
#pragma once
#include <cstdint>
namespace tensorflow {
typedef long long int64;
typedef int int32;
}  // namespace tensorflow

// Helpers shared by the translation units of libcpmpc.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

#include <string>

namespace cpmpc {
namespace detail {

// Records the thread-local error message returned by cpmpc_last_error() and passes `code` through.
int Fail(int code, const std::string& msg);
// Adds to the counter behind cpmpc_launch_count().
void CountLaunches(int n);

}  // namespace detail
}  // namespace cpmpc

#define CPMPC_CUDA_TRY(expr)                                                                                   \
  do {                                                                                                         \
    const cudaError_t err_ = (expr);                                                                           \
    if (err_ != cudaSuccess) {                                                                                 \
      return ::cpmpc::detail::Fail(CPMPC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(err_));      \
    }                                                                                                          \
  } while (0)

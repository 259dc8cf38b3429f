// msda_host.h — host-side helpers shared by the launch code (defined in msda_capi.cu).
#pragma once

#include <cuda_runtime.h>

namespace msda {

// Records a printf-style message for msda_last_error() on this thread and returns `code`.
int fail(int code, const char* fmt, ...);
// MSDA_OK, or MSDA_ERR_CUDA with the runtime's message.
int check_cuda(cudaError_t e, const char* what);
// Counts the launch (msda_launch_count) and reports a launch-time error, if any.
int after_launch(const char* what);

}  // namespace msda

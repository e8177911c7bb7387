// Error reporting and process-wide counters of the b200md C ABI.
#include <stdio.h>

#include "common.cuh"

namespace b200md {
unsigned long long g_launch_count = 0;
static thread_local char g_cuda_error_text[256] = "";

int set_cuda_error(cudaError_t e) {
  snprintf(g_cuda_error_text, sizeof(g_cuda_error_text), "CUDA error %d: %s", (int)e,
           cudaGetErrorString(e));
  return B200MD_ERR_CUDA;
}
}  // namespace b200md

extern "C" {

const char* b200md_error_string(int code) {
  switch (code) {
    case B200MD_OK: return "ok";
    case B200MD_ERR_INVALID_ARGUMENT: return "invalid argument";
    case B200MD_ERR_CUDA: return b200md::g_cuda_error_text[0] ? b200md::g_cuda_error_text : "CUDA error";
    case B200MD_ERR_UNSUPPORTED: return "configuration not supported by the sm_100a kernels";
    case B200MD_ERR_WORKSPACE_TOO_SMALL: return "workspace too small";
    default: return "unknown error";
  }
}

int b200md_abi_version(void) { return B200MD_ABI_VERSION; }

uint64_t b200md_launch_count(void) { return b200md::g_launch_count; }

}  // extern "C"

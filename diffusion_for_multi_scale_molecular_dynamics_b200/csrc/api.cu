// Error reporting and process-wide counters of the b200md C ABI.
#include <stdio.h>

#include <vector>

#include "common.cuh"

namespace b200md {
unsigned long long g_launch_count = 0;
static thread_local char g_cuda_error_text[256] = "";

bool g_profile_on = false;
struct ProfileRecord { int cls; cudaEvent_t e0, e1; };
static std::vector<ProfileRecord> g_profile_records;

void profile_begin(int, cudaStream_t st, cudaEvent_t* e0) {
  cudaEventCreate(e0);
  cudaEventRecord(*e0, st);
}
void profile_end(int cls, cudaStream_t st, cudaEvent_t e0) {
  cudaEvent_t e1;
  cudaEventCreate(&e1);
  cudaEventRecord(e1, st);
  g_profile_records.push_back({cls, e0, e1});
}

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

int b200md_profile_enable(int32_t on) {
  b200md::g_profile_on = on != 0;
  return B200MD_OK;
}

int b200md_profile_collect(double* ms, int64_t* launches, int32_t n_classes) {
  if (!ms || !launches || n_classes < B200MD_KERNEL_CLASSES) return B200MD_ERR_INVALID_ARGUMENT;
  for (auto& r : b200md::g_profile_records) {
    cudaError_t e = cudaEventSynchronize(r.e1);
    if (e != cudaSuccess) return b200md::set_cuda_error(e);
    float t = 0.f;
    cudaEventElapsedTime(&t, r.e0, r.e1);
    if (r.cls >= 0 && r.cls < n_classes) { ms[r.cls] += t; launches[r.cls] += 1; }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  b200md::g_profile_records.clear();
  return B200MD_OK;
}

}  // extern "C"

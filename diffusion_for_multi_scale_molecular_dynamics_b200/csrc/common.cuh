// Shared helpers for the b200md kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200md.h"

namespace b200md {

extern unsigned long long g_launch_count;  // host-side counter, see api.cu

#define B200MD_CHECK_LAUNCH()                                       \
  do {                                                              \
    ::b200md::g_launch_count++;                                     \
    cudaError_t e__ = cudaGetLastError();                           \
    if (e__ != cudaSuccess) return ::b200md::set_cuda_error(e__);   \
  } while (0)

int set_cuda_error(cudaError_t e);

// In-stream timing of one launch (see b200md_profile_enable): records an event pair around the scope when
// profiling is on; costs one predictable branch when it is off.
extern bool g_profile_on;
void profile_begin(int kernel_class, cudaStream_t st, cudaEvent_t* e0);
void profile_end(int kernel_class, cudaStream_t st, cudaEvent_t e0);
struct ProfileScope {
  int cls; cudaStream_t st; cudaEvent_t e0; bool on;
  ProfileScope(int kernel_class, cudaStream_t stream) : cls(kernel_class), st(stream), e0(nullptr), on(g_profile_on) {
    if (on) profile_begin(cls, st, &e0);
  }
  ~ProfileScope() { if (on) profile_end(cls, st, e0); }
};

static inline cudaStream_t as_stream(b200md_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Function attributes (the opt-in to more than 48 KB of dynamic shared memory) and the SM count belong to a DEVICE, not to
// the process: a "done once" flag must be kept per device ordinal.  `PerDeviceOnce once; if (once.needed()) { ...set...;
// once.done(); }` -- setting twice from two threads is harmless, skipping a device is not.
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}
struct PerDeviceOnce {
  static constexpr int kMaxDevices = 64;
  volatile bool flag[kMaxDevices] = {};
  bool needed() const { const int d = current_device(); return d < 0 || d >= kMaxDevices || !flag[d]; }
  void done() { const int d = current_device(); if (d >= 0 && d < kMaxDevices) flag[d] = true; }
};

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011).  Counter-based: value = f(key, counter), no state.
// ---------------------------------------------------------------------------------------------
struct Philox4 {
  uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                          uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n1 = lo1;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    uint32_t n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
  return o;
}

// Counter layout shared by every kernel that draws noise (documented in DESIGN.md):
//   c0 = unit * ceil(unit_size/4) + lane/4,  c1 = low 32 bits of the GLOBAL sample index,
//   c2 = step_id,  c3 = draw_id | (high 24 bits of the global sample index << 8);  key = seed.
__host__ __device__ __forceinline__ Philox4 draw_block(uint64_t seed, uint64_t global_sample,
                                                       uint32_t block_in_sample, uint32_t step_id,
                                                       uint32_t draw_id) {
  uint32_t c3 = (draw_id & 0xFFu) | ((uint32_t)(global_sample >> 32) << 8);
  return philox4x32_10(block_in_sample, (uint32_t)global_sample, step_id, c3, (uint32_t)seed,
                       (uint32_t)(seed >> 32));
}

// uniform in [0,1) with 24 bits, like torch.rand for float32
__host__ __device__ __forceinline__ float u01(uint32_t r) { return (float)(r >> 8) * 5.9604644775390625e-08f; }
// uniform in (0,1] for the Box-Muller radius (never 0)
__host__ __device__ __forceinline__ float u01_open(uint32_t r) {
  return ((float)(r >> 8) + 1.0f) * 5.9604644775390625e-08f;
}

// four standard normals from one Philox block (Box-Muller).  SFU logarithm and sine / cosine on a phase centred in
// [-pi, pi) (absolute error ~2^-21 there): the draws are statistical inputs, not parity outputs, and the accurate libm
// versions made the update kernels compute-bound instead of HBM-bound.
__device__ __forceinline__ void normal4(const Philox4& p, float out[4]) {
  const float r0 = sqrtf(-2.0f * __logf(u01_open(p.x)));
  const float r1 = sqrtf(-2.0f * __logf(u01_open(p.z)));
  float s0, c0, s1, c1;
  __sincosf(6.283185307179586f * (u01(p.y) - 0.5f), &s0, &c0);
  __sincosf(6.283185307179586f * (u01(p.w) - 0.5f), &s1, &c1);
  out[0] = r0 * c0; out[1] = r0 * s0; out[2] = r1 * c1; out[3] = r1 * s1;
}

__device__ __forceinline__ float gumbel_from_uniform(float u, float small_epsilon) {
  // -log(-log(clip(u, eps)))   generators/langevin_generator.py:100-107
  u = fmaxf(u, small_epsilon);
  return -logf(-logf(u));
}

// remainder(x,1) with the ==1 fix-up   utils/basis_transformations.py:117-118
__device__ __forceinline__ float wrap_unit(float v) {
  float r = v - floorf(v);  // == torch.remainder(v, 1.0) for finite v: fmod + sign fix gives the same value
  // torch.remainder(a, b) = a - b * floor(a / b); with b == 1 this is exact fp32 subtraction.
  if (r == 1.0f) r = 0.0f;
  return r;
}

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + expf(-v)); }

}  // namespace b200md

// Reciprocal throughput (cycles per warp instruction per scheduler) of the instructions of the TS epilogue on B200.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates scripts/pipe_rates.cu ; results: profiles/pipe_rates_r01j.txt
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(float* out, long long* cyc, float seed) {
  float x[16];
  uint32_t h[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { x[i] = seed + i * 0.001f + threadIdx.x * 1e-4f; h[i] = __float_as_uint(x[i]); }
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 64; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (OP == 0) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "+r"(h[i]) : "f"(x[i]), "f"(__uint_as_float(h[i])));
      if (OP == 1) asm volatile("{.reg .b16 a, b; mov.b32 {a, b}, %1; sub.rn.f32.f16 %0, a, %2;}" : "=f"(x[i]) : "r"(h[i]), "f"(x[i]));
      if (OP == 2) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(x[i]) : "f"(x[i]));
      if (OP == 3) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(x[i]) : "f"(x[i]));
      if (OP == 4) asm volatile("{.reg .b64 t; mov.b64 t, {%0, %1}; add.rn.f32x2 t, t, t; mov.b64 {%0, %1}, t;}" : "+f"(x[i]), "+f"(x[(i + 8) & 15]));
      if (OP == 5) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[i]) : "f"(seed));
      if (OP == 6) asm volatile("{.reg .b16 a, b; mov.b32 {a, b}, %0; cvt.f32.f16 %0, a;}" : "+r"(h[i]));
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP>
void run(const char* name) {
  float* out; long long* cyc;
  cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 8);
  for (int warps : {4, 8, 16}) {   // 1, 2, 4 warps per scheduler
    k<OP><<<1, warps * 32>>>(out, cyc, 1.0f);
    k<OP><<<1, warps * 32>>>(out, cyc, 1.0f);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-10s warps/scheduler %d: %.2f cycles per warp instruction per scheduler\n", name, warps / 4, (double)c / (64 * 16) / (warps / 4));
  }
}

int main() {
  run<0>("F2FP"); run<1>("FHADD"); run<2>("MUFU.EX2"); run<3>("MUFU.RCP"); run<4>("FADD2"); run<5>("FFMA"); run<6>("cvt.f32.f16");
  return 0;
}

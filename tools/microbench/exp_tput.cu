// Microbenchmark (diagnostic only): per-SM throughput of exp2 variants and packed fp32 math on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
#define ITERS 4096
template <int MODE>
__global__ void k(float* out, float seed) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = seed + i * 0.01f + threadIdx.x * 1e-4f;
  uint32_t u[8];
  for (int i = 0; i < 8; ++i) u[i] = __float_as_uint(a[i]);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); a[i] -= 1.0f; }
      if (MODE == 1) { asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i])); }
      if (MODE == 2) { asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i])); }
      if (MODE == 3) { a[i] = fmaf(a[i], 0.999f, 0.001f); }
      if (MODE == 4) {  // packed fp32x2 fma: 2 lanes of work per instruction
        unsigned long long x = (unsigned long long)u[i] | ((unsigned long long)u[i] << 32), y = x, z = x;
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(y), "l"(z));
        u[i] = (uint32_t)x ^ (uint32_t)(x >> 32);
      }
      if (MODE == 5) {  // Cody-Waite + degree-3 polynomial exp2 on the FMA/ALU pipes (FA4-style)
        float x = a[i];
        float fl = floorf(x);
        float f = x - fl;
        float p = fmaf(fmaf(fmaf(0.0555f, f, 0.2402f), f, 0.6931f), f, 1.0f);
        a[i] = __uint_as_float(__float_as_uint(p) + ((int)fl << 23)) - 1.5f;
      }
      if (MODE == 6) { asm volatile("cvt.rn.bf16x2.f32 %0, %1, %1;" : "=r"(u[i]) : "f"(a[i])); a[i] = __uint_as_float(u[i]); }
    }
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int ops_per_inst) {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps = 4; warps <= 16; warps *= 2) {
    k<MODE><<<148, warps * 32>>>(out, 0.5f);
    cudaEventRecord(e0);
    k<MODE><<<148, warps * 32>>>(out, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst = (double)ITERS * 8 * warps * 32;  // thread-instructions per SM
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double cyc = ms * 1e-3 * clk * 1e3;
    printf("%-28s warps/SM %2d: %.2f lane-inst/clk/SM  (%.2f results/clk/SM)  [%.3f ms, err %s]\n", name, warps, inst / cyc,
           inst * ops_per_inst / cyc, ms, cudaGetErrorString(cudaGetLastError()));
  }
}
int main() {
  run<0>("ex2.approx.ftz.f32 (+fadd)", 1);
  run<1>("ex2.approx.ftz.bf16x2", 2);
  run<2>("ex2.approx.f16x2", 2);
  run<3>("ffma f32", 1);
  run<4>("fma.rn.f32x2 (+xor)", 2);
  run<5>("poly exp2 (floor+3 fma+shift)", 1);
  run<6>("cvt.rn.bf16x2.f32", 1);
  return 0;
}

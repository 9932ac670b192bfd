// Microbenchmark (diagnostic only): the attention kernel's per-tile softmax arithmetic (no TMEM, no MMA) with a
// fraction of the exponentials moved from MUFU to a packed-fp32x2 polynomial on the FMA pipe.
// One thread = one row of 128 scores, 8 warps per CTA (two per SM sub-partition, like the two softmax warpgroups).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../video_blade_b200/csrc/ptx.cuh"
using namespace blade;

// POLY = number of pairs out of every 8 pairs that use the polynomial
template <int POLY>
__global__ void __launch_bounds__(256) k(uint32_t* out, float seed, int tiles, long long* cyc) {
  uint32_t s[4][32];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int i = 0; i < 32; ++i) s[c][i] = __float_as_uint(seed * (c * 32 + i) * 0.01f - threadIdx.x * 1e-3f);
  float m = 0.f, l = 0.f;
  uint32_t chk = 0;
  const float sl2 = 0.1275f;
  const long long t0 = clock64();
  for (int it = 0; it < tiles; ++it) {
    float mxc[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int i = 0; i < 32; i += 2)
        mxc[c] = fmaxf(mxc[c], fmaxf(__uint_as_float(s[c][i]), __uint_as_float(s[c][i + 1])));
    const float mx = fmaxf(fmaxf(mxc[0], mxc[1]), fmaxf(mxc[2], mxc[3])) * sl2;
    m = fmaxf(m, mx);
    const float neg_m = -m;
    const uint64_t sl2_2 = pack_f32x2(sl2, sl2), negm_2 = pack_f32x2(neg_m, neg_m);
    uint64_t ls2[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const uint64_t x = fma_f32x2(pack_u32x2(s[c][2 * i], s[c][2 * i + 1]), sl2_2, negm_2);
        float p0, p1;
        if ((i & 7) < POLY) {
          ex2_poly_x2(x, p0, p1);
        } else {
          p0 = ex2_approx(lo_f32(x));
          p1 = ex2_approx(hi_f32(x));
        }
        ls2[i & 3] = add_f32x2(ls2[i & 3], pack_f32x2(p0, p1));
        pk[i] = pack_bf16x2(p0, p1);
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        chk ^= pk[i];
        // feed something loop-carried back into the scores so that nothing is hoisted
        s[c][2 * i] ^= (pk[i] & 0x00010000u);
      }
    }
    const uint64_t lsa = add_f32x2(add_f32x2(ls2[0], ls2[1]), add_f32x2(ls2[2], ls2[3]));
    l += lo_f32(lsa) + hi_f32(lsa);
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = chk ^ __float_as_uint(l) ^ __float_as_uint(m);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int POLY>
void run(int warps) {
  uint32_t* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 256 * 4);
  cudaMalloc(&cyc, 8);
  const int tiles = 2000;
  k<POLY><<<148, warps * 32>>>(out, 0.37f, 10, cyc);
  k<POLY><<<148, warps * 32>>>(out, 0.37f, tiles, cyc);
  long long h = 0;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("poly pairs %d/8, %d warps/SM: %.0f cycles per 128x128 tile per warpgroup-warp  (%s)\n", POLY, warps,
         (double)h / tiles, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
  cudaFree(cyc);
}
int main() {
  for (int warps : {4, 8}) {
    run<0>(warps);
    run<1>(warps);
    run<2>(warps);
    run<3>(warps);
    run<4>(warps);
  }
  return 0;
}

// scaffold_kernels.cu -- NOT the hot path: three fused element-wise kernels for the token-wise glue of the random-init DiT
// scaffolds (video_blade_b200/dit.py) that bench_clip.py / bench.py's `clip` time.  In the reference these ops belong
// to diffusers' WanTransformerBlock (adaLN modulation in fp32 around every sub-layer); eager PyTorch runs each of them as
// 5-7 passes over [B,S,C] in fp32 (28-30 bytes per element), which made the glue -- not attention, not the GEMMs -- the
// largest share of the clip (profiles/r02_profile_clip_before_fused_glue.txt: 48 % of a block).  One pass each here.
//   blade_scaffold_ln_modulate     out = bf16( LayerNorm(float(x))[* w + b] * (1 + scale[b]) + shift[b] )
//   blade_scaffold_gated_residual  out = bf16( float(x) + float(y) * gate[b] )
//   blade_scaffold_rmsnorm         out = bf16( float(x) * rsqrt(mean(float(x)^2) + eps) * float(w) )
// x, y, out: [B, S, C] contiguous 16-bit; scale / shift / gate: fp32 [B, C]; C a multiple of 256, C <= 4096.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace blade {

template <bool IS_BF16>
__device__ __forceinline__ void sc_unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (IS_BF16) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    } else {
      const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
}
template <bool IS_BF16>
__device__ __forceinline__ uint4 sc_pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (IS_BF16) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    } else {
      __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

constexpr int kMaxChunks = 16;  // 16-byte chunks per lane: C <= 32 * 16 * 8 = 4096

// one warp per token row; MODE 0 = LayerNorm (no affine) + modulate, 1 = RMSNorm * weight
template <bool IS_BF16, int MODE, int NCH>
__global__ void __launch_bounds__(256) norm_rows_kernel(const uint16_t* __restrict__ x, const float* __restrict__ scale,
                                                        const float* __restrict__ shift, const uint16_t* __restrict__ w16,
                                                        const uint16_t* __restrict__ b16, uint16_t* __restrict__ out,
                                                        int64_t rows, int S, int C, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  if (row >= rows) return;
  const int nch = C / 256;  // chunks of 8 elements per lane (<= NCH; NCH is the compile-time bound that sizes v[])
  const uint4* src = reinterpret_cast<const uint4*>(x + row * C);
  float v[NCH][8];
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (c < nch) {
      sc_unpack8<IS_BF16>(__ldg(src + c * 32 + lane), v[c]);
#pragma unroll
      for (int i = 0; i < 8; ++i) sum += MODE == 0 ? v[c][i] : v[c][i] * v[c][i];
    }
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  float mean = 0.f, rstd;
  if (MODE == 0) {
    mean = sum / static_cast<float>(C);
    float var = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
      if (c < nch)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float d = v[c][i] - mean;
          var = fmaf(d, d, var);
        }
#pragma unroll
    for (int off = 16; off; off >>= 1) var += __shfl_xor_sync(0xffffffffu, var, off);
    rstd = rsqrtf(var / static_cast<float>(C) + eps);
  } else {
    rstd = rsqrtf(sum / static_cast<float>(C) + eps);
  }
  const int64_t b = row / S;
  uint4* dst = reinterpret_cast<uint4*>(out + row * C);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (c < nch) {
      const int col = (c * 32 + lane) * 8;
      float o[8];
      if (MODE == 0) {
        const float4* sp = reinterpret_cast<const float4*>(scale + b * C + col);
        const float4* hp = reinterpret_cast<const float4*>(shift + b * C + col);
        const float4 s0 = __ldg(sp), s1 = __ldg(sp + 1), h0 = __ldg(hp), h1 = __ldg(hp + 1);
        const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
        float wf[8], bf_[8];
        if (w16) {  // affine LayerNorm (CogVideoX's LayerNormZero): ((x - mean) * rstd * w + b), then the modulation
          sc_unpack8<IS_BF16>(__ldg(reinterpret_cast<const uint4*>(w16 + col)), wf);
          if (b16) {
            sc_unpack8<IS_BF16>(__ldg(reinterpret_cast<const uint4*>(b16 + col)), bf_);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) bf_[i] = 0.f;
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float n = (v[c][i] - mean) * rstd;
          if (w16) n = fmaf(n, wf[i], bf_[i]);
          o[i] = fmaf(n, 1.0f + sc[i], sh[i]);
        }
      } else {
        float wf[8];
        sc_unpack8<IS_BF16>(__ldg(reinterpret_cast<const uint4*>(w16 + col)), wf);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = v[c][i] * rstd * wf[i];
      }
      dst[c * 32 + lane] = sc_pack8<IS_BF16>(o);
    }
  }
}

template <bool IS_BF16>
__global__ void __launch_bounds__(256) gated_residual_kernel(const uint16_t* __restrict__ x, const uint16_t* __restrict__ y,
                                                             const float* __restrict__ gate, uint16_t* __restrict__ out,
                                                             int64_t chunks, int64_t chunks_per_batch, int C) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (e >= chunks) return;
  const int64_t b = e / chunks_per_batch;
  const int col = static_cast<int>((e * 8) % C);
  float xf[8], yf[8];
  sc_unpack8<IS_BF16>(__ldg(reinterpret_cast<const uint4*>(x) + e), xf);
  sc_unpack8<IS_BF16>(__ldg(reinterpret_cast<const uint4*>(y) + e), yf);
  const float4* gp = reinterpret_cast<const float4*>(gate + b * C + col);
  const float4 g0 = __ldg(gp), g1 = __ldg(gp + 1);
  const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = fmaf(yf[i], g[i], xf[i]);
  reinterpret_cast<uint4*>(out)[e] = sc_pack8<IS_BF16>(o);
}

static int check_scaffold(const void* a, int64_t B, int64_t S, int64_t C, int dtype) {
  BLADE_REQUIRE(a, BLADE_ERR_ARG, "null pointer");
  BLADE_REQUIRE((reinterpret_cast<uintptr_t>(a) & 15) == 0, BLADE_ERR_ALIGN, "pointer not 16B aligned");
  BLADE_REQUIRE(B >= 1 && S >= 1 && C >= 256 && C % 256 == 0 && C <= 256 * kMaxChunks, BLADE_ERR_SHAPE,
                "scaffold kernels need C a multiple of 256, <= %d (got %lld)", 256 * kMaxChunks, (long long)C);
  BLADE_REQUIRE(dtype == BLADE_BF16 || dtype == BLADE_F16, BLADE_ERR_DTYPE, "bf16 / f16 only");
  return BLADE_OK;
}

}  // namespace blade

using namespace blade;

extern "C" int blade_scaffold_ln_modulate(const void* x, const float* scale, const float* shift, const void* weight,
                                          const void* bias, void* out, int64_t B, int64_t S, int64_t C, float eps,
                                          int32_t dtype, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int e = check_scaffold(x, B, S, C, dtype)) return e;
  BLADE_REQUIRE(scale && shift && out, BLADE_ERR_ARG, "null pointer");
  const int64_t rows = B * S;
  const unsigned grid = static_cast<unsigned>(ceil_div(rows, 8));
#define LAUNCH_LN(BF, N_)                                                                                          \
  norm_rows_kernel<BF, 0, N_><<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(x), scale, shift,            \
                                                        static_cast<const uint16_t*>(weight),                     \
                                                        static_cast<const uint16_t*>(bias),                       \
                                                        static_cast<uint16_t*>(out), rows, (int)S, (int)C, eps)
  const bool bf = dtype == BLADE_BF16;
  const int nch = (int)(C / 256);
  if (nch <= 6) { if (bf) LAUNCH_LN(true, 6); else LAUNCH_LN(false, 6); }
  else if (nch <= 12) { if (bf) LAUNCH_LN(true, 12); else LAUNCH_LN(false, 12); }
  else { if (bf) LAUNCH_LN(true, kMaxChunks); else LAUNCH_LN(false, kMaxChunks); }
#undef LAUNCH_LN
  BLADE_CUDA_OK(cudaGetLastError());
  return BLADE_OK;
}

extern "C" int blade_scaffold_rmsnorm(const void* x, const void* weight, void* out, int64_t rows, int64_t C, float eps,
                                      int32_t dtype, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int e = check_scaffold(x, 1, rows, C, dtype)) return e;
  BLADE_REQUIRE(weight && out, BLADE_ERR_ARG, "null pointer");
  const unsigned grid = static_cast<unsigned>(ceil_div(rows, 8));
#define LAUNCH_RMS(BF, N_)                                                                                         \
  norm_rows_kernel<BF, 1, N_><<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(x), nullptr, nullptr,        \
                                                        static_cast<const uint16_t*>(weight), nullptr,            \
                                                        static_cast<uint16_t*>(out), rows, (int)rows, (int)C, eps)
  const bool bf = dtype == BLADE_BF16;
  const int nch = (int)(C / 256);
  if (nch <= 6) { if (bf) LAUNCH_RMS(true, 6); else LAUNCH_RMS(false, 6); }
  else if (nch <= 12) { if (bf) LAUNCH_RMS(true, 12); else LAUNCH_RMS(false, 12); }
  else { if (bf) LAUNCH_RMS(true, kMaxChunks); else LAUNCH_RMS(false, kMaxChunks); }
#undef LAUNCH_RMS
  BLADE_CUDA_OK(cudaGetLastError());
  return BLADE_OK;
}

extern "C" int blade_scaffold_gated_residual(const void* x, const void* y, const float* gate, void* out, int64_t B,
                                             int64_t S, int64_t C, int32_t dtype, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int e = check_scaffold(x, B, S, C, dtype)) return e;
  BLADE_REQUIRE(y && gate && out, BLADE_ERR_ARG, "null pointer");
  const int64_t chunks = B * S * C / 8;
  const unsigned grid = static_cast<unsigned>(ceil_div(chunks, 256));
  if (dtype == BLADE_BF16)
    gated_residual_kernel<true><<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(x), static_cast<const uint16_t*>(y),
                                                          gate, static_cast<uint16_t*>(out), chunks, S * C / 8, (int)C);
  else
    gated_residual_kernel<false><<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(x), static_cast<const uint16_t*>(y),
                                                           gate, static_cast<uint16_t*>(out), chunks, S * C / 8, (int)C);
  BLADE_CUDA_OK(cudaGetLastError());
  return BLADE_OK;
}

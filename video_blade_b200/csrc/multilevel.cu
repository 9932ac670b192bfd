// multilevel.cu -- the reference's experimental multi-level pooled sparse attention on B200.
//
// Reference (paths under cogvideox/sample_evaluate/Triton/): module cogvideo_newattn.py (N), Triton kernels
// kernels/block_sparse_attn_kernel_with_backward_9_10.py (K9).  A block-mask entry is a LEVEL: 0 = skip the (query
// block, key block) pair, 1 = attend the block's 128 keys, L in {2,4,8} = attend its 128/L mean-pooled keys and
// values with `+ log L` on the scaled score, everything inside ONE softmax per query row (K9:135-277, 339-692).
//
//   pyramid_kernel          K/V -> levels 2, 4, 8 by rounds of pair means over the replicate-padded tensor, each
//                           round rounded to the tensor dtype like the reference's successive `pooling` calls
//                           (K9:1252-1270, 1307-1316).  HBM-bound: reads K and V once, writes 7/8 of that.
//   multilevel_mask_kernel  transfer_attn_to_mask (N:154-207): per score row, rank the key blocks (value desc,
//                           index asc == torch.sort stable) with the register bitonic network, level = the rank
//                           range's level, last `force_last` rows / columns forced to level 1 (N:201-203); emits
//                           the level mask, the per-level entry counts and the row's list sorted by (level, block).
//   level_mask_to_index_kernel   the same list format from a caller-provided level mask
//                           (sparse_attention_fn(q,k,v,mask), K9:1578-1611).
//   the attention itself    asa_multilevel_attn_kernel in attn_kernel.cu (the block-sparse pipeline with tiles
//                           assembled from the pyramid and `+ log2 L` in the softmax).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>

#include "common.cuh"

namespace blade {

template <bool IS_BF16>
__device__ __forceinline__ float rt(float x) {
  return IS_BF16 ? __bfloat162float(__float2bfloat16_rn(x)) : __half2float(__float2half_rn(x));
}
template <bool IS_BF16>
__device__ __forceinline__ void unpack8m(const uint4& u, float (&f)[8]) {
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
__device__ __forceinline__ uint4 pack8m(const float (&f)[8]) {
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

struct PyrStrides {
  int64_t b, h, s;
};

// grid (ceil(groups / (256 / LPR)), H, B): one thread = 8 consecutive (padded) rows x 8 channels of K and of V
template <int D, bool IS_BF16>
__global__ void __launch_bounds__(256) pyramid_kernel(const uint16_t* __restrict__ k, const uint16_t* __restrict__ v,
                                                      PyrStrides sk, PyrStrides sv, int S, int H, int groups,
                                                      uint16_t* __restrict__ k2, uint16_t* __restrict__ v2,
                                                      uint16_t* __restrict__ k4, uint16_t* __restrict__ v4,
                                                      uint16_t* __restrict__ k8, uint16_t* __restrict__ v8) {
  constexpr int LPR = D / 8;
  const int g = blockIdx.x * (256 / LPR) + threadIdx.x / LPR;
  const int chunk = threadIdx.x % LPR;
  if (g >= groups) return;
  const int h = blockIdx.y, b = blockIdx.z;
  const int64_t bh = static_cast<int64_t>(b) * H + h;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const uint16_t* src = (t ? v : k) + b * (t ? sv.b : sk.b) + h * (t ? sv.h : sk.h);
    const int64_t ss = t ? sv.s : sk.s;
    uint4 raw[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      int row = g * 8 + r;
      row = row < S ? row : S - 1;  // replicate padding to the block multiple (K9:1239-1250)
      raw[r] = __ldg(reinterpret_cast<const uint4*>(src + row * ss) + chunk);
    }
    float x[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r) unpack8m<IS_BF16>(raw[r], x[r]);
    float p2[4][8], p4[2][8], p8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int r = 0; r < 4; ++r) p2[r][i] = rt<IS_BF16>((x[2 * r][i] + x[2 * r + 1][i]) * 0.5f);
#pragma unroll
      for (int r = 0; r < 2; ++r) p4[r][i] = rt<IS_BF16>((p2[2 * r][i] + p2[2 * r + 1][i]) * 0.5f);
      p8[i] = rt<IS_BF16>((p4[0][i] + p4[1][i]) * 0.5f);
    }
    uint16_t* o2 = (t ? v2 : k2) + (bh * (groups * 4) + g * 4) * D;
    uint16_t* o4 = (t ? v4 : k4) + (bh * (groups * 2) + g * 2) * D;
    uint16_t* o8 = (t ? v8 : k8) + (bh * groups + g) * D;
#pragma unroll
    for (int r = 0; r < 4; ++r) reinterpret_cast<uint4*>(o2 + r * D)[chunk] = pack8m<IS_BF16>(p2[r]);
#pragma unroll
    for (int r = 0; r < 2; ++r) reinterpret_cast<uint4*>(o4 + r * D)[chunk] = pack8m<IS_BF16>(p4[r]);
    reinterpret_cast<uint4*>(o8)[chunk] = pack8m<IS_BF16>(p8);
  }
}

// level code of a level value: 1,2,4,8 -> 0..3; anything else (0 = skip) -> -1
__device__ __forceinline__ int level_code(int lv) { return lv == 1 ? 0 : (lv == 2 ? 1 : (lv == 4 ? 2 : (lv == 8 ? 3 : -1))); }

// shared tail of the two list builders: lv[j] (smem, this warp's row, u8 levels) -> list sorted by (level, block id)
__device__ __forceinline__ void emit_level_lists(const uint8_t* lv, int nk, int lane, int32_t* irow, int32_t* c4row) {
  int base = 0;
  int counts[4];
#pragma unroll
  for (int lc = 0; lc < 4; ++lc) {
    const int start = base;
    for (int j0 = 0; j0 < nk; j0 += 32) {
      const int j = j0 + lane;
      const bool sel = j < nk && lv[j] == (1 << lc);
      const unsigned bal = __ballot_sync(0xffffffffu, sel);
      if (sel) irow[base + __popc(bal & ((1u << lane) - 1u))] = j;
      base += __popc(bal);
    }
    counts[lc] = base - start;
  }
  for (int j = base + lane; j < nk; j += 32) irow[j] = -1;
  if (lane == 0) *reinterpret_cast<int4*>(c4row) = make_int4(counts[0], counts[1], counts[2], counts[3]);
}

// one warp per score row, E elements per lane (nk <= 32 E <= 256); rank_level[p] = level of sorted position p
template <int E>
__global__ void __launch_bounds__(256) multilevel_mask_kernel(const float* __restrict__ scores, int64_t total_rows, int nq,
                                                              int nk, const uint8_t* __restrict__ rank_level,
                                                              int force_last, uint8_t* __restrict__ level_mask,
                                                              int32_t* __restrict__ idx, int32_t* __restrict__ cnt4) {
  constexpr int N = 32 * E;
  __shared__ uint8_t lvs[8][N];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  if (row >= total_rows) return;
  const int qi = static_cast<int>(row % nq);
  const float* src = scores + row * nk;
  // same 64-bit keys / network as select_bitonic_kernel (mask_kernels.cu): (value desc, index asc) total order
  uint64_t key[E];
#pragma unroll
  for (int r = 0; r < E; ++r) {
    const int p = lane * E + r;
    const float x = p < nk ? __ldg(src + p) + 0.0f : -INFINITY;
    const uint32_t b = __float_as_uint(x);
    const uint32_t mono = b ^ (static_cast<uint32_t>(static_cast<int32_t>(b) >> 31) | 0x80000000u);
    key[r] = (static_cast<uint64_t>(mono) << 32) | static_cast<uint32_t>(~p);
  }
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j >= 1; j >>= 1) {
      if (j >= E) {
        const bool keep_if_first = ((lane * E) & k) == 0 == (((lane * E) & j) == 0);
#pragma unroll
        for (int r = 0; r < E; ++r) {
          const uint64_t ok = __shfl_xor_sync(0xffffffffu, key[r], j / E);
          const bool keep = (key[r] > ok) == keep_if_first;
          key[r] = keep ? key[r] : ok;
        }
      } else {
#pragma unroll
        for (int r = 0; r < E; ++r) {
          if ((r & j) == 0) {
            const int r2 = r | j;
            const bool up = ((lane * E + r) & k) == 0;
            const bool swap = (key[r] > key[r2]) != up;
            const uint64_t t = key[r];
            key[r] = swap ? key[r2] : t;
            key[r2] = swap ? t : key[r2];
          }
        }
      }
    }
  }
  uint8_t* lv = lvs[warp];
  const bool full_row = force_last > 0 && qi >= nq - force_last;
#pragma unroll
  for (int r = 0; r < E; ++r) {
    const int p = lane * E + r;                                  // sorted position
    const int id = static_cast<int>(~static_cast<uint32_t>(key[r]));  // original block index
    if (id < nk) {
      int l = p < nk ? rank_level[p] : 0;
      if (full_row || (force_last > 0 && id >= nk - force_last)) l = 1;   // N:201-203
      lv[id] = static_cast<uint8_t>(l);
    }
  }
  __syncwarp();
  if (level_mask)
    for (int j = lane; j < nk; j += 32) level_mask[row * nk + j] = lv[j];
  emit_level_lists(lv, nk, lane, idx + row * nk, cnt4 + row * 4);
}

__global__ void __launch_bounds__(256) level_mask_to_index_kernel(const uint8_t* __restrict__ level_mask, int64_t total_rows,
                                                                  int nk, int32_t* __restrict__ idx, int32_t* __restrict__ cnt4) {
  extern __shared__ uint8_t sm_lv[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  if (row >= total_rows) return;
  uint8_t* lv = sm_lv + warp * nk;
  for (int j = lane; j < nk; j += 32) lv[j] = level_mask[row * nk + j];
  __syncwarp();
  emit_level_lists(lv, nk, lane, idx + row * nk, cnt4 + row * 4);
}

}  // namespace blade

using namespace blade;

extern "C" int blade_multilevel_pyramid(const BladeTensor* k, const BladeTensor* v, void* k2, void* v2, void* k4, void* v4,
                                        void* k8, void* v8, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int e = check_tensor16(k, "k")) return e;
  if (int e = check_tensor16(v, "v")) return e;
  BLADE_REQUIRE(k2 && v2 && k4 && v4 && k8 && v8, BLADE_ERR_ARG, "null pyramid output");
  for (int i = 0; i < 4; ++i) BLADE_REQUIRE(k->shape[i] == v->shape[i], BLADE_ERR_SHAPE, "k/v shapes differ");
  BLADE_REQUIRE(k->dtype == v->dtype, BLADE_ERR_DTYPE, "k/v dtypes differ");
  const int64_t B = k->shape[0], H = k->shape[1], S = k->shape[2], D = k->shape[3];
  BLADE_REQUIRE(H <= 65535 && B <= 65535, BLADE_ERR_SHAPE, "bad B/H");
  const int groups = static_cast<int>(ceil_div(S, 128) * 16);  // groups of 8 rows of the block-padded sequence
  PyrStrides sk{k->stride[0], k->stride[1], k->stride[2]}, sv{v->stride[0], v->stride[1], v->stride[2]};
  const int gpb = 256 / static_cast<int>(D / 8);
  dim3 grid(static_cast<unsigned>(ceil_div(groups, gpb)), static_cast<unsigned>(H), static_cast<unsigned>(B));
#define LAUNCH_PYR(DD, BF)                                                                                            \
  pyramid_kernel<DD, BF><<<grid, 256, 0, stream>>>(                                                                   \
      static_cast<const uint16_t*>(k->ptr), static_cast<const uint16_t*>(v->ptr), sk, sv, (int)S, (int)H, groups,    \
      static_cast<uint16_t*>(k2), static_cast<uint16_t*>(v2), static_cast<uint16_t*>(k4), static_cast<uint16_t*>(v4), \
      static_cast<uint16_t*>(k8), static_cast<uint16_t*>(v8))
  const bool bf = k->dtype == BLADE_BF16;
  if (D == 128) { if (bf) LAUNCH_PYR(128, true); else LAUNCH_PYR(128, false); }
  else          { if (bf) LAUNCH_PYR(64, true); else LAUNCH_PYR(64, false); }
#undef LAUNCH_PYR
  BLADE_CUDA_OK(cudaGetLastError());
  return BLADE_OK;
}

extern "C" int blade_multilevel_mask(const float* scores, int64_t B, int64_t H, int64_t nq, int64_t nk,
                                     const uint8_t* rank_level, int32_t force_last, uint8_t* level_mask, int32_t* idx,
                                     int32_t* cnt4, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BLADE_REQUIRE(scores && rank_level && idx && cnt4, BLADE_ERR_ARG, "null pointer");
  BLADE_REQUIRE(nq >= 1 && nk >= 1 && nk <= 256, BLADE_ERR_SHAPE, "multi-level mask supports 1 <= nk <= 256 (got %lld)",
                (long long)nk);
  BLADE_REQUIRE((reinterpret_cast<uintptr_t>(cnt4) & 15) == 0, BLADE_ERR_ALIGN, "cnt4 not 16B aligned");
  const int64_t rows = B * H * nq;
  const unsigned grid = static_cast<unsigned>(ceil_div(rows, 8));
#define LAUNCH_ML(EE)                                                                                                 \
  multilevel_mask_kernel<EE><<<grid, 256, 0, stream>>>(scores, rows, (int)nq, (int)nk, rank_level, force_last,       \
                                                       level_mask, idx, cnt4)
  if (nk <= 32) LAUNCH_ML(1);
  else if (nk <= 64) LAUNCH_ML(2);
  else if (nk <= 128) LAUNCH_ML(4);
  else LAUNCH_ML(8);
#undef LAUNCH_ML
  BLADE_CUDA_OK(cudaGetLastError());
  return BLADE_OK;
}

extern "C" int blade_level_mask_to_index(const uint8_t* level_mask, int64_t B, int64_t H, int64_t nq, int64_t nk,
                                         int32_t* idx, int32_t* cnt4, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BLADE_REQUIRE(level_mask && idx && cnt4, BLADE_ERR_ARG, "null pointer");
  BLADE_REQUIRE(nq >= 1 && nk >= 1 && nk <= 4096, BLADE_ERR_SHAPE, "bad nq/nk");
  BLADE_REQUIRE((reinterpret_cast<uintptr_t>(cnt4) & 15) == 0, BLADE_ERR_ALIGN, "cnt4 not 16B aligned");
  const int64_t rows = B * H * nq;
  level_mask_to_index_kernel<<<static_cast<unsigned>(ceil_div(rows, 8)), 256, 8 * nk, stream>>>(level_mask, rows, (int)nk,
                                                                                                idx, cnt4);
  BLADE_CUDA_OK(cudaGetLastError());
  return BLADE_OK;
}

extern "C" int blade_multilevel_attn_fwd(const BladeTensor* q, const BladeTensor* k, const BladeTensor* v,
                                         const BladeTensor* k2, const BladeTensor* v2, const BladeTensor* k4,
                                         const BladeTensor* v4, const BladeTensor* k8, const BladeTensor* v8,
                                         const int32_t* idx, const int32_t* cnt4, int64_t idx_stride, BladeTensor* out,
                                         float* lse, const int32_t* dst_row, float softmax_scale, void* workspace,
                                         size_t ws_bytes, void* stream) {
  MultiLevelArgs ml{{k2, k4, k8}, {v2, v4, v8}, cnt4};
  return launch_attn(q, k, v, idx, nullptr, idx_stride, nullptr, nullptr, 0, out, lse, dst_row, softmax_scale, 0, workspace,
                     ws_bytes, static_cast<cudaStream_t>(stream), nullptr, &ml);
}

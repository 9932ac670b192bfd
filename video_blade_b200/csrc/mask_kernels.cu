// mask_kernels.cu -- north-star kernel (a): mask generation for Adaptive Sparse Attention.
//
//   prep_block_kernel   HBM-bound: one pass over Q,K,V rows (optionally gathered into Gilbert-curve
//                       order, W:142-152) producing curve-ordered copies and fp32 block means over the
//                       replicate-padded sequence (W:25-36).  16-byte vector loads, a half-warp per 256-byte
//                       row, 8 independent loads in flight per thread.
//   pool_kernel         gap-means of K and V for the global branch (simple_pooling, W:88-93).
//   score_meanpool_kernel  coarse nb x nb block-score GEMM on the means + row softmax (fp32).
//   select_bitonic_kernel / select_rank_kernel
//                       energy-threshold block selection (transfer_attn_to_mask, W:214-229 / C:228-248):
//                       one warp per score row; total order (value desc, index asc) by a register bitonic
//                       network (rank counting for rows > 256 blocks), fp64 sequential prefix sums rounded to
//                       fp32 (== torch CPU sort(stable)+cumsum, bit-exact), ballot compaction into the
//                       ascending per-row block-index list.
//   mask_to_index_kernel   bool mask -> the same index-list format.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace blade {

template <bool IS_BF16>
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (IS_BF16) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    } else {
      __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      float2 t = __half22float2(h);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
}
template <bool IS_BF16>
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
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

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

struct Strides3 {
  int64_t b, h, s;  // element strides
};

// BladeAsaConfig.select_rounding: 0 = fp32 scores; 1 = bf16, 2 = f16 (the reference keeps Po in the model dtype, W:214-221)
__device__ __forceinline__ float round_sel(float x, int mode) {
  if (mode == 1) return __bfloat162float(__float2bfloat16_rn(x));
  if (mode == 2) return __half2float(__float2half_rn(x));
  return x;
}
// The cut of transfer_attn_to_mask(mode="energy") over the descending-sorted row `sorted[0..nk)` (fp64 copies of the
// fp32 values): first index whose prefix sum reaches energy_threshold * total, exclusive (W:217-224); nk if none below
// `lim`.  Sequential on purpose: bit-exact with torch's CPU cumsum, which accumulates fp32 inputs in fp64 and 16-bit
// inputs in fp32, rounding every emitted prefix to the tensor dtype.
__device__ __forceinline__ int energy_cut(const double* sorted, int nk, int lim, float thr, int mode) {
  if (mode == 0) {
    double acc = 0.0;
    int i = 0;
#pragma unroll 1
    for (; i + 8 <= nk; i += 8) {
      double t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] = sorted[i + u];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += t[u];
    }
    for (; i < nk; ++i) acc += sorted[i];
    const float th = __fmul_rn(thr, static_cast<float>(acc));
    acc = 0.0;
    for (i = 0; i < lim; ++i) {
      acc += sorted[i];
      if (static_cast<float>(acc) >= th) return i;
    }
    return nk;
  }
  float acc = 0.f;
  for (int i = 0; i < nk; ++i) acc = __fadd_rn(acc, static_cast<float>(sorted[i]));
  const float th = round_sel(__fmul_rn(thr, round_sel(acc, mode)), mode);
  acc = 0.f;
  for (int i = 0; i < lim; ++i) {
    acc = __fadd_rn(acc, static_cast<float>(sorted[i]));
    if (round_sel(acc, mode) >= th) return i;
  }
  return nk;
}
// The same cut for fp32 scores (mode 0) by the whole warp: the prefix sums of the first `lim` elements ARE the first
// partial sums of the total, so lane 0 runs ONE sequential fp64 chain over the row (bit-exact with torch's CPU cumsum)
// and leaves the first min(lim, 64) partials, rounded to fp32, in `pre` (the row's flag scratch, not yet in use); the
// comparison against 0.95 * total then takes one ballot per 32 prefixes instead of a dependent
// add -> convert -> compare -> branch chain per element (12 % of the fused kernel's samples under ncu).
__device__ __forceinline__ int energy_cut_warp(const double* sorted, float* pre, int nk, int lim, float thr, int lane) {
  float th = 0.f;
  if (lane == 0) {
    double acc = 0.0;
    int i = 0;
#pragma unroll 1
    for (; i + 8 <= nk; i += 8) {
      double t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] = sorted[i + u];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        acc += t[u];
        if (i < 64) pre[i + u] = static_cast<float>(acc);   // warp-uniform per iteration of the outer loop
      }
    }
    for (; i < nk; ++i) {
      acc += sorted[i];
      if (i < 64) pre[i] = static_cast<float>(acc);
    }
    th = __fmul_rn(thr, static_cast<float>(acc));
  }
  th = __shfl_sync(0xffffffffu, th, 0);
  __syncwarp();
  const int lim64 = lim < 64 ? lim : 64;
  for (int b = 0; b < lim64; b += 32) {
    const int i = b + lane;
    const bool hit = i < lim64 && i < nk && pre[i] >= th;
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (m) return b + __ffs(m) - 1;
  }
  if (lim <= 64) return nk;
  // longer limits (not a configuration the reference uses): finish sequentially from the 64th prefix on
  int k = nk;
  if (lane == 0) {
    double acc = 0.0;
    for (int i = 0; i < lim && i < nk; ++i) {
      acc += sorted[i];
      if (i >= 64 && static_cast<float>(acc) >= th) { k = i; break; }
    }
  }
  return __shfl_sync(0xffffffffu, k, 0);
}

// ------------------------------------------------------------------------------------------------
// q/k RMSNorm statistic (Wan: rms_norm_across_heads, MW:99-102): rstd[token] = rsqrt(mean over all H*D channels of
// x^2 + eps) for q and k.  One warp per token row ([B,S,H*D] memory, a contiguous 2*H*D-byte segment), 16-byte loads.
// out: fp32 [2][B*S] (q then k).  The normalisation itself happens inside prep_block_kernel.
// ------------------------------------------------------------------------------------------------
// Ulysses push variant: the statistic of MY token shard goes to every peer's full-length [2][out_rows] table at row
// offset out_row0 (stores over NVLink; the caller's barrier publishes them)
struct StatPeers {
  float* out[BLADE_MAX_PEERS];
  int n;            // 0 = single destination `out`
  int64_t out_rows; // rows of a destination table (tokens of the whole sequence)
  int64_t out_row0; // first row I own
};

template <bool IS_BF16>
__global__ void __launch_bounds__(256) rms_stat_kernel(const uint16_t* __restrict__ q, const uint16_t* __restrict__ k,
                                                       int64_t q_sb, int64_t q_ss, int64_t k_sb, int64_t k_ss, int S,
                                                       int64_t rows, int hd, float eps, float* __restrict__ out,
                                                       const StatPeers peers) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  if (row >= rows) return;
  const int64_t b = row / S, sidx = row % S;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const uint16_t* src = t == 0 ? q + b * q_sb + sidx * q_ss : k + b * k_sb + sidx * k_ss;
    float acc = 0.f;
    for (int c = lane; c < hd / 8; c += 32) {
      float f[8];
      unpack8<IS_BF16>(ldg_stream(reinterpret_cast<const uint4*>(src) + c), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc = fmaf(f[i], f[i], acc);
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    const float r = rsqrtf(acc / static_cast<float>(hd) + eps);
    if (peers.n == 0) {
      if (lane == 0) out[t * rows + row] = r;
    } else if (lane < peers.n) {
      peers.out[lane][t * peers.out_rows + peers.out_row0 + row] = r;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// prep: grid (nb, H, B), 256 threads.  One CTA = one block of `block` output rows of one head.
// ------------------------------------------------------------------------------------------------
// Ulysses pull: token t of tensor j lives on peer t / rows at local row t % rows of that peer's [rows, H_total, D]
// projection output; base[j][p] already points at my first head's channels (BladePeers)
struct PeerSrc {
  const uint16_t* base[3][BLADE_MAX_PEERS];
  int rows;  // 0 = off
};

template <int D, bool IS_BF16, bool COPY, bool ROPE, bool NORM, bool PEER>
__global__ void __launch_bounds__(256, (ROPE || NORM) ? 3 : 0) prep_block_kernel(const uint16_t* __restrict__ q, const uint16_t* __restrict__ k,
                                                         const uint16_t* __restrict__ v, Strides3 sq, Strides3 sk,
                                                         Strides3 sv, const int32_t* __restrict__ src_row,
                                                         uint16_t* __restrict__ q_r, uint16_t* __restrict__ k_r,
                                                         uint16_t* __restrict__ v_r, float* __restrict__ q_mean,
                                                         float* __restrict__ k_mean, int S, int H, int nb, int block,
                                                         const float* __restrict__ rope, int rope_first,
                                                         const float* __restrict__ rstd, const uint16_t* __restrict__ wq,
                                                         const uint16_t* __restrict__ wk, int norm_kind,
                                                         const int32_t* __restrict__ tok_row,
                                                         const uint16_t* __restrict__ bq, const uint16_t* __restrict__ bk,
                                                         float norm_eps, const __grid_constant__ PeerSrc peer,
                                                         int tmask /* bit t: process tensor t of (q, k, v) */) {
  constexpr int LPR = D / 8;        // lanes per row (16-byte chunks)
  constexpr int RPW = 32 / LPR;     // rows per warp-wide load
  constexpr int RPP = 8 * RPW;      // rows per pass of the 8 warps
  constexpr int SLOTS = RPP;        // partial-sum slots
  __shared__ float red[2][SLOTS][D];

  const int blk = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / LPR, chunk = lane % LPR;
  const int slot = warp * RPW + sub;
  const int passes = block / RPP;   // host guarantees divisibility
  const int row0 = blk * block;
  const int64_t out_base = (static_cast<int64_t>(b) * H + h) * S;

  float acc[2][8];
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[t][i] = 0.f;

#pragma unroll
  for (int t = 0; t < 3; ++t) {
    const uint16_t* src = t == 0 ? q : (t == 1 ? k : v);
    const Strides3 st = t == 0 ? sq : (t == 1 ? sk : sv);
    uint16_t* dst = t == 0 ? q_r : (t == 1 ? k_r : v_r);
    if (!((tmask >> t) & 1)) continue;                // this launch handles a subset of (q, k, v)
    if (t == 2 && !(COPY && dst != nullptr)) break;  // V is only needed for the copy
    const uint16_t* base = src + b * st.b + h * st.h;
    for (int p0 = 0; p0 < passes; p0 += 8) {
      uint4 val[8];
      int rows[8];
      int srcs[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int p = p0 + u;
        int r = row0 + p * RPP + slot;
        rows[u] = r;
        const int rc = r < S ? r : S - 1;  // replicate padding (W:35)
        const int sr = src_row ? __ldg(src_row + rc) : rc;
        srcs[u] = ((ROPE || NORM) && tok_row) ? __ldg(tok_row + rc) : sr;  // token index (rotary table / norm statistic)
        if (p < passes) {
          const uint16_t* rowp;
          if (PEER) {  // compile-time: pull the row over NVLink from the peer that owns token sr
            const int pp = sr / peer.rows;
            rowp = peer.base[t][pp] + h * st.h + static_cast<int64_t>(sr - pp * peer.rows) * st.s;
          } else {
            rowp = base + sr * st.s;
          }
          val[u] = ldg_stream(reinterpret_cast<const uint4*>(rowp) + chunk);
        }
      }
      if (NORM && t < 2 && norm_kind == 3) {
        // CogVideoX: LayerNorm over the D channels of this head (MC:54-57); a row is spread over LPR lanes
        float w[8], bias[8];
        unpack8<IS_BF16>(__ldg(reinterpret_cast<const uint4*>(t == 0 ? wq : wk) + chunk), w);
        const uint16_t* bp = t == 0 ? bq : bk;
        if (bp) {
          unpack8<IS_BF16>(__ldg(reinterpret_cast<const uint4*>(bp) + chunk), bias);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) bias[i] = 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (p0 + u >= passes) continue;  // warp-uniform
          float f[8];
          unpack8<IS_BF16>(val[u], f);
          float sum = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) sum += f[i];
#pragma unroll
          for (int off = LPR / 2; off; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
          const float mean = sum * (1.0f / D);
          float var = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            f[i] -= mean;
            var = fmaf(f[i], f[i], var);
          }
#pragma unroll
          for (int off = LPR / 2; off; off >>= 1) var += __shfl_xor_sync(0xffffffffu, var, off);
          const float r = rsqrtf(var * (1.0f / D) + norm_eps);
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = fmaf(f[i] * r, w[i], bias[i]);
          val[u] = pack8<IS_BF16>(f);
        }
      } else if (NORM && t < 2) {
        // q/k RMSNorm of the processor (MW:99-102), statistic from rms_stat_kernel; weights of this head's columns
        float w[8];
        unpack8<IS_BF16>(__ldg(reinterpret_cast<const uint4*>((t == 0 ? wq : wk) + h * D) + chunk), w);
        const float* rs = rstd + static_cast<int64_t>(t) * gridDim.z * S + static_cast<int64_t>(b) * S;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (p0 + u >= passes) continue;
          const float r = __ldg(rs + srcs[u]);
          float f[8];
          unpack8<IS_BF16>(val[u], f);
          if (norm_kind == 2) {  // diffusers: (x * rstd) -> tensor dtype, then * weight
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] *= r;
            unpack8<IS_BF16>(pack8<IS_BF16>(f), f);
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] *= w[i];
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = f[i] * r * w[i];
          }
          val[u] = pack8<IS_BF16>(f);
        }
      }
      if (ROPE && t < 2) {
        // rotary embedding where the reference applies it (modify_wan.py:108-116 / modify_cogvideo.py:59-64):
        // pairs (x[2i], x[2i+1]) times (cos, sin) of the token's position, fp32, rounded back to the tensor dtype
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (p0 + u >= passes || srcs[u] < rope_first) continue;
          const float4* tp = reinterpret_cast<const float4*>(rope + (static_cast<int64_t>(srcs[u] - rope_first) * (D / 2) + chunk * 4) * 2);
          const float4 c01 = __ldg(tp), c23 = __ldg(tp + 1);  // (cos0,sin0,cos1,sin1), (cos2,sin2,cos3,sin3)
          float f[8];
          unpack8<IS_BF16>(val[u], f);
          const float cs[8] = {c01.x, c01.y, c01.z, c01.w, c23.x, c23.y, c23.z, c23.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float a = f[2 * i], bb = f[2 * i + 1];
            f[2 * i] = a * cs[2 * i] - bb * cs[2 * i + 1];
            f[2 * i + 1] = a * cs[2 * i + 1] + bb * cs[2 * i];
          }
          val[u] = pack8<IS_BF16>(f);
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (p0 + u >= passes) break;
        if (COPY && dst != nullptr && rows[u] < S)
          reinterpret_cast<uint4*>(dst + (out_base + rows[u]) * D)[chunk] = val[u];
        if (t < 2) {
          float f[8];
          unpack8<IS_BF16>(val[u], f);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[t][i] += f[i];
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int i = 0; i < 8; ++i) red[t][slot][chunk * 8 + i] = acc[t][i];
  __syncthreads();
  for (int e = threadIdx.x; e < 2 * D; e += 256) {
    const int t = e / D, d = e % D;
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < SLOTS; ++j) s += red[t][j][d];
    float* dstm = t == 0 ? q_mean : k_mean;
    if (dstm && ((tmask >> t) & 1)) dstm[((static_cast<int64_t>(b) * H + h) * nb + blk) * D + d] = s / static_cast<float>(block);
  }
}

// ------------------------------------------------------------------------------------------------
// pool: grid (ceil(np/8), H, B), 256 threads, one warp per pooled row.
// ------------------------------------------------------------------------------------------------
template <int D, bool IS_BF16>
__global__ void __launch_bounds__(256) pool_kernel(const uint16_t* __restrict__ k, const uint16_t* __restrict__ v,
                                                   Strides3 sk, Strides3 sv, const int32_t* __restrict__ src_row,
                                                   uint16_t* __restrict__ k_pool, uint16_t* __restrict__ v_pool, int S,
                                                   int H, int np, int gap) {
  constexpr int LPR = D / 8;
  constexpr int RPW = 32 / LPR;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // reverse traversal: the pooling runs right after the gather kernel, whose LAST writes (high b, h, rows) are the
  // ones still resident in L2
  const int g = (gridDim.x - 1 - blockIdx.x) * 8 + warp;
  if (g >= np) return;
  const int h = gridDim.y - 1 - blockIdx.y, b = gridDim.z - 1 - blockIdx.z;
  const int sub = lane / LPR, chunk = lane % LPR;
  const float inv = 1.0f / static_cast<float>(gap);
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const uint16_t* base = (t == 0 ? k : v) + b * (t == 0 ? sk.b : sv.b) + h * (t == 0 ? sk.h : sv.h);
    const int64_t ss = t == 0 ? sk.s : sv.s;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int j0 = 0; j0 < gap; j0 += RPW * 4) {
      uint4 val[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u * RPW + sub;
        ok[u] = j < gap;
        int r = g * gap + j;
        r = r < S ? r : S - 1;  // replicate padding (W:89 -> W:35)
        const int sr = src_row ? __ldg(src_row + r) : r;
        if (ok[u]) val[u] = ldg_stream(reinterpret_cast<const uint4*>(base + sr * ss) + chunk);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (!ok[u]) continue;
        float f[8];
        unpack8<IS_BF16>(val[u], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += f[i];
      }
    }
#pragma unroll
    for (int off = 16; off >= LPR; off >>= 1)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
    if (sub == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] *= inv;
      uint16_t* dst = (t == 0 ? k_pool : v_pool) + ((static_cast<int64_t>(b) * H + h) * np + g) * D;
      reinterpret_cast<uint4*>(dst)[chunk] = pack8<IS_BF16>(acc);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// scores: grid (ceil(nb/16), B*H), 256 threads.  A CTA owns 16 score rows of one head; the head's K means stream
// through shared memory in chunks of 64 key blocks (prefetched global -> registers while the previous chunk is
// multiplied; rows padded to D+4 floats so that the per-lane float4 reads are bank-conflict free).  Thread
// (qg, kk) = (tid / 64, tid % 64) computes the 4 dot products of query rows 4qg..4qg+3 with key row kk of the
// chunk: one conflict-free LDS.128 of K and four broadcast LDS.128 of Q feed 16 FFMA, which keeps the kernel on
// the FMA pipe instead of the shared-memory port (the one-row-per-warp version spent 31 us there).  The FMA order
// inside a dot product (4 lanes of the float4, then (d0+d1)+(d2+d3)) and the softmax reduction order are fixed.
// dynamic smem: (16*(D+nb) + 64*(D+4)) floats
// ------------------------------------------------------------------------------------------------
constexpr int kScoreRows = 16;
constexpr int kFusedRows = 16;  // rows per CTA of the fused score + selection kernel at full size (8 was slower there: every CTA
                                // streams all K means); few heads (one Ulysses rank) take 8 or 4 rows per CTA to fill the machine
constexpr int kScoreChunk = 64;
// Body shared by score_meanpool_kernel (R = 16 rows per CTA) and the fused score + selection kernel (R = 8: twice the
// CTAs, one selection row per warp): leaves the CTA's R softmax rows in shared memory (srow, [R][nb]) and, if `scores` is given, writes them to global memory.
template <int D, int R>
__device__ __forceinline__ void score_rows(const float* __restrict__ qm, const float* __restrict__ km,
                                              float* __restrict__ scores, int nb, float scale, float* sm) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int DP = D + 4;
  constexpr int d4n = D / 4;
  float* sq = sm;                                  // [16][D]
  float* srow = sm + R * D;               // [16][nb]
  float* sk = sm + R * (D + nb);          // [64][DP]
  const int i0 = blockIdx.x * R;
  const int64_t bh = blockIdx.y;
  for (int e = threadIdx.x; e < R * d4n; e += 256) {
    const int r = e / d4n, c = e % d4n;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i0 + r < nb) v = __ldg(reinterpret_cast<const float4*>(qm + (bh * nb + i0 + r) * D) + c);
    reinterpret_cast<float4*>(sq + r * D)[c] = v;
  }
  const float* kb = km + bh * nb * D;
  constexpr int kPre = kScoreChunk * d4n / 256;  // 64 rows x D/4 float4 over 256 threads
  float4 pre[kPre];
  auto fetch = [&](int j0) {
#pragma unroll
    for (int u = 0; u < kPre; ++u) {
      const int e = threadIdx.x + u * 256;
      const int r = e / d4n, c = e % d4n;
      pre[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e < kScoreChunk * d4n && j0 + r < nb)
        pre[u] = __ldg(reinterpret_cast<const float4*>(kb + static_cast<int64_t>(j0 + r) * D) + c);
    }
  };
  fetch(0);
  const int qg = threadIdx.x >> 6, kk = threadIdx.x & 63;
  for (int j0 = 0; j0 < nb; j0 += kScoreChunk) {
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kPre; ++u) {
      const int e = threadIdx.x + u * 256;
      if (e < kScoreChunk * d4n) *reinterpret_cast<float4*>(sk + (e / d4n) * DP + 4 * (e % d4n)) = pre[u];
    }
    __syncthreads();
    if (j0 + kScoreChunk < nb) fetch(j0 + kScoreChunk);
    if (j0 + kk < nb) {
      const float4* kr = reinterpret_cast<const float4*>(sk + kk * DP);
      constexpr int A = R / 4;                     // query rows per thread group (4 groups of 64 threads)
      const float4* qr = reinterpret_cast<const float4*>(sq + A * qg * D);
      float acc[A][4];
#pragma unroll
      for (int a = 0; a < A; ++a)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[a][i] = 0.f;
#pragma unroll 2
      for (int c = 0; c < d4n; ++c) {
        const float4 kv = kr[c];
#pragma unroll
        for (int a = 0; a < A; ++a) {
          const float4 qv = qr[a * d4n + c];
          acc[a][0] = fmaf(qv.x, kv.x, acc[a][0]);
          acc[a][1] = fmaf(qv.y, kv.y, acc[a][1]);
          acc[a][2] = fmaf(qv.z, kv.z, acc[a][2]);
          acc[a][3] = fmaf(qv.w, kv.w, acc[a][3]);
        }
      }
#pragma unroll
      for (int a = 0; a < A; ++a)
        srow[(A * qg + a) * nb + j0 + kk] = ((acc[a][0] + acc[a][1]) + (acc[a][2] + acc[a][3])) * scale;
    }
  }
  __syncthreads();
  // row softmax: R / 8 rows per warp (R < 8: one row on each of the first R warps)
  constexpr int RPW = R >= 8 ? R / 8 : 1;
  for (int rr = 0; rr < RPW; ++rr) {
    const int r = warp * RPW + rr;
    const int i = i0 + r;
    if (r >= R || i >= nb) break;
    float* row = srow + r * nb;
    float mx = -INFINITY;
    for (int j = lane; j < nb; j += 32) mx = fmaxf(mx, row[j]);
#pragma unroll
    for (int off = 16; off; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    float sum = 0.f;
    for (int j = lane; j < nb; j += 32) {
      const float e = expf(row[j] - mx);
      row[j] = e;
      sum += e;
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    const float inv = 1.0f / sum;
    float* out = scores ? scores + (bh * nb + i) * nb : nullptr;
    for (int j = lane; j < nb; j += 32) {
      const float v = row[j] * inv;
      row[j] = v;
      if (out) out[j] = v;
    }
  }
}

template <int D>
__global__ void __launch_bounds__(256, 3) score_meanpool_kernel(const float* __restrict__ qm, const float* __restrict__ km,
                                                             float* __restrict__ scores, int nb, float scale) {
  extern __shared__ __align__(16) float sm[];
  score_rows<D, kScoreRows>(qm, km, scores, nb, scale, sm);
}

template <int E>
__device__ __forceinline__ int select_row_bitonic(const float* src, int nk, int qi, int nq, int lo, int hi, float thr,
                                                  int force_last, int rounding, double* sorted, int* flag,
                                                  int32_t* irow, uint8_t* mrow, int32_t* kcut_out, int lane);

// Fused mask generation back end (north-star kernel (a), second half): block-score GEMM on the means + row softmax +
// energy-threshold selection + index list, one launch; the fp32 score map only goes to memory when the caller asks for
// it.  Same arithmetic, bit for bit, as score_meanpool_kernel followed by select_bitonic_kernel<E> (the selection reads
// the rows from shared memory instead of global).  dynamic smem: the score part + 8 warps * (N doubles + N ints).
template <int D, int E, int R>
__global__ void __launch_bounds__(256, 3) score_select_kernel(const float* __restrict__ qm, const float* __restrict__ km,
                                                           float* __restrict__ scores_opt, int nb, float scale, int lo,
                                                           int hi, float thr, int force_last, int rounding,
                                                           int32_t* __restrict__ idx, int32_t* __restrict__ cnt,
                                                           uint8_t* __restrict__ mask,
                                                           unsigned long long* __restrict__ sel_acc) {
  constexpr int N = 32 * E;
  extern __shared__ __align__(16) float sm[];
  // programmatic dependent launch: everything above the first global read may overlap the producer's tail
  asm volatile("griddepcontrol.wait;" ::: "memory");
  score_rows<D, R>(qm, km, scores_opt, nb, scale, sm);
  __syncthreads();
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* srow = sm + R * D;
  // the selection scratch (8 warps x (N doubles + N ints) <= 24 KB) aliases the K-means chunk buffer (64 x (D+4) floats
  // >= 17 KB for D = 64, 33 KB for D = 128), dead once the scores are in srow; what does not fit follows the score region
  const size_t srow_end = static_cast<size_t>(R) * (D + nb);
  double* sorted = reinterpret_cast<double*>(sm + ((srow_end + 1) & ~size_t(1))) + warp * 2 * N;
  int* flag = reinterpret_cast<int*>(sorted + N);
  const int64_t bh = blockIdx.y;
  int total = 0;
  constexpr int RPW = R >= 8 ? R / 8 : 1;
  for (int rr = 0; rr < RPW; ++rr) {
    const int r = warp * RPW + rr;
    const int qi = blockIdx.x * R + r;
    if (r >= R || qi >= nb) break;
    const int64_t row = bh * nb + qi;
    const int base = select_row_bitonic<E>(srow + r * nb, nb, qi, nb, lo, hi, thr, force_last, rounding, sorted, flag,
                                           idx + row * nb, mask ? mask + row * nb : nullptr, nullptr, lane);
    if (lane == 0) cnt[row] = base;
    total += base;
  }
  if (sel_acc && lane == 0 && total) atomicAdd(sel_acc, static_cast<unsigned long long>(total));
}

// ------------------------------------------------------------------------------------------------
// select, long rows (nk > 256): one warp per score row, O(n^2) rank counting under (value desc, index asc), fp64
// sequential prefix sums.  dynamic smem: 8 * 2 * nk_pad floats.  Rows of up to 256 blocks take the bitonic kernel.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) select_rank_kernel(const float* __restrict__ scores, int64_t total_rows, int nq, int nk,
                                                          int lo_s, int hi_s, const int32_t* __restrict__ lo_bh,
                                                          const int32_t* __restrict__ hi_bh, float thr, int force_last,
                                                          int32_t* __restrict__ idx, int32_t* __restrict__ cnt,
                                                          uint8_t* __restrict__ mask, int32_t* __restrict__ kcut,
                                                          int rounding, unsigned long long* __restrict__ sel_acc) {
  extern __shared__ __align__(16) float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk_pad = (nk + 31) & ~31;
  float* vals = sm + warp * 2 * nk_pad;
  float* sorted = vals + nk_pad;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  if (row >= total_rows) return;
  const int qi = static_cast<int>(row % nq);
  const int64_t bh = row / nq;
  const float* src = scores + row * nk;
  for (int j = lane; j < nk_pad; j += 32) vals[j] = j < nk ? src[j] : -INFINITY;
  __syncwarp();
  // rank of element j: #{i : v_i > v_j or (v_i == v_j and i < j)}; padding is -inf at indices >= nk
  auto rank_of = [&](int j, float vj) {
    int r = 0;
    for (int i4 = 0; i4 < nk_pad; i4 += 4) {
      const float4 v = *reinterpret_cast<const float4*>(vals + i4);  // warp-uniform address: broadcast
      r += (v.x > vj) || (v.x == vj && i4 + 0 < j);
      r += (v.y > vj) || (v.y == vj && i4 + 1 < j);
      r += (v.z > vj) || (v.z == vj && i4 + 2 < j);
      r += (v.w > vj) || (v.w == vj && i4 + 3 < j);
    }
    return r;
  };
  for (int j = lane; j < nk; j += 32) {
    const float vj = vals[j];
    sorted[rank_of(j, vj)] = vj;
  }
  __syncwarp();
  const int lo = lo_bh ? lo_bh[bh] : lo_s;
  const int hi = hi_bh ? hi_bh[bh] : hi_s;
  int kfirst = nk;
  if (lane == 0) {
    const int lim = hi < nk ? hi : nk;  // beyond `hi` the clamp decides
    if (rounding == 0) {
      double acc = 0.0;
      for (int i = 0; i < nk; ++i) acc += static_cast<double>(sorted[i]);
      const float th = __fmul_rn(thr, static_cast<float>(acc));
      acc = 0.0;
      for (int i = 0; i < lim; ++i) {
        acc += static_cast<double>(sorted[i]);
        if (static_cast<float>(acc) >= th) {
          kfirst = i;
          break;
        }
      }
    } else {
      float acc = 0.f;
      for (int i = 0; i < nk; ++i) acc = __fadd_rn(acc, sorted[i]);
      const float th = round_sel(__fmul_rn(thr, round_sel(acc, rounding)), rounding);
      acc = 0.f;
      for (int i = 0; i < lim; ++i) {
        acc = __fadd_rn(acc, sorted[i]);
        if (round_sel(acc, rounding) >= th) {
          kfirst = i;
          break;
        }
      }
    }
  }
  kfirst = __shfl_sync(0xffffffffu, kfirst, 0);
  int kc = kfirst < lo ? lo : kfirst;
  kc = kc > hi ? hi : kc;
  if (kcut && lane == 0) kcut[row] = kc;
  const bool full_row = force_last > 0 && qi >= nq - force_last;
  int32_t* irow = idx + row * nk;
  uint8_t* mrow = mask ? mask + row * nk : nullptr;
  int base = 0;
  for (int s = 0; s < nk_pad / 32; ++s) {
    const int j = s * 32 + lane;
    const int rk = rank_of(j, vals[j]);
    const bool sel = j < nk && (rk < kc || full_row || (force_last > 0 && j >= nk - force_last));
    const unsigned bal = __ballot_sync(0xffffffffu, sel);
    if (sel) irow[base + __popc(bal & ((1u << lane) - 1u))] = j;
    if (mrow && j < nk) mrow[j] = sel ? 1 : 0;
    base += __popc(bal);
  }
  for (int j = base + lane; j < nk; j += 32) irow[j] = -1;
  if (lane == 0) {
    cnt[row] = base;
    if (sel_acc) atomicAdd(sel_acc, static_cast<unsigned long long>(base));
  }
}

// One score row through the selection: `src` = the row's nk fp32 scores (global or shared memory), sorted / flag = this
// warp's scratch (N doubles + N ints).  Returns the number of selected blocks (all lanes).
template <int E>
__device__ __forceinline__ int select_row_bitonic(const float* src, int nk, int qi, int nq, int lo, int hi, float thr,
                                                  int force_last, int rounding, double* sorted, int* flag,
                                                  int32_t* irow, uint8_t* mrow, int32_t* kcut_out, int lane) {
  constexpr int N = 32 * E;
  // One 64-bit key per element: high word = the value's bits mapped to an order-preserving unsigned, low word =
  // ~index.  "a before b" (value desc, index asc) is then a single unsigned compare ka > kb: 2 ISETP + 2 SEL per
  // compare-exchange instead of three compares, predicate logic and two selects (the ALU pipe bounds this kernel).
  // x + 0.0f folds -0.0 into +0.0, which torch.sort treats as equal.
  uint64_t key[E];
#pragma unroll
  for (int r = 0; r < E; ++r) {
    const int p = lane * E + r;
    const float x = p < nk ? src[p] + 0.0f : -INFINITY;  // padding sorts behind every real entry
    const uint32_t b = __float_as_uint(x);
    const uint32_t mono = b ^ (static_cast<uint32_t>(static_cast<int32_t>(b) >> 31) | 0x80000000u);
    key[r] = (static_cast<uint64_t>(mono) << 32) | static_cast<uint32_t>(~p);
  }
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j >= 1; j >>= 1) {
      if (j >= E) {                                          // partner lives in lane ^ (j / E), same r
        // p = lane*E + r with k, j multiples of E: direction and side depend on the lane only
        const bool keep_if_first = ((lane * E) & k) == 0 == (((lane * E) & j) == 0);
#pragma unroll
        for (int r = 0; r < E; ++r) {
          const uint64_t ok = __shfl_xor_sync(0xffffffffu, key[r], j / E);
          const bool keep = (key[r] > ok) == keep_if_first;
          key[r] = keep ? key[r] : ok;
        }
      } else {                                               // partner is register r ^ j of the same lane
#pragma unroll
        for (int r = 0; r < E; ++r) {
          if ((r & j) == 0) {
            const int r2 = r | j;
            const bool up = ((lane * E + r) & k) == 0;
            const bool swap = (key[r] > key[r2]) != up;      // lower slot must hold the "first" element iff up
            const uint64_t t = key[r];
            key[r] = swap ? key[r2] : t;
            key[r2] = swap ? t : key[r2];
          }
        }
      }
    }
  }
  float v[E];
  int id[E];
#pragma unroll
  for (int r = 0; r < E; ++r) {
    const uint32_t mono = static_cast<uint32_t>(key[r] >> 32);
    v[r] = __uint_as_float(mono ^ ((mono >> 31) ? 0x80000000u : 0xFFFFFFFFu));
    id[r] = static_cast<int>(~static_cast<uint32_t>(key[r]));
  }
#pragma unroll
  for (int r = 0; r < E; ++r) {
    sorted[lane * E + r] = static_cast<double>(v[r]);
    flag[lane * E + r] = 0;
  }
  __syncwarp();
  int kfirst = nk;
  if (rounding == 0) {   // warp-uniform
    kfirst = energy_cut_warp(sorted, reinterpret_cast<float*>(flag), nk, hi < nk ? hi : nk, thr, lane);
    __syncwarp();
#pragma unroll
    for (int r = 0; r < E; ++r) flag[lane * E + r] = 0;      // the prefix scratch was the flag array
    __syncwarp();
  } else {
    if (lane == 0) kfirst = energy_cut(sorted, nk, hi < nk ? hi : nk /* beyond `hi` the clamp decides */, thr, rounding);
    kfirst = __shfl_sync(0xffffffffu, kfirst, 0);
  }
  int kc = kfirst < lo ? lo : kfirst;
  kc = kc > hi ? hi : kc;
  if (kcut_out && lane == 0) *kcut_out = kc;
#pragma unroll
  for (int r = 0; r < E; ++r)
    if (lane * E + r < kc && id[r] < nk) flag[id[r]] = 1;    // sorted position < cut  ->  keep that block
  __syncwarp();

  const bool full_row = force_last > 0 && qi >= nq - force_last;
  int base = 0;
#pragma unroll
  for (int sblk = 0; sblk < E; ++sblk) {
    const int j = sblk * 32 + lane;
    const bool sel = j < nk && (flag[j] != 0 || full_row || (force_last > 0 && j >= nk - force_last));
    const unsigned bal = __ballot_sync(0xffffffffu, sel);
    if (sel) irow[base + __popc(bal & ((1u << lane) - 1u))] = j;
    if (mrow && j < nk) mrow[j] = sel ? 1 : 0;
    base += __popc(bal);
  }
  for (int j = base + lane; j < nk; j += 32) irow[j] = -1;
  __syncwarp();
  return base;
}

// ------------------------------------------------------------------------------------------------
// select (nk <= 256): one warp per score row, E = elements per lane (nk <= 32*E).  The row is sorted by a
// bitonic network held in registers (position p = lane*E + r; partners with j < E are in-lane, the others one
// __shfl_xor away) under the total order (value desc, index asc) -- the same order as torch.sort(stable=True),
// unique keys so the network's instability does not matter.  Then: fp64 sequential prefix sums by lane 0
// (bit-exact with torch's CPU cumsum), clamp, flag scatter, ballot compaction to the ascending index list.
// dynamic smem: 8 warps * 2 * 32*E doubles (sorted values as fp64 + int flags).
// ------------------------------------------------------------------------------------------------
template <int E>
__global__ void __launch_bounds__(256) select_bitonic_kernel(const float* __restrict__ scores, int64_t total_rows, int nq,
                                                             int nk, int lo_s, int hi_s, const int32_t* __restrict__ lo_bh,
                                                             const int32_t* __restrict__ hi_bh, float thr, int force_last,
                                                             int32_t* __restrict__ idx, int32_t* __restrict__ cnt,
                                                             uint8_t* __restrict__ mask, int32_t* __restrict__ kcut,
                                                             int rounding, unsigned long long* __restrict__ sel_acc) {
  constexpr int N = 32 * E;
  extern __shared__ __align__(16) float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* sorted = reinterpret_cast<double*>(sm) + warp * 2 * N;   // sorted values, widened in parallel
  int* flag = reinterpret_cast<int*>(sorted + N);                 // flag[original index] = selected
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  if (row >= total_rows) return;
  const int qi = static_cast<int>(row % nq);
  const int64_t bh = row / nq;
  const int lo = lo_bh ? lo_bh[bh] : lo_s;
  const int hi = hi_bh ? hi_bh[bh] : hi_s;
  const int base = select_row_bitonic<E>(scores + row * nk, nk, qi, nq, lo, hi, thr, force_last, rounding, sorted, flag,
                                         idx + row * nk, mask ? mask + row * nk : nullptr, kcut ? kcut + row : nullptr,
                                         lane);
  if (lane == 0) {
    cnt[row] = base;
    if (sel_acc) atomicAdd(sel_acc, static_cast<unsigned long long>(base));
  }
}

__global__ void __launch_bounds__(256) mask_to_index_kernel(const uint8_t* __restrict__ mask, int64_t total_rows, int nk,
                                                            int32_t* __restrict__ idx, int32_t* __restrict__ cnt) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  if (row >= total_rows) return;
  const uint8_t* mrow = mask + row * nk;
  int32_t* irow = idx + row * nk;
  int base = 0;
  for (int j0 = 0; j0 < nk; j0 += 32) {
    const int j = j0 + lane;
    const bool sel = j < nk && mrow[j] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, sel);
    if (sel) irow[base + __popc(bal & ((1u << lane) - 1u))] = j;
    base += __popc(bal);
  }
  for (int j = base + lane; j < nk; j += 32) irow[j] = -1;
  if (lane == 0) cnt[row] = base;
}

// 64-granular bool mask [rows64][nk64] -> per 128-row query tile: ascending list of 128-key tiles that contain at
// least one selected 64x64 block, each entry = tile id | (quadrant mask << 28), quadrant bit = 2*rowhalf + colhalf.
// One warp per query tile.
__global__ void __launch_bounds__(256) mask64_to_index_kernel(const uint8_t* __restrict__ mask, int64_t total_tiles, int nq64,
                                                              int nk64, int nq128, int nk128, int32_t* __restrict__ idx,
                                                              int32_t* __restrict__ cnt) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t tile = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  if (tile >= total_tiles) return;
  const int qt = static_cast<int>(tile % nq128);
  const int64_t bh = tile / nq128;
  const uint8_t* m0 = mask + (bh * nq64 + 2 * qt) * nk64;
  const bool has1 = 2 * qt + 1 < nq64;
  const uint8_t* m1 = m0 + nk64;
  int32_t* irow = idx + tile * nk128;
  int base = 0;
  for (int j0 = 0; j0 < nk128; j0 += 32) {
    const int kt = j0 + lane;
    unsigned fl = 0;
    if (kt < nk128) {
      const int c0 = 2 * kt, c1 = 2 * kt + 1;
      fl |= m0[c0] ? 1u : 0u;
      if (c1 < nk64) fl |= m0[c1] ? 2u : 0u;
      if (has1) {
        fl |= m1[c0] ? 4u : 0u;
        if (c1 < nk64) fl |= m1[c1] ? 8u : 0u;
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, fl != 0);
    if (fl) irow[base + __popc(bal & ((1u << lane) - 1u))] = static_cast<int32_t>(kt | (fl << 28));
    base += __popc(bal);
  }
  for (int j = base + lane; j < nk128; j += 32) irow[j] = -1;
  if (lane == 0) cnt[tile] = base;
}

}  // namespace blade

// ================================================================================================
// C ABI
// ================================================================================================
using namespace blade;

extern "C" int blade_asa_prep(const BladeTensor* q, const BladeTensor* k, const BladeTensor* v, const int32_t* src_row,
                              void* q_r, void* k_r, void* v_r, float* q_mean, float* k_mean, void* k_pool,
                              void* v_pool, int32_t block_size, int32_t sample_gap, void* stream_) {
  return blade_asa_prep_rope(q, k, v, src_row, q_r, k_r, v_r, q_mean, k_mean, k_pool, v_pool, block_size, sample_gap,
                             nullptr, 0, stream_);
}

extern "C" int blade_asa_prep_rope(const BladeTensor* q, const BladeTensor* k, const BladeTensor* v,
                                   const int32_t* src_row, void* q_r, void* k_r, void* v_r, float* q_mean,
                                   float* k_mean, void* k_pool, void* v_pool, int32_t block_size, int32_t sample_gap,
                                   const float* rope_cos_sin, int32_t rope_first_row, void* stream_) {
  return blade::prep_impl(q, k, v, src_row, q_r, k_r, v_r, q_mean, k_mean, k_pool, v_pool, block_size, sample_gap,
                          rope_cos_sin, rope_first_row, 3, static_cast<cudaStream_t>(stream_));
}

int blade::rms_stat_impl(const BladeTensor* q, const BladeTensor* k, float eps, float* out, cudaStream_t stream,
                         float* const* peer_out, int n_peers, int64_t out_rows, int64_t out_row0) {
  if (int e = check_tensor16(q, "q")) return e;
  if (int e = check_tensor16(k, "k")) return e;
  BLADE_REQUIRE(out || (peer_out && n_peers > 0), BLADE_ERR_ARG, "rstd output null");
  BLADE_REQUIRE(n_peers >= 0 && n_peers <= BLADE_MAX_PEERS, BLADE_ERR_ARG, "n_peers %d out of range", n_peers);
  StatPeers sp{};
  if (peer_out && n_peers > 0) {
    sp.n = n_peers;
    sp.out_rows = out_rows;
    sp.out_row0 = out_row0;
    for (int i = 0; i < n_peers; ++i) {
      BLADE_REQUIRE(peer_out[i], BLADE_ERR_ARG, "peer rstd table %d null", i);
      sp.out[i] = peer_out[i];
    }
    BLADE_REQUIRE(q->shape[0] == 1 && out_row0 + q->shape[2] <= out_rows, BLADE_ERR_SHAPE, "peer rstd rows out of range");
  }
  const int64_t B = q->shape[0], H = q->shape[1], S = q->shape[2], D = q->shape[3];
  for (const BladeTensor* t : {q, k}) {
    BLADE_REQUIRE(t->shape[0] == B && t->shape[1] == H && t->shape[2] == S && t->shape[3] == D && t->dtype == q->dtype,
                  BLADE_ERR_SHAPE, "q/k differ");
    BLADE_REQUIRE(t->stride[1] == D && t->stride[2] == H * D, BLADE_ERR_SHAPE,
                  "the RMSNorm statistic needs token-major q/k ([B,S,H*D] memory)");
  }
  BLADE_REQUIRE((H * D) % 8 == 0, BLADE_ERR_SHAPE, "H*D must be a multiple of 8");
  const int64_t rows = B * S;
  const unsigned grid = static_cast<unsigned>(ceil_div(rows, 8));
  const uint16_t *qp = static_cast<const uint16_t*>(q->ptr), *kp = static_cast<const uint16_t*>(k->ptr);
  if (q->dtype == BLADE_BF16)
    rms_stat_kernel<true><<<grid, 256, 0, stream>>>(qp, kp, q->stride[0], q->stride[2], k->stride[0], k->stride[2],
                                                    static_cast<int>(S), rows, static_cast<int>(H * D), eps, out, sp);
  else
    rms_stat_kernel<false><<<grid, 256, 0, stream>>>(qp, kp, q->stride[0], q->stride[2], k->stride[0], k->stride[2],
                                                     static_cast<int>(S), rows, static_cast<int>(H * D), eps, out, sp);
  BLADE_CUDA_OK(cudaGetLastError());
  return BLADE_OK;
}

extern "C" int blade_qk_rms_stat(const BladeTensor* q, const BladeTensor* k, float eps, float* rstd_out, void* stream) {
  return blade::rms_stat_impl(q, k, eps, rstd_out, static_cast<cudaStream_t>(stream));
}

extern "C" int blade_qk_rms_stat_peers(const BladeTensor* q, const BladeTensor* k, float eps, float* const* rstd_peers,
                                       int32_t n_peers, int64_t total_rows, int64_t first_row, void* stream) {
  return blade::rms_stat_impl(q, k, eps, nullptr, static_cast<cudaStream_t>(stream), rstd_peers, n_peers, total_rows,
                              first_row);
}

// parts: bit 0 = gather/copy/means kernel, bit 1 = gap-pooling kernel (reads the copies when they exist)
int blade::prep_impl(const BladeTensor* q, const BladeTensor* k, const BladeTensor* v, const int32_t* src_row, void* q_r,
                     void* k_r, void* v_r, float* q_mean, float* k_mean, void* k_pool, void* v_pool,
                     int32_t block_size, int32_t sample_gap, const float* rope_cos_sin, int32_t rope_first_row,
                     int parts, cudaStream_t stream, const PrepNorm* norm, const int32_t* tok_row,
                     const BladePeers* peers, int tmask, int stage) {
  if (int e = check_tensor16(q, "q")) return e;
  if (int e = check_tensor16(k, "k")) return e;
  if (int e = check_tensor16(v, "v")) return e;
  const int64_t B = q->shape[0], H = q->shape[1], S = q->shape[2], D = q->shape[3];
  for (int i = 0; i < 4; ++i)
    BLADE_REQUIRE(k->shape[i] == q->shape[i] && v->shape[i] == q->shape[i], BLADE_ERR_SHAPE,
                  "q/k/v shapes differ in dim %d (reference asserts equality, W:250-251)", i);
  BLADE_REQUIRE(q->dtype == k->dtype && q->dtype == v->dtype, BLADE_ERR_DTYPE, "q/k/v dtypes differ");
  BLADE_REQUIRE(block_size == 64 || block_size == 128, BLADE_ERR_ARG, "block_size %d not in {64,128}", block_size);
  BLADE_REQUIRE(S >= 1 && B >= 1 && H >= 1 && H <= 65535 && B <= 65535, BLADE_ERR_SHAPE, "bad B/H/S");
  const bool copy = q_r || k_r || v_r;
  // v is only rewritten when rows move (gather / peer pull); rope and norm touch q and k alone, so v_r may stay NULL then
  BLADE_REQUIRE(!copy || (q_r && k_r && (v_r || (!src_row && !(peers && peers->q[0])))), BLADE_ERR_ARG,
                "q_r/k_r must both be set, and v_r too when src_row / peers move rows");
  PeerSrc psrc{};
  if (peers && peers->q[0]) {
    BLADE_REQUIRE(peers->n_peers >= 1 && peers->n_peers <= BLADE_MAX_PEERS && peers->my_peer >= 0 &&
                      peers->my_peer < peers->n_peers && peers->rows_per_peer >= 1,
                  BLADE_ERR_ARG, "BladePeers: bad n_peers / my_peer / rows_per_peer");
    BLADE_REQUIRE(B == 1 && copy, BLADE_ERR_ARG, "peer pull needs B == 1 and the gathered copies");
    BLADE_REQUIRE(static_cast<int64_t>(peers->n_peers) * peers->rows_per_peer >= S, BLADE_ERR_SHAPE,
                  "peers hold %lld rows, sequence has %lld", (long long)peers->n_peers * peers->rows_per_peer, (long long)S);
    const int64_t h_total = static_cast<int64_t>(peers->n_peers) * H;
    for (const BladeTensor* t : {q, k, v})
      BLADE_REQUIRE(t->stride[1] == D && t->stride[2] == h_total * D, BLADE_ERR_SHAPE,
                    "peer pull: q/k/v must describe [rows, n_peers*H, D] token-major memory");
    for (int i = 0; i < peers->n_peers; ++i) {
      BLADE_REQUIRE(peers->q[i] && peers->k[i] && peers->v[i], BLADE_ERR_ARG, "peer %d: q/k/v pointer null", i);
      const int64_t off = static_cast<int64_t>(peers->my_peer) * H * D;
      psrc.base[0][i] = static_cast<const uint16_t*>(peers->q[i]) + off;
      psrc.base[1][i] = static_cast<const uint16_t*>(peers->k[i]) + off;
      psrc.base[2][i] = static_cast<const uint16_t*>(peers->v[i]) + off;
    }
    psrc.rows = peers->rows_per_peer;
  }
  BLADE_REQUIRE(copy || !src_row, BLADE_ERR_ARG, "src_row given but no output copies requested");
  BLADE_REQUIRE(copy || !rope_cos_sin, BLADE_ERR_ARG, "rotary embedding needs the q_r/k_r/v_r outputs");
  BLADE_REQUIRE(!rope_cos_sin || (reinterpret_cast<uintptr_t>(rope_cos_sin) & 15) == 0, BLADE_ERR_ALIGN,
                "rope table not 16B aligned");
  const int nb = static_cast<int>(ceil_div(S, block_size));
  const bool bf = q->dtype == BLADE_BF16;
  Strides3 sq{q->stride[0], q->stride[1], q->stride[2]}, sk{k->stride[0], k->stride[1], k->stride[2]},
      sv{v->stride[0], v->stride[1], v->stride[2]};
  StageTimer timer(stage != -2 ? stage : (parts == 2 ? 4 : 0), stream);
  const uint16_t *qp = static_cast<const uint16_t*>(q->ptr), *kp = static_cast<const uint16_t*>(k->ptr),
                 *vp = static_cast<const uint16_t*>(v->ptr);
  const float* rstd = nullptr;
  const uint16_t *wq = nullptr, *wk = nullptr, *bq = nullptr, *bk = nullptr;
  int norm_kind = 0;
  float norm_eps = 0.f;
  if (norm && norm->kind != 0 && (parts & 1) && (tmask & 3)) {
    BLADE_REQUIRE(norm->kind >= 1 && norm->kind <= 3, BLADE_ERR_ARG, "qk_norm kind %d not in {1,2,3}", norm->kind);
    BLADE_REQUIRE(copy, BLADE_ERR_ARG, "qk_norm needs the q_r/k_r/v_r outputs");
    BLADE_REQUIRE(norm->q_weight && norm->k_weight && (norm->kind == 3 || norm->rstd || norm->rstd_ext), BLADE_ERR_ARG,
                  "qk_norm weights / scratch missing");
    BLADE_REQUIRE((reinterpret_cast<uintptr_t>(norm->q_weight) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(norm->k_weight) & 15) == 0,
                  BLADE_ERR_ALIGN, "qk_norm weights not 16B aligned");
    if (norm->kind != 3) {
      if (!norm->rstd_ext)
        if (int e = rms_stat_impl(q, k, norm->eps, norm->rstd, stream)) return e;
      rstd = norm->rstd_ext ? norm->rstd_ext : norm->rstd;
    }
    wq = static_cast<const uint16_t*>(norm->q_weight);
    wk = static_cast<const uint16_t*>(norm->k_weight);
    bq = static_cast<const uint16_t*>(norm->q_bias);
    bk = static_cast<const uint16_t*>(norm->k_bias);
    norm_eps = norm->eps;
    norm_kind = norm->kind;
  }
  if ((parts & 1) && (q_mean || k_mean || copy)) {
    dim3 grid(nb, static_cast<unsigned>(H), static_cast<unsigned>(B));
#define LAUNCH_PREP_P(DD, BF, CP, RP, NM, PE)                                                                         \
  prep_block_kernel<DD, BF, CP, RP, NM, PE><<<grid, 256, 0, stream>>>(                                                \
      qp, kp, vp, sq, sk, sv, src_row, static_cast<uint16_t*>(q_r), static_cast<uint16_t*>(k_r),                      \
      static_cast<uint16_t*>(v_r), q_mean, k_mean, static_cast<int>(S), static_cast<int>(H), nb, block_size,          \
      rope_cos_sin, rope_first_row, rstd, wq, wk, norm_kind, tok_row, bq, bk, norm_eps, psrc, tmask)
#define LAUNCH_PREP(DD, BF, CP, RP, NM)                                                                               \
  do {                                                                                                                \
    if (CP && psrc.rows) LAUNCH_PREP_P(DD, BF, CP, RP, NM, (CP)); else LAUNCH_PREP_P(DD, BF, CP, RP, NM, false);      \
  } while (0)
#define LAUNCH_PREP_B(DD, CP, RP, NM)                                                                                 \
  do {                                                                                                                \
    if (bf) LAUNCH_PREP(DD, true, CP, RP, NM); else LAUNCH_PREP(DD, false, CP, RP, NM);                               \
  } while (0)
#define LAUNCH_PREP_D(DD)                                                                                             \
  do {                                                                                                                \
    if (norm_kind && rope_cos_sin) LAUNCH_PREP_B(DD, true, true, true);                                               \
    else if (norm_kind) LAUNCH_PREP_B(DD, true, false, true);                                                         \
    else if (rope_cos_sin) LAUNCH_PREP_B(DD, true, true, false);                                                      \
    else if (copy) LAUNCH_PREP_B(DD, true, false, false);                                                             \
    else LAUNCH_PREP_B(DD, false, false, false);                                                                      \
  } while (0)
    if (D == 128) LAUNCH_PREP_D(128); else LAUNCH_PREP_D(64);
#undef LAUNCH_PREP_D
#undef LAUNCH_PREP_B
#undef LAUNCH_PREP
#undef LAUNCH_PREP_P
    BLADE_CUDA_OK(cudaGetLastError());
  }
  if ((parts & 2) && sample_gap > 0 && k_pool && v_pool) {
    const int np = static_cast<int>(ceil_div(S, sample_gap));
    // read the curve-ordered copies when they exist (contiguous, no gather), else the sources
    const uint16_t* ks = copy ? static_cast<const uint16_t*>(k_r) : kp;
    const uint16_t* vs = v_r ? static_cast<const uint16_t*>(v_r) : vp;
    Strides3 ck{H * S * D, S * D, D};
    Strides3 pk = copy ? ck : sk, pv = v_r ? ck : sv;
    const int32_t* sr = copy ? nullptr : src_row;   // (no v copy implies no src_row: both tensors are read in place)
    dim3 grid(static_cast<unsigned>(ceil_div(np, 8)), static_cast<unsigned>(H), static_cast<unsigned>(B));
#define LAUNCH_POOL(DD, BF)                                                                                      \
  pool_kernel<DD, BF><<<grid, 256, 0, stream>>>(ks, vs, pk, pv, sr, static_cast<uint16_t*>(k_pool),             \
                                                static_cast<uint16_t*>(v_pool), static_cast<int>(S),            \
                                                static_cast<int>(H), np, sample_gap)
    if (D == 128) { if (bf) LAUNCH_POOL(128, true); else LAUNCH_POOL(128, false); }
    else          { if (bf) LAUNCH_POOL(64, true); else LAUNCH_POOL(64, false); }
#undef LAUNCH_POOL
    BLADE_CUDA_OK(cudaGetLastError());
  }
  return BLADE_OK;
}

extern "C" int blade_asa_scores_meanpool(const float* q_mean, const float* k_mean, float* scores, int64_t B, int64_t H,
                                         int64_t nb, int64_t D, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BLADE_REQUIRE(q_mean && k_mean && scores, BLADE_ERR_ARG, "null pointer");
  BLADE_REQUIRE((D == 64 || D == 128) && nb >= 1 && nb <= 4096, BLADE_ERR_SHAPE, "bad nb/D");
  const size_t smem = (kScoreRows * (D + nb) + kScoreChunk * (D + 4)) * sizeof(float);
  BLADE_REQUIRE(smem <= 200 * 1024, BLADE_ERR_SHAPE, "nb too large for score kernel");
  StageTimer timer(1, stream);
  dim3 grid(static_cast<unsigned>(ceil_div(nb, kScoreRows)), static_cast<unsigned>(B * H));
  const float scale = 1.0f / sqrtf(static_cast<float>(D));
  auto kern = D == 128 ? score_meanpool_kernel<128> : score_meanpool_kernel<64>;
  if (smem > 48 * 1024)
    BLADE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, 256, smem, stream>>>(q_mean, k_mean, scores, (int)nb, scale);
  BLADE_CUDA_OK(cudaGetLastError());
  return BLADE_OK;
}

// fused scores + selection (nb <= 256); returns BLADE_OK, or -1 when the shape needs the two-kernel path
int blade::score_select_impl(const float* q_mean, const float* k_mean, float* scores_opt, int64_t B, int64_t H, int64_t nb,
                             int64_t D, const BladeAsaConfig* cfg, int32_t* idx, int32_t* cnt, uint8_t* mask_opt,
                             cudaStream_t stream, bool pdl) {
  if (nb > 256 || (D != 64 && D != 128)) return -1;
  const int E = nb <= 32 ? 1 : (nb <= 64 ? 2 : (nb <= 128 ? 4 : 8));
  // rows per CTA: 16, or fewer when that leaves SMs without a CTA (the kernel's time is the latency of ONE CTA: with 3 Wan
  // heads 16 rows per CTA is 48 CTAs of ~50 us each).  BLADE_FUSED_ROWS = 4 | 8 | 16 forces one (A/B).
  static const int env_rows = getenv("BLADE_FUSED_ROWS") ? atoi(getenv("BLADE_FUSED_ROWS")) : 0;
  int R = kFusedRows;
  while (R > 8 && B * H * ceil_div(nb, R) < device_sm_count()) R >>= 1;   // measured: 3 Wan heads 44 / 30 / 32 us at 16 / 8 / 4
  if (env_rows == 4 || env_rows == 8 || env_rows == 16) R = env_rows;
  const size_t srow_end = ((static_cast<size_t>(R) * (D + nb) + 1) & ~size_t(1)) * sizeof(float);
  const size_t chunk_bytes = static_cast<size_t>(kScoreChunk) * (D + 4) * sizeof(float);
  const size_t sel_bytes = 8 * (2 * 32 * E) * sizeof(double);
  const size_t smem = srow_end + (chunk_bytes > sel_bytes ? chunk_bytes : sel_bytes) + 16;
  if (smem > 200 * 1024) return -1;
  BLADE_REQUIRE(cfg->min_retain >= 1 && cfg->max_retain >= 1, BLADE_ERR_ARG, "retain bounds must be >= 1");
  BLADE_REQUIRE(cfg->select_rounding >= 0 && cfg->select_rounding <= 2, BLADE_ERR_ARG, "select_rounding %d not in {0,1,2}",
                cfg->select_rounding);
  StageTimer t1(1, stream);   // the fused launch is reported as stage 1 (scores); stage 2 (select) collapses to ~0
  const float scale = 1.0f / sqrtf(static_cast<float>(D));
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3(static_cast<unsigned>(ceil_div(nb, R)), static_cast<unsigned>(B * H));
  lc.blockDim = dim3(256);
  lc.dynamicSmemBytes = smem;
  lc.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = at;
  lc.numAttrs = pdl ? 1 : 0;
  const int nbi = static_cast<int>(nb);
#define LAUNCH_FUSED(DD, EE)                                                                                          \
  do {                                                                                                                \
    auto kern = R == 16 ? score_select_kernel<DD, EE, 16> : (R == 8 ? score_select_kernel<DD, EE, 8> : score_select_kernel<DD, EE, 4>); \
    if (smem > 48 * 1024)                                                                                             \
      BLADE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));              \
    BLADE_CUDA_OK(cudaLaunchKernelEx(&lc, kern, q_mean, k_mean, scores_opt, nbi, scale, (int)cfg->min_retain,         \
                                     (int)cfg->max_retain, cfg->energy_threshold, (int)cfg->force_last,               \
                                     (int)cfg->select_rounding, idx, cnt, mask_opt, cfg->selected_acc));              \
  } while (0)
#define LAUNCH_FUSED_D(DD)                                                                                            \
  do {                                                                                                                \
    if (E == 1) LAUNCH_FUSED(DD, 1); else if (E == 2) LAUNCH_FUSED(DD, 2); else if (E == 4) LAUNCH_FUSED(DD, 4);      \
    else LAUNCH_FUSED(DD, 8);                                                                                         \
  } while (0)
  if (D == 128) LAUNCH_FUSED_D(128); else LAUNCH_FUSED_D(64);
#undef LAUNCH_FUSED_D
#undef LAUNCH_FUSED
  { StageTimer t2(2, stream); }
  return BLADE_OK;
}

extern "C" int blade_asa_select(const float* scores, int64_t B, int64_t H, int64_t nq, int64_t nk,
                                const BladeAsaConfig* cfg, const int32_t* lo_bh, const int32_t* hi_bh, int32_t* idx,
                                int32_t* cnt, uint8_t* mask_opt, int32_t* kcut_opt, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BLADE_REQUIRE(scores && idx && cnt && cfg, BLADE_ERR_ARG, "null pointer");
  BLADE_REQUIRE(nq >= 1 && nk >= 1 && nk <= 4096, BLADE_ERR_SHAPE, "bad nq/nk");
  BLADE_REQUIRE(cfg->min_retain >= 1 && cfg->max_retain >= 1, BLADE_ERR_ARG, "retain bounds must be >= 1");
  BLADE_REQUIRE(cfg->select_rounding >= 0 && cfg->select_rounding <= 2, BLADE_ERR_ARG, "select_rounding %d not in {0,1,2}",
                cfg->select_rounding);
  const int64_t rows = B * H * nq;
  const int nk_pad = static_cast<int>((nk + 31) & ~31);
  const size_t smem = 8 * 2 * nk_pad * sizeof(float);
  StageTimer timer(2, stream);
#define LAUNCH_BITONIC(EE)                                                                                       \
  do {                                                                                                             \
    const size_t sm_b = 8 * (2 * 32 * EE) * sizeof(double);                                                           \
    select_bitonic_kernel<EE><<<static_cast<unsigned>(ceil_div(rows, 8)), 256, sm_b, stream>>>(                    \
        scores, rows, (int)nq, (int)nk, cfg->min_retain, cfg->max_retain, lo_bh, hi_bh, cfg->energy_threshold,     \
        cfg->force_last, idx, cnt, mask_opt, kcut_opt, cfg->select_rounding, cfg->selected_acc);                   \
  } while (0)
  if (nk <= 32) LAUNCH_BITONIC(1);
  else if (nk <= 64) LAUNCH_BITONIC(2);
  else if (nk <= 128) LAUNCH_BITONIC(4);
  else if (nk <= 256) LAUNCH_BITONIC(8);
  else {  // long rows: O(n^2) rank counting fallback
    if (smem > 48 * 1024)
      BLADE_CUDA_OK(cudaFuncSetAttribute(select_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    select_rank_kernel<<<static_cast<unsigned>(ceil_div(rows, 8)), 256, smem, stream>>>(
        scores, rows, (int)nq, (int)nk, cfg->min_retain, cfg->max_retain, lo_bh, hi_bh, cfg->energy_threshold,
        cfg->force_last, idx, cnt, mask_opt, kcut_opt, cfg->select_rounding, cfg->selected_acc);
  }
#undef LAUNCH_BITONIC
  BLADE_CUDA_OK(cudaGetLastError());
  return BLADE_OK;
}

extern "C" int blade_mask_to_index(const uint8_t* mask, int64_t B, int64_t H, int64_t nq, int64_t nk, int32_t* idx,
                                   int32_t* cnt, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BLADE_REQUIRE(mask && idx && cnt, BLADE_ERR_ARG, "null pointer");
  const int64_t rows = B * H * nq;
  mask_to_index_kernel<<<static_cast<unsigned>(ceil_div(rows, 8)), 256, 0, stream>>>(mask, rows, (int)nk, idx, cnt);
  BLADE_CUDA_OK(cudaGetLastError());
  return BLADE_OK;
}

extern "C" int blade_mask64_to_index(const uint8_t* mask, int64_t B, int64_t H, int64_t nq64, int64_t nk64, int32_t* idx,
                                     int32_t* cnt, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BLADE_REQUIRE(mask && idx && cnt, BLADE_ERR_ARG, "null pointer");
  const int64_t nq128 = ceil_div(nq64, 2), nk128 = ceil_div(nk64, 2);
  const int64_t tiles = B * H * nq128;
  mask64_to_index_kernel<<<static_cast<unsigned>(ceil_div(tiles, 8)), 256, 0, stream>>>(
      mask, tiles, (int)nq64, (int)nk64, (int)nq128, (int)nk128, idx, cnt);
  BLADE_CUDA_OK(cudaGetLastError());
  return BLADE_OK;
}

// estimator_kernel.cu -- the reference's sampled-max block-score estimator (SURVEY 8a rows a4 + a5) on tcgen05.
//
// Reference: efficient_attn_with_pooling (W:62-87) = pad_to_multiple (W:25-36) + random_sample_tokens (W:37-60:
// the SAME num_keep = 32 intra-block offsets for every block, drawn per (b,h) and separately for q and k) +
// the Triton kernel _attn_fwd (P:87-199, non-causal) + row normalisation (P:250-251).  With
// s = (q~ . k~) * (1/sqrt(D)) * 1.44269504:
//     R[r, j]  = max_{c in sampled k-block j} s[r, c]           stored in q.dtype          (P:54-55)
//     m[r]     = max_j (fp32 value before the cast)                                         (P:52-56)
//     Po[i, j] = max_{r in sampled q-block i} exp2(R[r, j] - m[r])   stored in q.dtype      (P:72-82, l_i == 1)
//     Po      /= Po.sum(-1)                                          in q.dtype             (P:250-251)
// The Triton kernel materialises R (50 MB at Wan size) and makes two passes; here one CTA owns 128 sampled
// query rows (= 4 q-blocks) of one head, streams all sampled keys through a TMA ring, keeps R for its rows in
// shared memory (bf16, 128 x nb) and never touches HBM for it.
//
//   sample_tokens_kernel   gather of the sampled rows into contiguous [B,H,nb*32,D] tensors (replicate padding
//                          of the ragged last block = clamped row index)
//   sampled_score_kernel   warps 0-3: TMEM -> registers, 32-column segment maxima, running row max, R to smem,
//                          final exp2 / block max / normalisation;  warp 4: TMA producer;  warp 5: MMA issuer
//                          (S = Q~ K~^T, 128x128x D, S double-buffered in TMEM)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace blade {

constexpr int kKeep = 32;   // sampled tokens per block (W:62)
constexpr int kTileR = 128; // sampled rows / keys per MMA tile = 4 blocks
#ifndef BLADE_EST_SBUF
#define BLADE_EST_SBUF 4
#endif
// S buffers in TMEM of the first kernel (2 or 4; measured: no difference -- the kernel is bound by L2 -> SM ingest)
constexpr int kSBuf = BLADE_EST_SBUF;

struct Strides3e {
  int64_t b, h, s;
};

// grid (nb, H, B), 256 threads: 32 sampled rows of q and of k per block, 16-byte chunks
template <int D>
__global__ void __launch_bounds__(256) sample_tokens_kernel(const uint16_t* __restrict__ q, const uint16_t* __restrict__ k,
                                                            Strides3e sq, Strides3e sk, const int32_t* __restrict__ q_off,
                                                            const int32_t* __restrict__ k_off, uint16_t* __restrict__ q_s,
                                                            uint16_t* __restrict__ k_s, int S, int H, int nb, int block) {
  constexpr int LPR = D / 8;
  const int blk = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int64_t bh = static_cast<int64_t>(b) * H + h;
  for (int e = threadIdx.x; e < 2 * kKeep * LPR; e += 256) {
    const int t = e / (kKeep * LPR);           // 0 = q, 1 = k
    const int r = (e / LPR) % kKeep, chunk = e % LPR;
    const int off = __ldg((t ? k_off : q_off) + bh * kKeep + r);
    int row = blk * block + off;
    row = row < S ? row : S - 1;                // replicate padding (W:35)
    const Strides3e st = t ? sk : sq;
    const uint16_t* src = (t ? k : q) + b * st.b + h * st.h + static_cast<int64_t>(row) * st.s;
    uint16_t* dst = (t ? k_s : q_s) + ((bh * nb + blk) * kKeep + r) * D;
    reinterpret_cast<uint4*>(dst)[chunk] = __ldg(reinterpret_cast<const uint4*>(src) + chunk);
  }
}

template <int D>
struct EstSmem {
  static constexpr int kTileBytes = kTileR * D * 2;
  static constexpr int kStages = D == 128 ? 4 : 6;
  static constexpr int kQOff = 0;
  static constexpr int kKOff = kTileBytes;
  static constexpr int kROff = kKOff + kStages * kTileBytes;
  static constexpr int kRStride = 258;               // padded row (129 words): conflict-free per-row access
  static constexpr int kRBytes = kTileR * kRStride * 2;  // nb <= 256
  static constexpr int kMiscOff = kROff + kRBytes;
  static constexpr int kTotal = kMiscOff + 256 + 1024;
};

struct EstMisc {
  uint64_t q_full, k_full[6], k_empty[6], s_full[4], s_empty[4];
  uint32_t tmem_base;
};

// CS = cluster size along the query-tile axis (1 or 2).  The kernel is bound by L2 -> SM traffic (every CTA streams all
// sampled keys of its head: 2 MB x 768 CTAs at Wan size), so with CS = 2 the two CTAs of a cluster -- two query tiles of
// the SAME head -- share every key tile: each issues the TMA for one half of it (tmKh: 64-row boxes) with
// .multicast::cluster to both, and a ring slot is released by BOTH issuers' tcgen05.commit (multicast to the two
// k_empty barriers, count 2).
template <int D, bool IS_BF16, int CS>
__global__ void __launch_bounds__(192, 1)
sampled_score_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmKh, float* __restrict__ scores, int nb, float scale_log2) {
  using L = EstSmem<D>;
  constexpr int kStages = L::kStages;
  constexpr int kTileBytes = L::kTileBytes;
  constexpr int kSub = D / 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + L::kQOff;
  uint8_t* sK = smem + L::kKOff;
  uint16_t* sR = reinterpret_cast<uint16_t*>(smem + L::kROff);  // [128 rows][258] bf16/f16 bits
  EstMisc* mz = reinterpret_cast<EstMisc*>(smem + L::kMiscOff);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int nkt = (nb + 3) / 4;  // key tiles of 4 blocks

  if (threadIdx.x == 0) {
    mbar_init(&mz->q_full, 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&mz->k_full[s], 1);
      mbar_init(&mz->k_empty[s], CS);
    }
    for (int i = 0; i < kSBuf; ++i) {
      mbar_init(&mz->s_full[i], 1);
      mbar_init(&mz->s_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc<kSBuf * 128>(&mz->tmem_base);
  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();   // the peer's barriers are initialised before anything is multicast at them
  tc_fence_after();
  const uint32_t tmem_base = mz->tmem_base;
  const uint32_t crank = CS > 1 ? cluster_ctarank() : 0;

  if (warp == 4) {
    // ------------------------------ TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(&mz->q_full, kTileBytes);
      for (int dh = 0; dh < kSub; ++dh)
        tma_load_4d(sQ + dh * (kTileR * 128), &tmQ, &mz->q_full, dh * 64, qt * kTileR, h, b, kEvictFirst);
    }
    __syncwarp();
    uint32_t slot = 0, ph = 0;
    for (int jt = 0; jt < nkt; ++jt) {
      mbar_wait(&mz->k_empty[slot], ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&mz->k_full[slot], kTileBytes);   // my half + the peer's half
        if (CS > 1) {
          constexpr int kHalf = kTileR / CS;
          for (int dh = 0; dh < kSub; ++dh)
            tma_load_4d_mc(sK + slot * kTileBytes + dh * (kTileR * 128) + crank * (kHalf * 128), &tmKh, &mz->k_full[slot],
                           dh * 64, jt * kTileR + crank * kHalf, h, b, static_cast<uint16_t>((1u << CS) - 1), kEvictLast);
        } else {
          for (int dh = 0; dh < kSub; ++dh)
            tma_load_4d(sK + slot * kTileBytes + dh * (kTileR * 128), &tmK, &mz->k_full[slot], dh * 64, jt * kTileR, h,
                        b, kEvictLast);
        }
      }
      __syncwarp();
      if (++slot == kStages) {
        slot = 0;
        ph ^= 1;
      }
    }
  } else if (warp == 5) {
    // ------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_f16(kTileR, kTileR, IS_BF16, false, false);
    const uint32_t sQ_addr = smem_u32(sQ), sK_addr = smem_u32(sK);
    mbar_wait(&mz->q_full, 0);
    uint32_t slot = 0, ph = 0;
    for (int jt = 0; jt < nkt; ++jt) {
      const int bsel = jt % kSBuf;
      mbar_wait(&mz->s_empty[bsel], ((jt / kSBuf) & 1) ^ 1);  // the reducers have drained this S buffer
      mbar_wait(&mz->k_full[slot], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t adesc = make_smem_desc(sQ_addr, 16, 1024, 2);
        const uint64_t bdesc = make_smem_desc(sK_addr + slot * kTileBytes, 16, 1024, 2);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t koff = static_cast<uint32_t>((k >> 2) * (kTileR * 128 / 16) + (k & 3) * 2);
          umma_ss(tmem_base + bsel * kTileR, adesc + koff, bdesc + koff, idesc, k > 0);
        }
        if (CS > 1) tc_commit_mc(&mz->k_empty[slot], static_cast<uint16_t>((1u << CS) - 1));
        else tc_commit(&mz->k_empty[slot]);
        tc_commit(&mz->s_full[bsel]);
      }
      __syncwarp();
      if (++slot == kStages) {
        slot = 0;
        ph ^= 1;
      }
    }
  } else {
    // ------------------------------ reducers: one sampled query row per thread
    const int row = warp * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    uint16_t* myR = sR + row * L::kRStride;
    float m = -INFINITY;
    for (int jt = 0; jt < nkt; ++jt) {
      const int bsel = jt % kSBuf;
      mbar_wait(&mz->s_full[bsel], (jt / kSBuf) & 1);
      tc_fence_after();
      uint32_t s[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(tmem_base + lane_base + bsel * kTileR + c * 32, s[c]);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&mz->s_empty[bsel]);  // S is in registers: the buffer can be overwritten
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          mx0 = fmaxf(mx0, fmaxf(__uint_as_float(s[c][i]), __uint_as_float(s[c][i + 1])));
          mx1 = fmaxf(mx1, fmaxf(__uint_as_float(s[c][i + 2]), __uint_as_float(s[c][i + 3])));
        }
        const int j = jt * 4 + c;
        float bm = fmaxf(mx0, mx1) * scale_log2;  // P:52 `tl.max(qk, 1) * qk_scale`
        if (j >= nb) bm = -INFINITY;              // zero-filled key blocks beyond the sequence
        m = fmaxf(m, bm);
        if (j < 256) myR[j] = IS_BF16 ? __bfloat16_as_ushort(__float2bfloat16_rn(bm)) : __half_as_ushort(__float2half_rn(bm));
      }
    }
    // Po[i, j] = max over the 32 rows of q-block i (= this warp) of exp2(R[r,j] - m[r]), stored in q.dtype (P:72-82).
    // Rounding is monotonic, so each thread first overwrites its own R row in place with the ROUNDED exponentials
    // (one MUFU per entry, no cross-lane traffic); then lane L takes the column maxima of columns j = 32*jj + L over
    // the 32 rows of its warp's block straight from shared memory (conflict-free: consecutive lanes read
    // consecutive 16-bit words of one padded row).
    const int qblk = qt * 4 + warp;
    for (int j = 0; j < nb; ++j) {
      const float r = IS_BF16 ? __bfloat162float(__ushort_as_bfloat16(myR[j])) : __half2float(__ushort_as_half(myR[j]));
      const float e = ex2_approx(r - m);
      myR[j] = IS_BF16 ? __bfloat16_as_ushort(__float2bfloat16_rn(e)) : __half_as_ushort(__float2half_rn(e));
    }
    __syncwarp();
    float po[8];
    float sum = 0.f;
    const uint16_t* blkR = sR + (warp * 32) * L::kRStride;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int j = jj * 32 + lane;
      float mxv = 0.f;  // exponentials are >= 0
      if (j < nb) {
#pragma unroll 8
        for (int rr = 0; rr < 32; ++rr) {
          const uint16_t u = blkR[rr * L::kRStride + j];
          mxv = fmaxf(mxv, IS_BF16 ? __bfloat162float(__ushort_as_bfloat16(u)) : __half2float(__ushort_as_half(u)));
        }
        sum += mxv;
      }
      po[jj] = mxv;
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    // P:250-251 in q.dtype: Sum = sum(Po) (fp32 accumulate, rounded), Po /= Sum (rounded)
    const float sum_r = IS_BF16 ? bf16_round(sum) : __half2float(__float2half_rn(sum));
    if (qblk < nb) {
      float* out = scores + ((static_cast<int64_t>(b) * gridDim.y + h) * nb + qblk) * nb;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int j = jj * 32 + lane;
        if (j < nb) {
          const float v = po[jj] / sum_r;
          out[j] = IS_BF16 ? bf16_round(v) : __half2float(__float2half_rn(v));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();   // nobody leaves while the peer may still multicast into this CTA
  if (warp == 5) tmem_dealloc<kSBuf * 128>(tmem_base);
}


// ------------------------------------------------------------------------------------------------
// v2: the same arithmetic with the shared memory spent on the key ring ONLY:
//   * Q lives in TMEM: every reducer thread loads its own sampled query row from global memory once and stores it as
//     the A operand (row = lane, 2 elements per 32-bit column), the score MMA runs TS-mode (A from TMEM, B = key tile);
//   * R lives in TMEM: 4 block maxima per tile = two 32-bit columns per row, written with tcgen05.st, read back once
//     (128 columns) for the exp2 / block-max epilogue, which stages the exponentials through the then idle ring;
//   * the ring takes all of shared memory: 7 stages of 32 KB (d = 128) / 14 of 16 KB (d = 64).
// Measured (Wan size, same box): v1 0.239 ms -> v2 0.221 ms.  What bounds both (profiles/r02z_estimator_cluster.txt):
// every SM ingests all sampled keys of its head, 2 MB per CTA in ~30 us = 67 GB/s per SM -- the per-SM share of the
// L2 -> SM fabric (the attention kernel tops out at 61 GB/s per SM).  Three experiments agree: TMA multicast inside a
// 2-CTA cluster halves the L2 reads but not the bytes each SM receives (no gain); 7 ring stages instead of 4 (this
// kernel: 7 %); 4 S buffers instead of 2 in v1 (no gain).  Only fewer bytes per SM per FLOP help: cta_group::2 MMAs (each
// SM of a pair holds half of every key tile) -- not built.
// TMEM columns: [0,256) S double buffer | [256, 256 + D/2) Q | [320, 448) R.
// ------------------------------------------------------------------------------------------------
template <int D>
struct EstSmem2 {
  static constexpr int kTileBytes = kTileR * D * 2;
  static constexpr int kStages = (224 * 1024) / kTileBytes;
  static constexpr int kRStride = 258;
  static constexpr int kMiscOff = kStages * kTileBytes;
  static constexpr int kTotal = kMiscOff + 512 + 1024;
  static_assert(kTileR * kRStride * 2 <= kStages * kTileBytes, "the epilogue's staging area lives in the ring");
};
struct EstMisc2 {
  uint64_t q_ready, k_full[14], k_empty[14], s_full[2], s_empty[2];
  uint32_t tmem_base;
};
constexpr uint32_t kEstColQ = 256, kEstColR = 320;

template <int D, bool IS_BF16>
__global__ void __launch_bounds__(192, 1)
sampled_score_kernel_v2(const uint16_t* __restrict__ q_s, const __grid_constant__ CUtensorMap tmK, float* __restrict__ scores,
                        int nb, float scale_log2) {
  using L = EstSmem2<D>;
  constexpr int kStages = L::kStages;
  constexpr int kTileBytes = L::kTileBytes;
  constexpr int kSub = D / 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint16_t* sR = reinterpret_cast<uint16_t*>(smem);  // epilogue only: [128 rows][258] bf16/f16 bits, over the idle ring
  EstMisc2* mz = reinterpret_cast<EstMisc2*>(smem + L::kMiscOff);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int nkt = (nb + 3) / 4;  // key tiles of 4 blocks
  const int64_t Ss = static_cast<int64_t>(nb) * kKeep;

  if (threadIdx.x == 0) {
    mbar_init(&mz->q_ready, 4);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&mz->k_full[s], 1);
      mbar_init(&mz->k_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&mz->s_full[i], 1);
      mbar_init(&mz->s_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc<512>(&mz->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = mz->tmem_base;

  if (warp == 4) {
    // ------------------------------ TMA producer
    uint32_t slot = 0, ph = 0;
    for (int jt = 0; jt < nkt; ++jt) {
      mbar_wait(&mz->k_empty[slot], ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&mz->k_full[slot], kTileBytes);
        for (int dh = 0; dh < kSub; ++dh)
          tma_load_4d(sK + slot * kTileBytes + dh * (kTileR * 128), &tmK, &mz->k_full[slot], dh * 64, jt * kTileR, h,
                      b, kEvictLast);
      }
      __syncwarp();
      if (++slot == kStages) {
        slot = 0;
        ph ^= 1;
      }
    }
  } else if (warp == 5) {
    // ------------------------------ MMA issuer: S = Q (TMEM) x K^T (ring)
    constexpr uint32_t idesc = make_idesc_f16(kTileR, kTileR, IS_BF16, false, false);
    const uint32_t sK_addr = smem_u32(sK);
    mbar_wait(&mz->q_ready, 0);
    tc_fence_after();
    uint32_t slot = 0, ph = 0;
    for (int jt = 0; jt < nkt; ++jt) {
      const int bsel = jt & 1;
      mbar_wait(&mz->s_empty[bsel], ((jt >> 1) & 1) ^ 1);  // the reducers have drained this S buffer
      mbar_wait(&mz->k_full[slot], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t bdesc = make_smem_desc(sK_addr + slot * kTileBytes, 16, 1024, 2);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t koff = static_cast<uint32_t>((k >> 2) * (kTileR * 128 / 16) + (k & 3) * 2);
          umma_ts(tmem_base + bsel * kTileR, tmem_base + kEstColQ + k * 8, bdesc + koff, idesc, k > 0);
        }
        tc_commit(&mz->k_empty[slot]);
        tc_commit(&mz->s_full[bsel]);
      }
      __syncwarp();
      if (++slot == kStages) {
        slot = 0;
        ph ^= 1;
      }
    }
  } else {
    // ------------------------------ reducers: one sampled query row per thread
    const int row = warp * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    {
      // my query row -> TMEM as the MMA's A operand (rows beyond the sampled sequence are zero, like TMA's fill)
      const int64_t grow = static_cast<int64_t>(qt) * kTileR + row;
      const uint4* src = reinterpret_cast<const uint4*>(
          q_s + ((static_cast<int64_t>(b) * gridDim.y + h) * Ss + (grow < Ss ? grow : 0)) * D);
#pragma unroll
      for (int c = 0; c < D / 64; ++c) {
        uint32_t w[32];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          uint4 v = make_uint4(0u, 0u, 0u, 0u);
          if (grow < Ss) v = __ldg(src + c * 8 + u);
          w[4 * u] = v.x;
          w[4 * u + 1] = v.y;
          w[4 * u + 2] = v.z;
          w[4 * u + 3] = v.w;
        }
        tmem_st32(tmem_base + lane_base + kEstColQ + c * 32, w);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&mz->q_ready);
    }
    float m = -INFINITY;
    for (int jt = 0; jt < nkt; ++jt) {
      const int bsel = jt & 1;
      mbar_wait(&mz->s_full[bsel], (jt >> 1) & 1);
      tc_fence_after();
      uint32_t s[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(tmem_base + lane_base + bsel * kTileR + c * 32, s[c]);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&mz->s_empty[bsel]);  // S is in registers: the buffer can be overwritten
      uint32_t rb[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          mx0 = fmaxf(mx0, fmaxf(__uint_as_float(s[c][i]), __uint_as_float(s[c][i + 1])));
          mx1 = fmaxf(mx1, fmaxf(__uint_as_float(s[c][i + 2]), __uint_as_float(s[c][i + 3])));
        }
        const int j = jt * 4 + c;
        float bm = fmaxf(mx0, mx1) * scale_log2;  // P:52 `tl.max(qk, 1) * qk_scale`
        if (j >= nb) bm = -INFINITY;              // zero-filled key blocks beyond the sequence
        m = fmaxf(m, bm);
        rb[c] = IS_BF16 ? __bfloat16_as_ushort(__float2bfloat16_rn(bm)) : __half_as_ushort(__float2half_rn(bm));
      }
      // R[row, 4 jt .. 4 jt + 3] (rounded to q.dtype, P:54-55) -> two TMEM columns of my lane
      tmem_st2(tmem_base + lane_base + kEstColR + jt * 2, rb[0] | (rb[1] << 16), rb[2] | (rb[3] << 16));
    }
    // Po[i, j] = max over the 32 rows of q-block i (= this warp) of exp2(R[r,j] - m[r]), stored in q.dtype (P:72-82).
    // Each thread turns its own R row into ROUNDED exponentials (rounding is monotonic) and writes them to the staging
    // area -- the ring is idle: the last s_full means every MMA, hence every TMA load, has completed -- then lane L takes
    // the maxima of columns j = 32 jj + L over the 32 rows of its warp's block.
    tmem_wait_st();
    const int qblk = qt * 4 + warp;
    uint16_t* myR = sR + row * L::kRStride;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_base + lane_base + kEstColR + c * 32, r);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int j = (c * 32 + i) * 2;
        if (j < nb) {   // warp-uniform
          const uint16_t u0 = static_cast<uint16_t>(r[i] & 0xFFFFu), u1 = static_cast<uint16_t>(r[i] >> 16);
          const float r0 = IS_BF16 ? __bfloat162float(__ushort_as_bfloat16(u0)) : __half2float(__ushort_as_half(u0));
          const float r1 = IS_BF16 ? __bfloat162float(__ushort_as_bfloat16(u1)) : __half2float(__ushort_as_half(u1));
          const float e0 = ex2_approx(r0 - m), e1 = ex2_approx(r1 - m);
          myR[j] = IS_BF16 ? __bfloat16_as_ushort(__float2bfloat16_rn(e0)) : __half_as_ushort(__float2half_rn(e0));
          myR[j + 1] = IS_BF16 ? __bfloat16_as_ushort(__float2bfloat16_rn(e1)) : __half_as_ushort(__float2half_rn(e1));
        }
      }
    }
    __syncwarp();
    float po[8];
    float sum = 0.f;
    const uint16_t* blkR = sR + (warp * 32) * L::kRStride;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int j = jj * 32 + lane;
      float mxv = 0.f;  // exponentials are >= 0
      if (j < nb) {
#pragma unroll 8
        for (int rr = 0; rr < 32; ++rr) {
          const uint16_t u = blkR[rr * L::kRStride + j];
          mxv = fmaxf(mxv, IS_BF16 ? __bfloat162float(__ushort_as_bfloat16(u)) : __half2float(__ushort_as_half(u)));
        }
        sum += mxv;
      }
      po[jj] = mxv;
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    // P:250-251 in q.dtype: Sum = sum(Po) (fp32 accumulate, rounded), Po /= Sum (rounded)
    const float sum_r = IS_BF16 ? bf16_round(sum) : __half2float(__float2half_rn(sum));
    if (qblk < nb) {
      float* out = scores + ((static_cast<int64_t>(b) * gridDim.y + h) * nb + qblk) * nb;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int j = jj * 32 + lane;
        if (j < nb) {
          const float v = po[jj] / sum_r;
          out[j] = IS_BF16 ? bf16_round(v) : __half2float(__float2half_rn(v));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<512>(tmem_base);
}

}  // namespace blade

using namespace blade;

extern "C" int blade_asa_sample_tokens(const BladeTensor* q, const BladeTensor* k, const int32_t* q_off,
                                       const int32_t* k_off, void* q_s, void* k_s, int32_t block_size, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int e = check_tensor16(q, "q")) return e;
  if (int e = check_tensor16(k, "k")) return e;
  BLADE_REQUIRE(q_off && k_off && q_s && k_s, BLADE_ERR_ARG, "null pointer");
  BLADE_REQUIRE(block_size == 64 || block_size == 128, BLADE_ERR_ARG, "block_size %d not in {64,128}", block_size);
  const int64_t B = q->shape[0], H = q->shape[1], S = q->shape[2], D = q->shape[3];
  for (int i = 0; i < 4; ++i) BLADE_REQUIRE(k->shape[i] == q->shape[i], BLADE_ERR_SHAPE, "q/k shapes differ");
  const int nb = static_cast<int>(ceil_div(S, block_size));
  Strides3e sq{q->stride[0], q->stride[1], q->stride[2]}, sk{k->stride[0], k->stride[1], k->stride[2]};
  dim3 grid(nb, static_cast<unsigned>(H), static_cast<unsigned>(B));
  if (D == 128)
    sample_tokens_kernel<128><<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(q->ptr), static_cast<const uint16_t*>(k->ptr), sq, sk,
                                                        q_off, k_off, static_cast<uint16_t*>(q_s), static_cast<uint16_t*>(k_s),
                                                        (int)S, (int)H, nb, block_size);
  else
    sample_tokens_kernel<64><<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(q->ptr), static_cast<const uint16_t*>(k->ptr), sq, sk,
                                                       q_off, k_off, static_cast<uint16_t*>(q_s), static_cast<uint16_t*>(k_s),
                                                       (int)S, (int)H, nb, block_size);
  BLADE_CUDA_OK(cudaGetLastError());
  return BLADE_OK;
}

extern "C" int blade_asa_scores_sampled(const void* q_s, const void* k_s, float* scores, int64_t B, int64_t H, int64_t nb,
                                        int64_t D, int32_t dtype, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BLADE_REQUIRE(q_s && k_s && scores, BLADE_ERR_ARG, "null pointer");
  BLADE_REQUIRE(D == 64 || D == 128, BLADE_ERR_SHAPE, "head dim %lld not in {64,128}", (long long)D);
  BLADE_REQUIRE(nb >= 1 && nb <= 256, BLADE_ERR_SHAPE, "sampled estimator supports nb <= 256 blocks (got %lld)", (long long)nb);
  BLADE_REQUIRE(dtype == BLADE_BF16 || dtype == BLADE_F16, BLADE_ERR_DTYPE, "dtype");
  const int64_t Ss = nb * kKeep;
  CUtensorMap tmQ, tmK, tmKh;
  if (int e = make_tmap(&tmQ, q_s, dtype, B, H, Ss, D, H * Ss * D, Ss * D, D)) return e;
  if (int e = make_tmap(&tmK, k_s, dtype, B, H, Ss, D, H * Ss * D, Ss * D, D)) return e;
  if (int e = make_tmap(&tmKh, k_s, dtype, B, H, Ss, D, H * Ss * D, Ss * D, D, kTileR / 2)) return e;
  StageTimer timer(1, stream);
  // v2 (Q and R in TMEM, the ring takes all of shared memory) is the default; BLADE_EST_V1=1 selects the first kernel (A/B)
  static const bool use_v1 = getenv("BLADE_EST_V1") && atoi(getenv("BLADE_EST_V1")) != 0;
  if (!use_v1) {
    dim3 grid2(static_cast<unsigned>(ceil_div(nb, 4)), static_cast<unsigned>(H), static_cast<unsigned>(B));
    const float sl2 = (1.0f / sqrtf(static_cast<float>(D))) * 1.44269504f;  // P:163 literal
#define LAUNCH_EST2(DD, BF)                                                                                          \
  do {                                                                                                               \
    auto kern = sampled_score_kernel_v2<DD, BF>;                                                                     \
    BLADE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, EstSmem2<DD>::kTotal));    \
    kern<<<grid2, 192, EstSmem2<DD>::kTotal, stream>>>(static_cast<const uint16_t*>(q_s), tmK, scores, (int)nb, sl2); \
  } while (0)
    const bool bf2 = dtype == BLADE_BF16;
    if (D == 128) { if (bf2) LAUNCH_EST2(128, true); else LAUNCH_EST2(128, false); }
    else          { if (bf2) LAUNCH_EST2(64, true); else LAUNCH_EST2(64, false); }
#undef LAUNCH_EST2
    BLADE_CUDA_OK(cudaGetLastError());
    return BLADE_OK;
  }
  // BLADE_EST_CLUSTER=2: clusters of two query tiles of one head share the key tiles (TMA multicast).  Measured
  // (profiles/r02z_estimator_cluster.txt): L2 reads halve (1.6 -> 0.85 GB) but the kernel does not get faster (0.240 vs
  // 0.244 ms) -- every SM still ingests each full 32 KB tile (see the note above the v2 kernel).  Default: no cluster.
  static const int env_cs = getenv("BLADE_EST_CLUSTER") ? atoi(getenv("BLADE_EST_CLUSTER")) : 1;
  const int cs = env_cs == 2 ? 2 : 1;
  const unsigned qtiles = static_cast<unsigned>(ceil_div(nb, 4));
  const float scale_log2 = (1.0f / sqrtf(static_cast<float>(D))) * 1.44269504f;  // P:163 literal
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3((qtiles + cs - 1) / cs * cs, static_cast<unsigned>(H), static_cast<unsigned>(B));  // an odd last tile is padded
  lc.blockDim = dim3(192);
  lc.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  lc.attrs = at;
  lc.numAttrs = 1;
  const int nbi = static_cast<int>(nb);
#define LAUNCH_EST(DD, BF, CSZ)                                                                                     \
  do {                                                                                                              \
    auto kern = sampled_score_kernel<DD, BF, CSZ>;                                                                  \
    lc.dynamicSmemBytes = EstSmem<DD>::kTotal;                                                                      \
    BLADE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, EstSmem<DD>::kTotal));    \
    BLADE_CUDA_OK(cudaLaunchKernelEx(&lc, kern, tmQ, tmK, tmKh, scores, nbi, scale_log2));                          \
  } while (0)
#define LAUNCH_EST_D(DD, BF) do { if (cs == 2) LAUNCH_EST(DD, BF, 2); else LAUNCH_EST(DD, BF, 1); } while (0)
  const bool bf = dtype == BLADE_BF16;
  if (D == 128) { if (bf) LAUNCH_EST_D(128, true); else LAUNCH_EST_D(128, false); }
  else          { if (bf) LAUNCH_EST_D(64, true); else LAUNCH_EST_D(64, false); }
#undef LAUNCH_EST_D
#undef LAUNCH_EST
  BLADE_CUDA_OK(cudaGetLastError());
  return BLADE_OK;
}

/*
 * blade_asa.h -- C ABI of the B200-native Adaptive Sparse Attention (ASA) hot path.
 *
 * Drop-in boundary for the kernels Video-BLADE's ASA module calls (reference paths relative to the
 * VIDEO-BLADE tree; W = wanx/train/special_attentions_local/TrainRelated/wanx_blocksparseattn.py,
 * C = cogvideox/.../cogvideo_blocksparseattn.py, P = .../attn_pooling_kernel.py):
 *
 *   reference interface                                              replaced by
 *   ---------------------------------------------------------------  -----------------------------
 *   GilbertRearranger.__init__ / gilbert3d        (W:102-129, G:6-167)  blade_gilbert_tables
 *   index_select x3 + pad_to_multiple + simple_pooling
 *                                     (W:25-36, 88-93, 142-152, 344-345)  blade_asa_prep
 *   attn_with_pooling (Triton _attn_fwd)  (W:62-87 -> P:201-253)         blade_asa_scores_meanpool
 *                                                                         blade_asa_scores_sampled
 *   transfer_attn_to_mask(mode="energy")  (W:162-233, C:177-249)         blade_asa_select
 *   block_sparse_attn -> block_sparse_attn_func (external CUDA lib)
 *                                     (W:278-309, call site W:301-305)    blade_block_sparse_attn_fwd
 *   standard_attn + LSE merge + reversed_rearrange
 *                                     (W:21-24, 348-370, 154-159)         blade_asa_attn_fwd
 *   AdaptiveBlockSparseAttnTrain.forward  (W:383-408, C:405-427)         blade_asa_forward
 *   processor steps in front of the module, optionally fused into the gather (blade_asa_forward via
 *   BladeAsaConfig): rotary embedding (modify_wan.py:108-116 / modify_cogvideo.py:59-64) -> rope_cos_sin;
 *   q/k normalisation (modify_wan.py:99-102 RMSNorm over all heads / modify_cogvideo.py:54-57 LayerNorm
 *   per head) -> qk_norm (BladeQkNorm), statistic helper blade_qk_rms_stat
 *
 * Conventions
 *   - plain C, no C++/torch types; every pointer marked "device" is a CUDA device pointer owned by the
 *     caller.  The library never allocates, frees or retains device memory: scratch comes from the
 *     caller-provided workspace (size from blade_asa_workspace_bytes).
 *   - all calls are asynchronous on `stream` (a cudaStream_t passed as void*); no host sync.
 *   - return 0 on success, a BLADE_ERR_* code otherwise; blade_last_error() gives the message of the
 *     last failure on the calling thread.  No C++ exception crosses the boundary.
 *   - tensors are [B, H, S, D] with ELEMENT strides (the reference passes transposed views of
 *     [B, S, H, D] memory, modify_wan.py:104-106); the last dim must be contiguous.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef BLADE_ASA_H_
#define BLADE_ASA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BLADE_ABI_VERSION 2

enum BladeStatus {
  BLADE_OK = 0,
  BLADE_ERR_SHAPE = 1,      /* inconsistent / unsupported shape (W:250-251 asserts) */
  BLADE_ERR_DTYPE = 2,      /* unsupported element type */
  BLADE_ERR_ALIGN = 3,      /* pointer or stride not 16-byte aligned / last dim not contiguous */
  BLADE_ERR_LAUNCH = 4,     /* CUDA launch / driver failure */
  BLADE_ERR_ARG = 5,        /* bad argument (e.g. unknown mode, W:193-194,232 raise ValueError) */
  BLADE_ERR_WORKSPACE = 6,  /* workspace too small */
  BLADE_ERR_NO_DEVICE = 7   /* no sm_100 device */
};

enum BladeDType { BLADE_BF16 = 0, BLADE_F16 = 1, BLADE_F32 = 2, BLADE_I32 = 3, BLADE_U8 = 4 };

typedef struct BladeTensor {
  void* ptr;          /* device */
  int64_t shape[4];   /* B, H, S, D */
  int64_t stride[4];  /* element strides; stride[3] must be 1 */
  int32_t dtype;      /* BladeDType */
  int32_t _pad;
} BladeTensor;

/* q/k normalisation of the attention processor, fused into the gather (MW:99-102: `attn.norm_q(query)`,
 * `attn.norm_k(key)` between the projections and the head split; Wan2.1 uses RMSNorm over all heads' channels).
 * kind 1: y = x * rsqrt(mean_c(x^2) + eps) * w in fp32, rounded once to the tensor dtype;
 * kind 2: diffusers' RMSNorm rounding: (x * rstd) -> tensor dtype, then * w -> tensor dtype.
 * q and k must be token-major ([B,S,H*D] memory: stride_h == D, stride_s == H*D); weights [H*D] in the tensor dtype.
 * kind 3: LayerNorm over the D channels of every head (CogVideoX, MC:54-57: `attn.norm_q(query)` after the head
 * split): y = (x - mean) * rsqrt(var + eps) * w + b in fp32, rounded once; weights / biases [D] in the tensor dtype
 * (bias pointers may be NULL); any q/k strides. */
typedef struct BladeQkNorm {
  int32_t kind;          /* 0 = none */
  float eps;
  const void* q_weight;  /* device */
  const void* k_weight;  /* device */
  const float* rstd;     /* optional device fp32 [2][B*S] (q then k), indexed by TOKEN: a statistic computed elsewhere
                            (blade_qk_rms_stat on the token shard + all-gather, when the heads are sharded); NULL =
                            computed here, which needs token-major q/k */
  const void* q_bias;    /* kind 3 only, device or NULL */
  const void* k_bias;
} BladeQkNorm;

/* Sequence-parallel (Ulysses) data plane over NVLink peer memory (absent from the reference, whose only
 * multi-GPU mode is whole-pipeline replicas, simple_multiprocess_sampler.py:296-339).  Outside attention rank r of a
 * group of n_peers owns rows_per_peer consecutive TOKENS of every head ([rows_per_peer, H_total, D] token-major
 * memory); inside it owns all tokens of H_total / n_peers heads.  Instead of an all-to-all in front of and behind the
 * layer, the gather kernel PULLS its heads' rows of q/k/v straight out of the peers' projection outputs, and the
 * attention epilogue PUSHES every output row into the owning peer's [rows_per_peer, H_total, D] buffer, tile by tile
 * while the tensor cores work on the next tile.  All pointers are device addresses valid on THIS GPU (peer mappings:
 * CUDA IPC / VMM, e.g. torch.distributed._symmetric_memory); the caller provides the cross-GPU barriers
 * (peers' q/k/v complete before the call is enqueued; outputs visible after a barrier behind it). */
#define BLADE_MAX_PEERS 8
typedef struct BladePeers {
  int32_t n_peers;         /* Ulysses degree P (1..8) */
  int32_t my_peer;         /* my index in the group: I own heads [my_peer*H, (my_peer+1)*H) of H_total = n_peers*H */
  int32_t rows_per_peer;   /* tokens owned by each peer: token t lives on peer t / rows_per_peer */
  int32_t _pad;
  const void* q[BLADE_MAX_PEERS];  /* peer p's [rows_per_peer, H_total, D] projections; NULL = q/k/v arguments are local */
  const void* k[BLADE_MAX_PEERS];
  const void* v[BLADE_MAX_PEERS];
  void* out[BLADE_MAX_PEERS];      /* peer p's [rows_per_peer, H_total, D] attention output; NULL = `out` argument */
} BladePeers;

/* Knobs of the reference module (W:9-16, C:9-16 and the literals W:62,325,341). */
typedef struct BladeAsaConfig {
  int32_t block_size;        /* 128 (W:325); 64 accepted for the selection/score stages */
  int32_t sample_gap;        /* 30 wan / 15 cog (W:15); 0 disables the pooled branch */
  int32_t min_retain;        /* max(1,int(nb*min_retain_ratio)) computed by the host (W:215 / C:230) */
  int32_t max_retain;        /* max(1,int(nb*max_retain_ratio))                      (W:216 / C:231) */
  float energy_threshold;    /* 0.95 (W:341) */
  int32_t force_last;        /* 0 wan; 2 cog: last two block rows/cols forced on (C:247-248) */
  int32_t num_keep;          /* 32 (W:62) -- sampled estimator only */
  int32_t estimator;         /* 0 = block-mean-pool (north-star kernel (a)), 1 = sampled-max (P) */
  int32_t exact_merge;       /* 1 = reproduce the reference's bf16 op chain W:351-370 (default) */
  int32_t rope_first_row;    /* first source row that is rotated (0 wan; text_length cog, MC:59-64) */
  const float* rope_cos_sin; /* device, fp32 [rows, D/2, 2] (cos, sin) per token and pair, or NULL: rotary embedding
                                fused into the gather, applied to q and k where the processor does (MW:108-116) */
  const BladeQkNorm* qk_norm; /* host pointer or NULL (blade_asa_forward only; needs the output copies, i.e. a gather) */
  const int32_t* token_row;   /* optional device int32 [S]: token index of output row r, for the rotary table and the
                                 norm statistic, when src_row addresses a packed buffer instead of tokens (Ulysses
                                 receive layout); NULL = src_row[r] (or r) is the token index */
  /* ---- ABI version 2 ---- */
  const int32_t* sample_q_off; /* estimator 1: device int32 [B,H,num_keep] intra-block offsets of the sampled query tokens */
  const int32_t* sample_k_off; /*   ... and key tokens (the reference draws them with torch.rand + topk per call, W:49-51) */
  int32_t select_rounding;     /* prefix sums / threshold of the selection: 0 = fp32 scores (fp64-sequential sums rounded
                                  to fp32 == torch CPU cumsum of fp32); 1 = bf16, 2 = f16: the reference's arithmetic when
                                  Po stays in the model dtype (W:214-221: fp32-sequential sums, each prefix and
                                  `energy_threshold * total` rounded to that dtype) */
  int32_t _pad2;
  unsigned long long* selected_acc; /* optional device counter: += number of selected (q-block, k-block) pairs of this
                                  call -- the numerator of the reference's sparsity statistic (W:372) without its
                                  per-layer host sync (W:398) */
  const BladePeers* peers;     /* host pointer or NULL: Ulysses pull/push over peer memory (see BladePeers) */
} BladeAsaConfig;

/* ---- introspection -------------------------------------------------------------------------- */
int blade_abi_version(void);
const char* blade_last_error(void);
/* Returns BLADE_OK when a CUDA device with compute capability 10.x is current. */
int blade_device_check(void);

/* ---- a2: Gilbert curve tables (host memory, int64[w*h*d] each) ------------------------------ */
int blade_gilbert_tables(int32_t width, int32_t height, int32_t depth,
                         int64_t* curve2raster, int64_t* raster2curve);

/* ---- workspace ------------------------------------------------------------------------------ */
size_t blade_asa_workspace_bytes(int64_t B, int64_t H, int64_t S, int64_t D, const BladeAsaConfig* cfg);
/* workspace of the stand-alone attention entry points below (item counter + one parked pooled-branch tile per SM
 * and stream; 16-byte aligned).  blade_block_sparse_attn*_fwd accept NULL: items are then assigned round-robin
 * instead of being claimed from the counter. */
size_t blade_attn_workspace_bytes(int64_t D);

/* ---- prep: gather into curve order + block means + gap-pooled K/V ---------------------------
 * src_row (device int32[S], may be NULL = identity): row of q/k/v that lands at output row r
 *   (GilbertRearranger.rearrange, W:142-152 / C:141-154).
 * q_r,k_r,v_r: contiguous [B,H,S,D] copies in output order (ptr may be NULL when src_row == NULL and
 *   no copy is wanted).  q_mean,k_mean: fp32 [B,H,nb,D] means over replicate-padded blocks (W:25-36).
 * k_pool,v_pool: [B,H,ceil(S/gap),D] in the input dtype, fp32-accumulated means (W:88-93). */
int blade_asa_prep(const BladeTensor* q, const BladeTensor* k, const BladeTensor* v,
                   const int32_t* src_row,
                   void* q_r, void* k_r, void* v_r,
                   float* q_mean, float* k_mean,
                   void* k_pool, void* v_pool,
                   int32_t block_size, int32_t sample_gap, void* stream);

/* Same, with the rotary embedding of the processor (MW:108-116 / MC:59-64) applied to q and k on the fly:
 * rope_cos_sin fp32 [rows, D/2, 2]; source rows < rope_first_row are left alone.  Needs q_r/k_r/v_r. */
int blade_asa_prep_rope(const BladeTensor* q, const BladeTensor* k, const BladeTensor* v,
                        const int32_t* src_row, void* q_r, void* k_r, void* v_r,
                        float* q_mean, float* k_mean, void* k_pool, void* v_pool,
                        int32_t block_size, int32_t sample_gap,
                        const float* rope_cos_sin, int32_t rope_first_row, void* stream);

/* rstd[token] = rsqrt(mean over all H*D channels of x^2 + eps) for q and k (token-major [B,S,H*D] memory):
 * rstd_out fp32 [2][B*S].  The statistic half of BladeQkNorm, for callers that shard heads after computing it. */
int blade_qk_rms_stat(const BladeTensor* q, const BladeTensor* k, float eps, float* rstd_out, void* stream);
/* Ulysses variant: q, k are MY token shard ([1,H_total,rows,D] views of token-major memory); the statistic is stored into
 * every peer's full-length table rstd_peers[p] (fp32 [2][total_rows], peer-mapped device pointers) at rows
 * [first_row, first_row + rows) -- an all-gather done by the producing kernel's own stores over NVLink. */
int blade_qk_rms_stat_peers(const BladeTensor* q, const BladeTensor* k, float eps, float* const* rstd_peers,
                            int32_t n_peers, int64_t total_rows, int64_t first_row, void* stream);

/* ---- score estimators: fp32 [B,H,nb,nb] row-normalised block scores -------------------------- */
int blade_asa_scores_meanpool(const float* q_mean, const float* k_mean, float* scores,
                              int64_t B, int64_t H, int64_t nb, int64_t D, void* stream);

/* ---- a4 + a5: the reference's sampled-max estimator (efficient_attn_with_pooling, W:62-87 -> P:201-253) -----
 * blade_asa_sample_tokens: q_off/k_off device int32 [B,H,32] = intra-block offsets of the sampled tokens (the
 *   reference draws them with torch.rand + topk per call, W:49-51: same offsets for every block of a (b,h));
 *   q_s/k_s: contiguous [B,H,nb*32,D] outputs (ragged last block padded by replicating the last token, W:35).
 * blade_asa_scores_sampled: fp32 [B,H,nb,nb] holding the reference's q.dtype-rounded Po (nb <= 256). */
int blade_asa_sample_tokens(const BladeTensor* q, const BladeTensor* k, const int32_t* q_off, const int32_t* k_off,
                            void* q_s, void* k_s, int32_t block_size, void* stream);
int blade_asa_scores_sampled(const void* q_s, const void* k_s, float* scores, int64_t B, int64_t H, int64_t nb,
                             int64_t D, int32_t dtype, void* stream);

/* ---- a6: energy-threshold block selection ---------------------------------------------------
 * scores fp32 [B,H,nq,nk] -> idx int32 [B,H,nq,nk] (ascending block ids, -1 padded),
 * cnt int32 [B,H,nq], optional mask u8 [B,H,nq,nk], optional kcut int32 [B,H,nq].
 * Order: value descending, block index ascending; prefix sums fp64-sequential rounded to fp32
 * (== torch CPU sort(stable)+cumsum).  lo_bh/hi_bh: optional device int32[B*H] per-head bounds
 * (C:230-231); NULL -> cfg->min_retain / max_retain. */
int blade_asa_select(const float* scores, int64_t B, int64_t H, int64_t nq, int64_t nk,
                     const BladeAsaConfig* cfg, const int32_t* lo_bh, const int32_t* hi_bh,
                     int32_t* idx, int32_t* cnt, uint8_t* mask_opt, int32_t* kcut_opt, void* stream);

/* bool block mask [B,H,nq,nk] (u8) -> idx/cnt in the same format (for callers that bring a mask). */
int blade_mask_to_index(const uint8_t* mask, int64_t B, int64_t H, int64_t nq, int64_t nk,
                        int32_t* idx, int32_t* cnt, void* stream);

/* ---- a7/a8: block-sparse FlashAttention forward (replaces block_sparse_attn_func) -----------
 * out [B,H,S,D] (any strides), lse fp32 [B,H,S] (natural log, may be NULL).
 * dst_row: optional device int32[S]: output row r is written to row dst_row[r]. */
int blade_block_sparse_attn_fwd(const BladeTensor* q, const BladeTensor* k, const BladeTensor* v,
                                const int32_t* idx, const int32_t* cnt, int64_t idx_stride,
                                BladeTensor* out, float* lse, const int32_t* dst_row,
                                float softmax_scale, void* workspace, size_t ws_bytes, void* stream);

/* ---- block_size 64 (BASELINE config 1): 64x64 mask granularity on the 128x128 tensor-core tiles -------------
 * blade_mask64_to_index: bool mask u8 [B,H,nq64,nk64] -> per 128-row query tile an ascending list of 128-key tiles,
 *   entry = tile id | (quadrant mask << 28) (bit 2*rowhalf+colhalf); idx int32 [B,H,ceil(nq64/2),ceil(nk64/2)].
 * blade_block_sparse_attn64_fwd / blade_asa_attn64_fwd: as the block-128 entry points, on such lists. */
int blade_mask64_to_index(const uint8_t* mask, int64_t B, int64_t H, int64_t nq64, int64_t nk64,
                          int32_t* idx, int32_t* cnt, void* stream);
int blade_block_sparse_attn64_fwd(const BladeTensor* q, const BladeTensor* k, const BladeTensor* v,
                                  const int32_t* idx, const int32_t* cnt, int64_t idx_stride,
                                  BladeTensor* out, float* lse, const int32_t* dst_row,
                                  float softmax_scale, void* workspace, size_t ws_bytes, void* stream);
int blade_asa_attn64_fwd(const BladeTensor* q, const BladeTensor* k, const BladeTensor* v,
                         const int32_t* idx, const int32_t* cnt, int64_t idx_stride,
                         const BladeTensor* k_pool, const BladeTensor* v_pool, int32_t sample_gap,
                         BladeTensor* out, const int32_t* dst_row,
                         float softmax_scale, int32_t exact_merge,
                         void* workspace, size_t ws_bytes, void* stream);

/* ---- a8+a10+a11: sparse branch + pooled branch + reference-exact LSE merge, one launch ------- */
int blade_asa_attn_fwd(const BladeTensor* q, const BladeTensor* k, const BladeTensor* v,
                       const int32_t* idx, const int32_t* cnt, int64_t idx_stride,
                       const BladeTensor* k_pool, const BladeTensor* v_pool, int32_t sample_gap,
                       BladeTensor* out, const int32_t* dst_row,
                       float softmax_scale, int32_t exact_merge,
                       void* workspace, size_t ws_bytes, void* stream);

/* ---- a1: the whole layer (AdaptiveBlockSparseAttnTrain.forward) ------------------------------
 * q,k,v in the caller's token order; src_row/dst_row as above (NULL = use_rearrange False).
 * scores_in: optional fp32 [B,H,nb,nb] to bypass the estimator (parity contract); scores_out / mask_out optional.
 * cfg->estimator 0 = block-mean-pool scores, 1 = the reference's sampled-max estimator (needs cfg->sample_*_off).
 * cfg->block_size 64 (BASELINE config 1): scores / mask_out / cnt_out are 64-granular ([.., nb64, nb64] / [.., nb64]);
 * idx_out, if given, receives the quadrant-flagged 128-tile lists [B,H,ceil(nb64/2),ceil(nb64/2)].
 * The running sparsity numerator goes to cfg->selected_acc (device counter, no host sync). */
int blade_asa_forward(const BladeTensor* q, const BladeTensor* k, const BladeTensor* v,
                      const int32_t* src_row, const int32_t* dst_row,
                      const BladeAsaConfig* cfg, const float* scores_in,
                      BladeTensor* out, float* scores_out, uint8_t* mask_out,
                      int32_t* idx_out, int32_t* cnt_out,
                      void* workspace, size_t ws_bytes, void* stream);

/* ---- f4: multi-level pooled sparse attention (the reference's experimental path) -------------------------------
 * Reference (under cogvideox/sample_evaluate/Triton/): cogvideo_newattn.py (N) and the Triton kernels
 * kernels/block_sparse_attn_kernel_with_backward_9_10.py (K9).  A block-mask entry is a LEVEL: 0 = skip, 1 = the
 * block's 128 keys, L in {2,4,8} = its 128/L mean-pooled keys/values with `+ log L` on the scaled score; one softmax

 * per query row (K9:135-277, 339-692).
 *
 *   reference interface                                           replaced by
 *   pooling x3 over the padded K, V       (K9:1252-1270, 1307-1316)   blade_multilevel_pyramid
 *   transfer_attn_to_mask(attn, ratios)   (N:154-207)                 blade_multilevel_mask
 *   sparse_attention_fn(q,k,v,mask,None)  (K9:1578-1611 -> _fwd_kernel K9:339-692)
 *                                                   blade_level_mask_to_index + blade_multilevel_attn_fwd
 *
 * blade_multilevel_pyramid: k,v [B,H,S,D] (any strides) -> contiguous k2,v2 [B,H,nb*64,D], k4,v4 [B,H,nb*32,D],
 *   k8,v8 [B,H,nb*16,D] (nb = ceil(S/128)); every round of pair means is rounded to the tensor dtype.
 * blade_multilevel_mask: scores fp32 [B,H,nq,nk] (nk <= 256); rank_level device u8 [nk]: level of the block ranked p-th
 *   (value descending, index ascending) -- the host expands the reference's ratio table into it; the last force_last
 *   rows and columns are set to level 1 (N:201-203).  Outputs: level_mask u8 [B,H,nq,nk] (optional), idx int32
 *   [B,H,nq,nk] = the row's block ids sorted by (level, block id), -1 padded; cnt4 int32 [B,H,nq,4] = entries of level
 *   1, 2, 4, 8 (16-byte aligned).
 * blade_multilevel_attn_fwd: every query row needs at least one non-zero entry (the reference forces the last two
 *   columns); zero-filled keys beyond the sequence in a level-1 tail block take part in the softmax with score 0 and
 *   value 0, exactly like the reference kernel's masked loads (K9:108-119). */
int blade_multilevel_pyramid(const BladeTensor* k, const BladeTensor* v, void* k2, void* v2, void* k4, void* v4,
                             void* k8, void* v8, void* stream);
int blade_multilevel_mask(const float* scores, int64_t B, int64_t H, int64_t nq, int64_t nk,
                          const uint8_t* rank_level, int32_t force_last, uint8_t* level_mask, int32_t* idx,
                          int32_t* cnt4, void* stream);
int blade_level_mask_to_index(const uint8_t* level_mask, int64_t B, int64_t H, int64_t nq, int64_t nk,
                              int32_t* idx, int32_t* cnt4, void* stream);
int blade_multilevel_attn_fwd(const BladeTensor* q, const BladeTensor* k, const BladeTensor* v,
                              const BladeTensor* k2, const BladeTensor* v2, const BladeTensor* k4,
                              const BladeTensor* v4, const BladeTensor* k8, const BladeTensor* v8,
                              const int32_t* idx, const int32_t* cnt4, int64_t idx_stride, BladeTensor* out,
                              float* lse, const int32_t* dst_row, float softmax_scale, void* workspace,
                              size_t ws_bytes, void* stream);

/* Backward of blade_multilevel_attn_fwd (replaces the Triton backward kernels K9:695-1237 / launch K9:1375-1576 behind
 * `sparse_attention_fn`): given the forward's inputs, its output `out`, its natural-log `lse` and the incoming gradient
 * `d_out` ([B,H,S,D], any strides), writes dq / dk / dv (tensor dtype, any strides).  The gradient through the pooled
 * K/V copies is folded back onto k and v (pair means; the pyramid's intermediate roundings are treated as identity,
 * like the reference's autograd).  Workspace: blade_multilevel_bwd_workspace_bytes (fp32 accumulators per level, zeroed
 * by the call; 1 KiB aligned). */
size_t blade_multilevel_bwd_workspace_bytes(int64_t B, int64_t H, int64_t S, int64_t Sk, int64_t D);
int blade_multilevel_attn_bwd(const BladeTensor* q, const BladeTensor* k, const BladeTensor* v,
                              const BladeTensor* k2, const BladeTensor* v2, const BladeTensor* k4,
                              const BladeTensor* v4, const BladeTensor* k8, const BladeTensor* v8,
                              const int32_t* idx, const int32_t* cnt4, int64_t idx_stride,
                              const BladeTensor* out, const BladeTensor* d_out, const float* lse,
                              float softmax_scale, BladeTensor* dq, BladeTensor* dk, BladeTensor* dv,
                              void* workspace, size_t ws_bytes, void* stream);

/* ---- benchmark scaffold (NOT the hot path) ---------------------------------------------------------------------
 * Fused token-wise glue of the random-init DiT scaffolds that the 8-step clip benchmark times (video_blade_b200/dit.py;
 * in the reference these ops are diffusers' transformer block, around `attn1`).  x, y, out: contiguous [B,S,C] tensors of
 * `dtype` (BLADE_BF16 / BLADE_F16); scale, shift, gate: fp32 [B,C]; C a multiple of 256, at most 4096.
 *   ln_modulate:     out = LayerNorm(x) [* weight + bias] * (1 + scale[b]) + shift[b]  (fp32 arithmetic, one rounding;
 *                    weight / bias [C] in `dtype`, NULL = no affine)
 *   rmsnorm:         out = x * rsqrt(mean(x^2) + eps) * weight                          (weight [C] in `dtype`)
 *   gated_residual:  out = x + y * gate[b] */
int blade_scaffold_ln_modulate(const void* x, const float* scale, const float* shift, const void* weight,
                               const void* bias, void* out, int64_t B, int64_t S, int64_t C, float eps, int32_t dtype,
                               void* stream);
int blade_scaffold_rmsnorm(const void* x, const void* weight, void* out, int64_t rows, int64_t C, float eps,
                           int32_t dtype, void* stream);
int blade_scaffold_gated_residual(const void* x, const void* y, const float* gate, void* out, int64_t B, int64_t S,
                                  int64_t C, int32_t dtype, void* stream);

/* ---- measurement hook ------------------------------------------------------------------------
 * When set, the library records the given cudaEvent_t pair (passed as void*) on the caller's stream
 * immediately before / after the launches of one stage, so a harness can time a kernel INSIDE a whole-layer
 * call.  stage: 0 = prep (gather + block means), 1 = scores, 2 = select, 3 = attention, 4 = gap pooling (runs on the
 * library's side stream, concurrently with 1 and 2).  NULL,NULL clears the slot.  Per calling thread. */
int blade_profile_events(int32_t stage, void* start_event, void* stop_event);

/* ---- bring-up probes (tests only): single-tile tcgen05 GEMMs dumped from TMEM --------------- */
int blade_probe_qk(const void* q_tile /*bf16 [128,D]*/, const void* k_tile /*bf16 [128,D]*/,
                   float* s_out /*[128,128]*/, int32_t D, void* stream);
int blade_probe_pv(const float* p_tile /*[128,128]*/, const void* v_tile /*bf16 [128,D]*/,
                   float* o_out /*[128,D]*/, int32_t D, void* stream);

/* ---- schedule replay (tests only, HOST code, no GPU needed) -------------------------------------
 * The persistent attention kernel decodes its work items (tile pairs, solo tiles, half tiles) with integer code
 * shared by host and device.  This entry point replays the decode of one launch on host block counts `cnt_host`
 * (int32 [B,H,nq]) for a device with `sm_count` SMs and writes 12 ints per item to items_out:
 *   item, bh, {qb, pooled_tiles, list_offset, list_entries} x 2 streams, merge, split (0 | 1 + 2*slot + half)
 * (qb == nq: stream without a tile).  dynamic_queue / half_tiles select the scheduling variants of
 * blade_asa_attn_fwd.  Returns BLADE_ERR_WORKSPACE (and *n_items_out) when max_items is too small. */
int blade_debug_attn_schedule(int64_t B, int64_t H, int64_t nq, int32_t n_pool_tiles, int32_t sm_count,
                              const int32_t* cnt_host, int32_t dynamic_queue, int32_t half_tiles,
                              int32_t* items_out, int64_t max_items, int32_t* n_items_out);

#ifdef __cplusplus
}
#endif
#endif /* BLADE_ASA_H_ */

/* mca_b200.h — C ABI of libmca_b200.so, the B200 (sm_100a) implementation of the mca-paper training hot path.
 *
 * The reference (josiahbjorgaard/mca-paper) has no FFI layer: its boundary is the Python nn.Module API of
 * model.py / encoders.py / utils/contrastive_loss_with_temperature.py.  This header is what a
 * Python (ctypes) or C++ host binds instead of the ATen calls those modules make; every entry point cites the
 * reference lines it replaces.  Conventions:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in `_host`
 *   - the library never allocates or frees device memory and never synchronises; work is enqueued on `stream`
 *     (a cudaStream_t passed as void*), so every call is CUDA-graph capturable
 *   - return value: MCA_OK or one of the MCA_ERR_* codes below
 *   - bf16 tensors are raw 16-bit storage (`void*`), fp32 tensors are `float*`
 *   - token matrices are row-major [B*N, cols]: row = sample*N + position (the packed layout of model.py:464)
 */
#ifndef MCA_B200_H
#define MCA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCA_MAX_MODALITIES 32 /* modality blocks of the packed sequence (EAO replicates modalities per pass) */

enum {
  MCA_OK = 0,
  MCA_ERR_SHAPE = 1,     /* bad dimension / divisibility (reference: AssertionError, model.py:186,416,437) */
  MCA_ERR_ALIGN = 2,     /* pointer or stride not 16-byte aligned */
  MCA_ERR_CUDA = 3,      /* launch / driver failure */
  MCA_ERR_NONFINITE = 4, /* reference: Exception on non-finite tokens, encoders.py:197-213 (device flag) */
  MCA_ERR_ARG = 5
};

/* epilogue selectors of mca_gemm_bf16 */
enum {
  MCA_EPI_BF16 = 0,      /* out0(bf16) = alpha*acc + bias */
  MCA_EPI_F32 = 1,       /* out0(f32)[z] = alpha*acc + bias, one slab per k-split z */
  MCA_EPI_RESID = 2,     /* out0(f32) = alpha*acc + aux0(f32); optional out1(bf16) copy */
  MCA_EPI_GEGLU = 3,     /* out1(bf16) = u = acc (interleaved value|gate), out0(bf16) = gelu(gate)*value */
  MCA_EPI_GEGLU_BWD = 4, /* acc = dL/dh, aux0 = u; out0(bf16) = dL/du */
  MCA_EPI_F32_ACC = 5    /* out0(f32) += alpha*acc: every k-split reduce-adds (TMA) into ONE slab the caller zeroed */
};

/* ---- static attention schedule (host-built once from token_types / attn_mask, model.py:383-430) ---- */
typedef struct { int start, len; } mca_attn_tile;                 /* a run of <=128 positions inside a sample */
typedef struct { int tile, flags; } mca_attn_ref;                 /* flags bit0: tile holds disallowed pairs   */
typedef struct { int start, len, kt_off, kt_cnt; } mca_attn_qtile; /* query tile + its slice of the ref list   */

/* ---- weight pack / gradient unpack descriptor (state_dict layout <-> kernel layout) ---- */
typedef struct {
  long long src_off;      /* element offset in the flat fp32 parameter (or gradient) buffer */
  long long dst_off;      /* element offset in the bf16 operand arena (or fp32 partial-gradient arena) */
  int rows, cols;         /* source matrix shape */
  int dst_ld;             /* row stride of the kernel-layout matrix */
  int dst_row0;           /* first destination row */
  int mode;               /* 0 identity rows, 1 GEGLU interleave: 64 value rows then 64 gate rows per 128 block */
  int half;               /* rows in the value half (mode 1) */
  float scale;            /* pack: multiplies weights; unpack: multiplies gradients */
  int n_splits;           /* unpack: number of split-K slabs to sum */
  long long split_stride; /* unpack: elements between slabs */
} mca_pack_desc;

/* ---- one contrastive pair of MCAPretrainingLoss (model.py:160-168,198-220) ---- */
typedef struct {
  int a_row, b_row;   /* rows of the pooled block [B,R,d] */
  uint32_t all_mask;  /* modalities that must ALL be present for a sample to count */
  uint32_t any_mask;  /* modalities of which at least one must be present (0 = no constraint) */
  int is_fcl;         /* counted in 'fcl_loss' (1) or 'no-fcl_loss' (0), model.py:221-222 */
} mca_loss_pair;

typedef struct {
  float lr, beta1, beta2, eps, weight_decay, max_norm; /* max_norm <= 0: no clipping */
  int lr_mode;                                         /* 0 constant, 1 cosine, 2 constant_with_warmup, 3 linear (1-3: linear warm-up) */
  long long warmup_steps, total_steps;
  long long sched_stride;                              /* scheduler.step() calls per optimiser step: accelerate's
                                                          prepared scheduler advances num_processes times per step
                                                          (train_accel_gpu.py:93,119); 0 or 1 = once */
} mca_adamw_cfg;

int mca_version(void);

/* Dense contraction out[M,N] = A[M,K] * B[N,K]^T on tcgen05 tensor cores (bf16 in, fp32 accumulate in TMEM).
 * Replaces every nn.Linear matmul of the step and its autograd transposes: model.py:83 (to_q,to_kv), :105 (to_out),
 * :49-51 (feed-forward, with model.py:37-38 GEGLU fused as an epilogue), encoders.py:190 (token projection).
 * a_mn_major/b_mn_major = 0: operand stored [rows, K] with K contiguous (ld = row stride in elements);
 *                       = 1: operand stored [K, rows] with rows contiguous (the autograd transposes need no copies).
 * k_splits > 1 (MCA_EPI_F32 only) splits K; slab z of out0 holds the partial sum of split z; the effective split
 * count is mca_gemm_effective_splits(K, k_splits). N must be a multiple of 32. */
int mca_gemm_bf16(const void* A, int a_mn_major, long long lda, const void* B, int b_mn_major, long long ldb, int M,
                  int N, int K, int k_splits, int mode, void* out0, long long ld0, void* out1, long long ld1,
                  const void* aux0, long long ldaux, const float* bias, float alpha, void* stream);
int mca_gemm_effective_splits(int K, int k_splits);

/* Modality-dropout / pad-mask builder (model.py:455-466; collator masks encoders.py:307,339).  masks_host[m] is the
 * DEVICE pointer of modality m's attention_mask [B, len m] (elem size 1 = bool, 8 = int64; non-zero = padded).
 * Outputs: padding [B,N] bytes (== reference `padding`), pad_mod (per-modality [B,len] bytes, modality-major),
 * present [B,n_mod] (== modality_sample_mask, model.py:458), live_count [B,n_mod], live_idx [B,N] (packed varlen
 * gather indices: live positions of each modality in order, -1 filled), cu_live [B*n_mod+1] (exclusive cumsum),
 * kt_class [B,n_kt] (0 all keys live, 1 mixed, 2 all padded -> tile skipped), kt_live [B,n_kt,4] (bit j of the
 * 128-bit word = key j of the tile exists and is not padded), any_absent (1 int). */
int mca_build_offsets(const void* const* masks_host, const int* elem_sizes_host, const int* lens_host, int n_mod,
                      int B, int N, const int* kt_start, const int* kt_len, int n_kt, uint8_t* padding,
                      uint8_t* pad_mod, uint8_t* present, int* live_count, int* live_idx, int* cu_live,
                      uint8_t* kt_class, uint32_t* kt_live, int* any_absent, void* stream);

/* LayerNorm over d = 512 (model.py:24-31; eps 1e-5).  Optional: pad [rows] bytes -> zero output (encoders.py:205),
 * pe [seg_len,512] added after (encoders.py:208-209), row scatter out_row = (r/seg_len)*out_rows_per_b + out_row_off
 * + r%seg_len (writes straight into the packed token buffer).  stats = (mean, rstd) per row. */
int mca_layernorm512_fwd(const float* x, const float* gamma, const float* beta, float* y32, void* y16, float* stats,
                         const uint8_t* pad, const float* pe, int seg_len, int out_rows_per_b, int out_row_off,
                         long long rows, void* stream);
/* dy_delta_bf16 (optional, same row mapping as dy): gradient of the branch GEMM, added to dy before the backward */
int mca_layernorm512_bwd(const float* dy, const void* dy_delta_bf16, const float* x, const float* stats,
                         const float* gamma, float* dx32, void* dx16, float* dgamma, float* dbeta, const uint8_t* pad,
                         int seg_len, int out_rows_per_b, int out_row_off, long long rows, void* stream);
/* Residual add fused with the next LayerNorm (model.py:118-122; quirk Q1: the residual is the NORMED tensor):
 * xnew = LN(xprev; stats_prev, gamma_prev, beta_prev) + y  (fp32, optional), out = LN(xnew; gamma, beta) as the bf16
 * operand of the next GEMM, stats = its (mean, rstd).  The normed residual is recomputed, never stored in fp32. */
int mca_add_layernorm512_fwd(const float* xprev, const float* stats_prev, const float* gamma_prev,
                             const float* beta_prev, const void* y_bf16, float* xnew, const float* gamma,
                             const float* beta, void* out_bf16, float* stats, long long rows, void* stream);
/* Input LayerNorm of EmbeddedSequenceEncoder (encoders.py:189,199): y = bf16 GEMM operand zero-padded to kpad cols;
 * sets *nonfinite_flag when any token is not finite (encoders.py:197-198). */
int mca_layernorm_in_fwd(const float* x, const float* w, const float* b, const uint8_t* pad, void* y_bf16,
                         float* stats, int width, int kpad, long long rows, int* nonfinite_flag, void* stream);
int mca_layernorm_in_param_bwd(const float* dy, int ld_dy, const float* x, const float* stats, const uint8_t* pad,
                               float* dw, float* db, int width, long long rows, void* stream);
int mca_colsum(const float* a, int ld, float* out, int width, long long rows, void* stream);

/* TabularEncoder pieces (encoders.py:25-37,55-72,90-96): in-place max_norm renormalisation of the embedding table
 * (nn.Embedding(max_norm=1.0) semantics), h1 = relu(w1*min(v,max_value)+b1) as the bf16 operand of the d x d Linear
 * (vpad[r] = v[r]==padding_value), and the parameter gradients of that first Linear. */
int mca_embedding_renorm(float* emb, int rows, int d, float max_norm, void* stream);
int mca_tabular_fwd(const float* values, const float* w1, const float* b1, void* h1_bf16, uint8_t* vpad,
                    float max_value, float padding_value, int d, long long rows, void* stream);
int mca_tabular_bwd(const float* dh1, const float* values, const float* w1, const float* b1, float* dw1, float* db1,
                    const void* unused0, const void* unused1, float max_value, float padding_value, int d,
                    long long rows, void* stream);

/* fp32 state_dict-layout weights -> bf16 kernel-layout operands, and the inverse for gradients. */
int mca_pack_weights(const float* params, void* arena_bf16, const mca_pack_desc* descs_dev, int n_desc, void* stream);
int mca_unpack_grads(float* grads, const float* partials, const mca_pack_desc* descs_dev, int n_desc, void* stream);

/* Index-driven embedding tables: SequenceEncoder (encoders.py:145-166) and SparseTabularEncoder (:100-120) look rows up
 * by int64 indices [B, L].  nn.Embedding(max_norm) renormalises the LOOKED-UP rows in place, once each
 * (flags_scratch: uint8[rows], zero on entry and exit); an out-of-range index sets bit 1 of *bad_index_flag.
 * gather: dst[b*dst_rows_per_b + dst_row_off + l] (+)= emb[idx[b,l]] (+ pe[l]); scatter_add: the gradient of the
 * gather into demb, skipping row skip_row (padding_idx gets no gradient). */
int mca_embedding_renorm_indexed(float* emb, const long long* idx, long long n_idx, int rows, int d, float max_norm,
                                 uint8_t* flags_scratch, int* bad_index_flag, void* stream);
int mca_embedding_gather(const float* emb, const long long* idx, int rows_emb, int B, int L, int d, const float* pe,
                         float* dst, int dst_rows_per_b, int dst_row_off, int accumulate, void* stream);
int mca_embedding_scatter_add(const float* dsrc, const long long* idx, int rows_emb, int B, int L, int d,
                              int src_rows_per_b, int src_row_off, int skip_row, float* demb, void* stream);
/* PatchEncoder "matrix" mode (encoders.py:217-274): values [B,H,W] -> tokens [B*(H/p1)*(W/p2), p1*p2] in
 * 'b (h p1) (w p2) -> b (h w) (p1 p2)' order, mask[t] = all(patch == pad_token). */
int mca_patchify(const float* values, int B, int H, int W, int p1, int p2, float pad_token, float* tokens, uint8_t* mask,
                 void* stream);
/* nn.Dropout(p) on the token rows [b*rows_per_b + row_off + l, :d]: counter-based mask from (seed, *counter_dev,
 * element), so calling it again with the same counter on the gradient applies the same mask (no stored mask). */
int mca_dropout_rows(float* x, int B, int L, int d, int rows_per_b, int row_off, float p, unsigned long long seed,
                     const long long* counter_dev, void* stream);

/* Device-side collate, the producer of the hot path's batch (MultimodalCollator encoders.py:374-403 over
 * EmbeddedSequenceCollator :314-343, SequenceCollator :286-311, MatrixCollator :346-364).  The host stages the live rows
 * of the present samples back to back (row_off[B+1], absent modality = empty range) and these kernels expand them:
 * rows: out[b,l,:] = l < len_b ? src[row_off[b]+l,:] : fill (clean != 0: torch.nan_to_num), rows beyond L truncated,
 *       mask[b,l] = (l >= len_b) as bytes (nullptr: no mask);
 * values: out[b,l] = l < len_b ? src[off[b]+l] : pad_token, mask[b,l] = (out[b,l] == pad_token) as int64. */
int mca_collate_rows(const float* src, const int* row_off, int B, int L, int E, float fill, int clean, float* out,
                     uint8_t* mask, void* stream);
int mca_collate_values_f32(const float* src, const int* off, int B, int L, float pad_token, float* out, long long* mask,
                           void* stream);
int mca_collate_values_i64(const long long* src, const int* off, int B, int L, long long pad_token, long long* out,
                           long long* mask, void* stream);

/* fusion tokens: broadcast into the packed buffer (model.py:460-461) / batch-sum of their gradient */
int mca_broadcast_rows(const float* src, float* dst, int F, int d, int B, int rows_per_b, int row_off, void* stream);
int mca_batchsum_rows(const float* src, float* out, int F, int d, int B, int rows_per_b, int row_off, int accumulate,
                      void* stream);
int mca_cast_f32_bf16(const float* src, long long ld_src, void* dst, long long ld_dst, long long rows, int cols,
                      void* stream);

/* Block-sparse masked multi-head attention (model.py:85-100), dim_head = 64.  qkv: bf16 [B*N, 3*H*64] = (Q*scale | K | V),
 * out: bf16 [B*N, H*64], lse: [B,H,N] natural-log row log-sum-exp (+inf marks a fully masked row).
 * q_tiles may be listed in any order (heaviest first balances the SMs); tile_grp[n_kt] = key group shared by all
 * keys of the tile or 255 when the tile mixes groups; kt_live = live-key bit words from mca_build_offsets.
 * Varlen query skipping (north_star subsystem 1: tokens of absent / padded modalities are never read): skip_ok [B] bytes
 * from mca_query_skip_flags (NULL = off); for a sample whose flag is set, a query tile whose rows are ALL padded
 * (padding [B,N]) visits no key tile: its rows get the fully-masked value and LSE = +inf, the backward leaves it out. */
int mca_attn_fwd(const void* qkv, const mca_attn_qtile* q_tiles, int n_qt, const mca_attn_ref* kt_list,
                 const mca_attn_tile* k_tiles, int n_kt, const uint32_t* rowbits, const uint8_t* keygrp,
                 const uint8_t* tile_grp, const uint8_t* kt_class, const uint32_t* kt_live, const uint8_t* padding,
                 const uint8_t* skip_ok, const int* any_absent, float* vmean, void* out, float* lse, int B, int N, int H,
                 void* stream);
/* Backward: dout bf16 [B*N, H*64] -> dqkv bf16 [B*N, 3*H*64].  k_tiles_q lists, for every key tile, the query tiles
 * that attend it (transposed schedule).  dq_accum: fp32 [B*N, H*64] scratch, delta: [B,H,N], ucorr: [B, H*64]. */
int mca_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, const mca_attn_qtile* k_tiles_q,
                 int n_kt, const mca_attn_ref* qt_list, const mca_attn_tile* q_tiles, int n_qt, const uint32_t* rowbits,
                 const uint8_t* keygrp, const uint8_t* tile_grp, const uint8_t* padding, const uint8_t* kt_class,
                 const uint8_t* skip_ok, float* delta, float* ucorr, float* dq_accum, void* dqkv, int B, int N, int H,
                 void* stream);
/* skip_ok[b] for the varlen query skipping above, from `present` [B, n_blk] of mca_build_offsets (first n_mod columns =
 * modalities).  mode 0: never; 1 (exact): only samples in which EVERY modality is present — padded rows of such a sample
 * feed nothing (every consumer masks padded keys), so results are unchanged; with an absent modality the reference's
 * fully-masked pooling row averages ALL N final tokens, padded ones included (quirk Q4/Q8), and they must be computed;
 * 2 (fast): every sample — padded query rows are never scored; the pooled embedding of an ABSENT modality (a negative
 * column of the loss) then averages substitute values for the padded rows: a documented approximation. */
int mca_query_skip_flags(const uint8_t* present, int n_blk, int n_mod, int B, int mode, uint8_t* skip_ok, void* stream);

/* Attention pooling core (model.py:472-473): qp [R,H*64] fp32 scaled queries, kv bf16 [B*N, 2*H*64] (K|V),
 * rowbits[R] allowed key groups per pooled row, probs [B,H,R,N] saved for the backward, out [B,R,H*64]. */
int mca_pool_attn_fwd(const float* qp, const void* kv, const uint8_t* padding, const uint8_t* keygrp,
                      const uint32_t* rowbits, float* probs, uint8_t* full_masked, float* out, int B, int H, int R,
                      int N, void* stream);
int mca_pool_attn_bwd(const float* dout, const float* qp, const void* kv, const float* probs,
                      const uint8_t* full_masked, float* ds_scratch, void* dkv, float* dqp, int B, int H, int R, int N,
                      void* stream);
/* small fp32 GEMM with arbitrary strides: C[m,n] = alpha*sum_k A(m,k)B(n,k) (+C) (+add[m % add_rows, n]; add_rows = 0:
 * add[m, n]) — the return-token projections around attention pooling (model.py:472-473, R*B <= 128 rows) */
int mca_small_gemm_f32(const float* A, long long sam, long long sak, const float* Bm, long long sbn, long long sbk,
                       float* C, long long ldc, const float* add, long long ldadd, int add_rows, int M, int N, int K,
                       float alpha, int accumulate, void* stream);

/* All-pairs temperature-scaled InfoNCE (model.py:196-232 + utils/contrastive_loss_with_temperature.py:71-100,187).
 * pooled_all: [GB, R, d] all-gathered pooled tokens, local rows at rank*B; losses[n_pairs] (NaN = no selected row);
 * summary = {loss, fcl_loss, no-fcl_loss, #non-NaN}; w_default[p] = d loss / d losses[p]. */
int mca_contrastive_allpairs_fwd(const float* pooled_all, const uint8_t* present, const mca_loss_pair* plan_dev,
                                 int n_pairs, float* logit_scale, int B, int GB, int R, int d, int n_mod, int rank,
                                 float scale_min, float scale_max, float* losses, float* summary, float* w_default,
                                 void* stream);
/* dpooled_all [GB,R,d] and dscale must be zeroed by the caller; w[p] = upstream gradient of losses[p]. */
int mca_contrastive_allpairs_bwd(const float* pooled_all, const uint8_t* present, const mca_loss_pair* plan_dev,
                                 int n_pairs, float* logit_scale, int B, int GB, int R, int d, int n_mod, int rank,
                                 const float* w, float* dpooled_all, float* dscale, void* stream);

/* Free-standing calls of the reference surface (not on the fused MCA.forward path, which never materialises logits or
 * attention probabilities).  Functional contrastive_loss_with_temperature(...) -> ContrastiveLossOutput
 * (utils/contrastive_loss_with_temperature.py:40-108): logits[i,j] = exp(*logit_scale) * <a_i, b_all_j> (:71,82-88);
 * F.cross_entropy rows with label_smoothing (cross_entropy_kwargs, :94-100): row_loss / row_lse [rows]; backward:
 * dlogits[rows,cols] = g_row[i] * dCE/dz * exp(*logit_scale) (logit_scale may be NULL: factor 1),
 * *dscale += sum dCE/dz * z (caller zeroes it; may be NULL). */
int mca_scaled_logits_f32(const float* a, const float* b_all, const float* logit_scale, int n_a, int n_b, int d,
                          float* logits, void* stream);
int mca_cross_entropy_fwd(const float* logits, long long ld, const long long* labels, int rows, int cols,
                          float label_smoothing, float* row_loss, float* row_lse, void* stream);
int mca_cross_entropy_bwd(const float* logits, long long ld, const long long* labels, int rows, int cols,
                          float label_smoothing, const float* row_lse, const float* g_row, const float* logit_scale,
                          float* dlogits, float* dscale, void* stream);
/* Attention(..., return_attn=True) (model.py:96,102-103): probs [B,H,N,N] fp32 recomputed from the saved lse of
 * mca_attn_fwd; fully masked rows (lse = +inf) are uniform over all N keys (quirk Q4). */
int mca_attn_probs(const void* qkv, const float* lse, const uint32_t* rowbits, const uint8_t* keygrp,
                   const uint8_t* padding, int B, int N, int H, float* probs, void* stream);

/* Peer-memory exchange of the pooled block for data parallelism inside one NVLink/NVSwitch domain (replaces the
 * all_gather / reduce_scatter of utils/distributed.py:23-56 and torch.distributed.nn.functional.all_gather's backward).
 * The gathered [GB,R,d] buffers live in P2P-mapped symmetric memory, one per rank, addressable by all ranks:
 *   forward : mca_p2p_push_rows(own block -> slot `rank` of every rank's gathered buffer), mca_xgpu_barrier,
 *             mca_contrastive_allpairs_fwd on the local gathered buffer;
 *   backward: mca_contrastive_allpairs_bwd into the local gradient buffer, mca_xgpu_barrier,
 *             mca_p2p_reduce_rows (sum of every rank's slice of this rank's rows). */
/* dst_peers_dev[g][off_elems + i] = src[i], i < n, for every rank g (posted NVLink stores). */
int mca_p2p_push_rows(const float* src, float* const* dst_peers_dev, long long off_elems, long long n, int world,
                      void* stream);
/* dst[i] = sum_g src_peers_dev[g][off_elems + i], i < n: the pull form of a reduce-scatter over peer memory. */
int mca_p2p_reduce_rows(const float* const* src_peers_dev, long long off_elems, float* dst, long long n, int world,
                        void* stream);
/* Flag barrier across the GPUs of one node: flags_peers_dev[g] = rank g's uint32[world] flag array (peer-mapped, zeroed
 * once), epoch_dev = this rank's barrier counter (device, zeroed once).  err_flag_dev is set if a peer does not arrive
 * within ~10 s (the kernel then returns instead of hanging the GPU).  Optional payload: *payload (one double of this
 * rank) is stored into payload_peers_dev[g][rank] of every rank before the flag is raised. */
int mca_xgpu_barrier(uint32_t* const* flags_peers_dev, int world, int rank, uint32_t* epoch_dev, int* err_flag_dev,
                     const double* payload, double* const* payload_peers_dev, void* stream);

/* clip_grad_norm_(max_norm) + AdamW + LR schedule on flat buffers (train_accel_gpu.py:80-86,116-119).
 * step_dev: device int64 step counter (incremented here); grads are multiplied by grad_scale first (1/world). */
int mca_clip_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                        double* sumsq_scratch, long long* step_dev, float* total_norm_out, float grad_scale,
                        const mca_adamw_cfg* cfg_host, void* stream);

/* Data-parallel optimiser over peer memory (replaces DDP's gradient all-reduce, train_accel_gpu.py:93,115, followed by
 * clip_grad_norm_ + AdamW on every rank): rank r owns flat elements [shard_off, shard_off + shard_n).
 * mca_dp_reduce_shard: grads_local[shard] = sum_g grads_peers_dev[g][shard] (pulled over NVLink), *sumsq_local = its sum
 * of squares.  After a barrier that carries the G partial sums (mca_xgpu_barrier payload), mca_dp_adamw_shard clips by the
 * global norm, applies AdamW to the shard (moments of the shard only) and stores the new parameters into
 * params_peers_dev[g][shard] of EVERY rank (all-gather by posted stores); a last barrier publishes them. */
int mca_dp_reduce_shard(const float* const* grads_peers_dev, float* grads_local, long long shard_off, long long shard_n,
                        int world, double* sumsq_local, void* stream);
int mca_dp_adamw_shard(float* const* params_peers_dev, int world, int rank, const float* grads_local, float* exp_avg,
                       float* exp_avg_sq, long long shard_off, long long shard_n, const double* sumsq_slots,
                       long long* step_dev, float* total_norm_out, float grad_scale, const mca_adamw_cfg* cfg_host,
                       void* stream);

/* The same two kernels over the NVSwitch MULTICAST mapping of the symmetric flat buffers (NVLS): the gradient shard is read
 * with multimem.ld_reduce (the switch returns the sum over all ranks: 1/G of the pull traffic) and the updated parameters are
 * written with multimem.st (one store, replicated by the switch into every rank's buffer).  *_multicast = the multicast
 * address of the flat gradient / parameter buffer (torch symmetric memory: handle.multicast_ptr + offset). */
int mca_dp_reduce_shard_mc(const float* grads_multicast, float* grads_local, long long shard_off, long long shard_n, int world,
                           double* sumsq_local, void* stream);
int mca_dp_adamw_shard_mc(float* const* params_peers_dev, float* params_multicast, int world, int rank, const float* grads_local,
                          float* exp_avg, float* exp_avg_sq, long long shard_off, long long shard_n, const double* sumsq_slots,
                          long long* step_dev, float* total_norm_out, float grad_scale, const mca_adamw_cfg* cfg_host,
                          void* stream);

/* ---- Mean pooling of the EAO baseline: MeanTokenProjectionPool(token_types=None, projection=False), model.py:235-280 as
 * EAO.single_pass uses it (model.py:553-556,562-563).  x: bf16 [B, N, 512] final-normed tokens of all passes of a sample
 * laid back to back; padding [B, N] bytes (non-zero = padded key); pass_start [R + 1] token offsets; pooled [B, R, 512] =
 * mean of the live tokens of each pass (zeros when there are none, model.py:270-271); cnt [B, R] live counts (kept for
 * the backward); scratch: mca_mean_pool_scratch_floats(B, R) floats.  Backward: dx [B, N, 512] fp32 =
 * dpooled[b, tok_pass[t]] / cnt for live tokens, 0 for padded ones (tok_pass [N] = pass of every token). */
int mca_mean_pool_scratch_floats(int B, int R);
int mca_mean_pool_fwd(const void* x_bf16, const uint8_t* padding, const int* pass_start, int B, int N, int R, int d,
                      float* pooled, float* cnt, float* scratch, void* stream);
int mca_mean_pool_bwd(const float* dpooled, const uint8_t* padding, const int* tok_pass, const float* cnt, int B, int N,
                      int R, int d, float* dx, void* stream);

/* ---- Embedding-space evaluation metrics (utils/metrics.py; eval loop train_accel_gpu.py:136-184, infer_accel_gpu.py:115-147,
 * lp_accel_gpu.py:70-95).  All inputs fp32 row-major [rows, D] on the device; `scratch` holds at least
 * mca_metric_scratch_doubles(M) doubles; results are single floats / int64 ranks on the device (no host sync). */
long long mca_metric_scratch_doubles(long long M);
/* inv[i] = 1 / max(||x_i||_2, eps)  (F.normalize: eps 1e-12, utils/metrics.py:21-22,27; nn.CosineSimilarity: eps 1e-8, :75-76) */
int mca_row_inv_norms(const float* x, long long M, int D, float eps, float* inv, void* stream);
/* lalign, utils/metrics.py:20-23: out = mean_i ||x_i - y_i||_2 ^ alpha, rows L2-normalised first when norm != 0
 * (M == 0 -> NaN like the mean of an empty tensor). */
int mca_alignment(const float* x, const float* y, long long M, int D, float alpha, int norm, double* scratch, float* out,
                  void* stream);
/* lunif, utils/metrics.py:26-29: out = log(mean_{i<j} exp(-t * ||x_i - x_j||^2))  (torch.pdist pairs; M < 2 -> NaN).
 * inv_scratch: M floats. */
int mca_uniformity(const float* x, long long M, int D, float t, int norm, float* inv_scratch, double* scratch, float* out,
                   void* stream);
/* get_rank_metrics / get_rank, utils/metrics.py:73-92: ranks[i] = #{ j : cos(emb_i, targets_j) > cos(emb_i, targets_idx[i]) }
 * with idx[i] the sample's own row in `targets` (the reference passes the sample's index in the unmasked array).
 * inv_e [M], inv_t [T], own [M] are float scratch; idx outside [0, T) is the caller's error (the reference raises
 * IndexError; the host wrapper checks). */
int mca_retrieval_ranks(const float* emb, const float* targets, const long long* idx, long long M, long long T, int D,
                        float* inv_e, float* inv_t, float* own, long long* ranks, void* stream);

/* ---- Linear probe on frozen embeddings (lp_accel_gpu.py:22-35 FineTuneDataset, :100-104 nn.Linear(num_emb, num_labels),
 * :118-131 losses, :160-231 epoch loop with clip_grad_norm_ / AdamW / get_scheduler stepped once per batch).  One launch =
 * one EPOCH: a thread-block cluster keeps the parameters and AdamW moments in shared memory and exchanges the per-CTA
 * gradient partials through distributed shared memory.  X [rows, 512], Y [rows, n_out] fp32; order [n] = dataset indices in
 * visiting order (the RandomSampler permutation; arange for evaluation); state = [3][n_out*512 + n_out] (W row-major then
 * bias; exp_avg; exp_avg_sq); loss_kind 0 L1Loss, 1 MSELoss, 2 BCEWithLogitsLoss, 3 CrossEntropyLoss with probability
 * targets (all mean-reduced); train = 0 evaluates (no update).  pred [rows, n_out] gets every visited row's prediction,
 * *loss_sum += sum over the batches of the batch-mean loss (the reference's epoch_loss).  n_out <= 8. */
int mca_probe_epoch(const float* X, const float* Y, const int* order, int n, int batch_size, int n_out, int loss_kind,
                    int train, float* state, long long* step_dev, const mca_adamw_cfg* cfg_host, float* pred,
                    double* loss_sum, float* last_grad_norm, void* stream);
/* torchmetrics.PearsonCorrCoef of (pred, y) over n values (lp_accel_gpu.py:148-149). */
int mca_probe_pcc(const float* pred, const float* y, long long n, float* out, void* stream);

/* ---- fp32-parity forward mode (csrc/exact.cu).  The reference computes in fp32 end to end (train_accel_gpu.py:21 default
 * Accelerator(), no autocast; model.py:73-105), and north_star asks for loss / embeddings within 1e-3 of it.  Every dense
 * contraction of the forward then runs as a 3-term bf16 split product on the SAME tensor-core GEMM: with x = hi + lo
 * (hi = bf16(x), lo = bf16(x - hi)),  A W^T ~= A_hi W_hi^T + A_hi W_lo^T + A_lo W_hi^T  is one mca_gemm_bf16 over K' = 3K
 * with operands [A_hi | A_hi | A_lo] and [W_hi | W_lo | W_hi].  These entry points produce those operands and the fp32
 * pieces between the GEMMs; the attention core runs in fp32 on the CUDA cores (online softmax) and leaves the bf16 copies
 * and the log-sum-exp the regular backward consumes.  Selected with Engine.set_precision("fp32") / MCA_PRECISION=fp32. */
/* dst3 bf16 [rows, 3*kpad] <- src fp32 [rows, cols] (zero padded to kpad).  weight_layout 0: [hi|hi|lo], 1: [hi|lo|hi]. */
int mca_x_split_f32(const float* src, long long ld_src, void* dst3, long long rows, int cols, int kpad, int weight_layout,
                    void* stream);
/* mca_pack_weights with every kernel-layout matrix written as [hi|lo|hi] (row stride 3*dst_ld, arena offsets x 3). */
int mca_x_pack_weights_split(const float* params, void* arena3_bf16, const mca_pack_desc* descs_dev, int n_desc, void* stream);
/* encoders.py:187-190 first LayerNorm (as mca_layernorm_in_fwd), output as the [hi|hi|lo] operand [rows, 3*kpad]. */
int mca_x_layernorm_in_split(const float* x, const float* w, const float* b, const uint8_t* pad, void* y3, int width, int kpad,
                             long long rows, void* stream);
/* encoders.py:60-75 first stage (as mca_tabular_fwd), output as the [hi|hi|lo] operand [rows, 3*d]. */
int mca_x_tabular_split(const float* values, const float* w1, const float* b1, void* h3, float max_value, int d,
                        long long rows, void* stream);
/* model.py:35-38 on the fp32 FF1 output u32 [M, 2*IP] ([64 value | 64 gate] per 128 columns): h3 = [hi|hi|lo] of
 * h = x * gelu(g) [M, 3*IP]; h16 [M, IP] and u16 [M, 2*IP] = what MCA_EPI_GEGLU leaves for the backward. */
int mca_x_geglu_f32(const float* u32, void* h3, void* h16, void* u16, long long M, int IP, void* stream);
/* model.py:85-100 in fp32: qkv32 [B*N, 3*H*64] (q pre-scaled) -> out32 / out16 [B*N, H*64], lse [B,H,N]; allowed(q,k) =
 * rowbits[q] >> keygrp[k] & 1 and padding[b,k] == 0; a row with no allowed live key gets vmean (mean of V over all N, Q4). */
int mca_x_attn_fwd_f32(const float* qkv32, const uint32_t* rowbits, const uint8_t* keygrp, const uint8_t* padding, float* vmean,
                       float* out32, void* out16, float* lse, int B, int N, int H, void* stream);
/* mca_pool_attn_fwd on fp32 K | V rows. */
int mca_x_pool_attn_fwd_f32(const float* qp, const float* kv32, const uint8_t* padding, const uint8_t* keygrp,
                            const uint32_t* rowbits, float* probs, uint8_t* full_masked, float* out, int B, int H, int R, int N,
                            void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MCA_B200_H */

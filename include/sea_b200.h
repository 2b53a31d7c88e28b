/* sea_b200 — C ABI of the B200-native (sm_100a) SEA hot path.
 *
 * The reference (ParsaEsmati/SEA) has no FFI of its own: its hot path is the Python nn.Module
 * surface of models/temporal.py, models/base_blocks.py and models/encoder_decoder.py, which
 * dispatches into ATen.  Each entry point below replaces the ATen dispatches of one group of
 * reference call sites (cited per function as file:line into the reference tree) and is what a
 * maintainer binds from Python (ctypes stub in INTEGRATION.md; the shipped binding is
 * sea_b200/_lib.py).
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*;
 *  - the caller owns all memory (inputs, outputs, workspace); nothing here allocates device
 *    memory or synchronises the device; all work is enqueued on `stream` and is CUDA-graph
 *    capturable;
 *  - row-major everywhere; `ld*` are leading dimensions in ELEMENTS;
 *  - return value: 0 = ok, <0 = SEA_ERR_* (invalid argument / unsupported shape), >0 = cudaError_t;
 *  - there is no CPU implementation behind any of these calls.
 */
#ifndef SEA_B200_H_
#define SEA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* sea_stream_t;

enum {
  SEA_OK = 0,
  SEA_ERR_INVALID = -1,     /* null pointer, negative size, misaligned pointer/stride */
  SEA_ERR_UNSUPPORTED = -2, /* shape outside the supported set (see each function)     */
  SEA_ERR_NO_DEVICE = -3,   /* no sm_100 device / driver entry point missing           */
  SEA_ERR_WORKSPACE = -4    /* workspace too small                                      */
};

enum { SEA_ACT_NONE = 0, SEA_ACT_GELU = 1 };
enum { SEA_GEMM_A_MN = 1, SEA_GEMM_B_MN = 2 };
enum { SEA_PREC_BF16 = 0, SEA_PREC_FP32 = 1 };
enum { SEA_NORM_LN = 0, SEA_NORM_ADALN = 1 };
#define SEA_MAX_STREAMS 4

/* ------------------------------------------------------------------ library ---------------- */
int sea_init(int device);              /* caches SM count + driver entry points for `device`  */
const char* sea_strerror(int code);    /* static string for SEA_ERR_* / cudaError_t           */
int sea_version(void);
int sea_num_sms(void);
void sea_set_pdl(int on);           /* programmatic dependent launch between library kernels (default on) */

/* ------------------------------------------------------------------ K1: GEMM ----------------
 * C[M,N] = A[M,K] * B[N,K]^T with fused epilogue.  Replaces every nn.Linear on the path:
 *   q/k/v + RoPE        models/base_blocks.py:179-184, 271-276, 314-324
 *   projection + skip   models/base_blocks.py:201, 293 ; models/temporal.py:136
 *   cross_down/up       models/temporal.py:177-178, 185, 191
 *   MLP Linear 1 / 2    models/base_blocks.py:22-26, 44-47 ; models/temporal.py:145
 *   proj                models/temporal.py:146
 *   AdaLN cond_mlp[2]   models/base_blocks.py:337-341, 344
 * A, B are bf16 (tcgen05 kind::f16, fp32 accumulation in TMEM).  fp32-accurate products are
 * obtained by calling this on 3-way bf16 splits concatenated along K (sea_split3_*).
 *
 * v   = sum_k A[m,k] B[n,k] + bias[n] + residual[m,n]
 * v   = rope(v)                                  (columns < rope_cols, interleaved pairs)
 * v   = v * gelu'(gelu_grad_of[m,n])             (backward of GELU fused, optional)
 * out_f32[m,n]      = v                          (optional)
 * out_pre_bf16[m,n] = bf16(v)                    (optional)
 * out_bf16[m,n]     = bf16(act(v))               (optional)
 */
typedef struct sea_gemm_epilogue {
  const float* bias;        /* [N] or NULL */
  const float* residual;    /* [M, ld_residual] fp32 or NULL (may alias out_f32) */
  int64_t ld_residual;
  const void* gelu_grad_of; /* bf16 [M, ld_gelu] pre-activation saved by forward, or NULL */
  int64_t ld_gelu;
  int32_t act;              /* SEA_ACT_* applied to out_bf16 only */
  int32_t rope_cols;        /* 0 = no RoPE; else multiple of head_dim */
  int32_t head_dim;         /* multiple of 32 */
  int32_t seq_len;          /* position of row m is (m % seq_len); rope_sign=-1 rotates back */
  float rope_sign;          /* +1 forward, -1 inverse rotation (backward) */
  int32_t rope_ld;          /* positions per table row (>= seq_len) */
  const float* rope_table;  /* PAIR-MAJOR [head_dim/2, rope_ld, 2] fp32 (cos, sin): coalesced across rows */
  float* out_f32;
  int64_t ld_out_f32;
  void* out_pre_bf16;
  int64_t ld_out_pre_bf16;
  void* out_bf16;
  int64_t ld_out_bf16;
  int32_t rope_pos0;        /* position of row m is rope_pos0 + (m % seq_len): the incremental (KV-cached)
                               step rotates its single new token by its absolute position */
  int32_t res_rows_per_batch; /* > 0: the residual is a [B, *, ...] buffer with more rows per trajectory than this
                                 call covers: row m lives at residual + (m / res_rows_per_batch) * res_batch_stride
                                 + (m % res_rows_per_batch) * ld_residual */
  int64_t res_batch_stride;
  float dropout_p;          /* > 0: nn.Dropout on (acc + bias) BEFORE the residual is added
                               (MLP.forward, models/base_blocks.py:44-47 followed by the skip, temporal.py:145);
                               element (m, n) uses mask index m * N + n of `dropout_site` (sea_dropout_mask) */
  uint32_t dropout_site;
  uint64_t dropout_seed;
} sea_gemm_epilogue;

typedef struct sea_gemm_problem {
  const void* a; /* bf16 [M,K], lda */
  int64_t lda;
  const void* b; /* bf16 [N,K], ldb  (nn.Linear weight layout) */
  int64_t ldb;
  sea_gemm_epilogue epi;
  int32_t b_is_static; /* promise: B was completely written before the last NON-overlapped kernel
                          boundary of the stream (a weight matrix).  The TMA producer then requests its
                          first ring of B tiles before griddepcontrol.wait, so the weight fetch overlaps
                          the upstream kernel's tail.  0 = B may come from a kernel still in flight. */
  int32_t mn_major;    /* mask of SEA_GEMM_A_MN / SEA_GEMM_B_MN.  0: A [M,K], B [N,K] (K contiguous).
                          SEA_GEMM_B_MN: B is stored [K,N] (N contiguous, ldb = its row pitch) — the input
                          gradient dx = dy W reads the nn.Linear weight W [N_out, K_in] as it is.
                          A_MN | B_MN: A stored [K,M] and B [K,N]: C = A^T B — the weight gradient
                          dW = dy^T x (autograd of every nn.Linear on the path, train/train_temporal.py:257).
                          No transpose is materialised in either case (MN-major UMMA descriptors).
                          All problems of a launch must agree; A_MN needs M % 8 == 0. */
} sea_gemm_problem;

/* `num_problems` (1..4) same-shape problems in ONE launch (the V field streams are independent). */
int sea_gemm_bf16_tn(int num_problems, const sea_gemm_problem* host_problems, int M, int N, int K,
                     sea_stream_t stream);
/* Same, but the K loop is cut into chunks of `k_chunk` elements (multiple of 64; 0 = one chunk):
 * every chunk gets a fresh TMEM accumulator and the chunk sums are added in fp32 round-to-nearest
 * by the epilogue through out_f32 (required; it may alias residual for an in-place += ).  Used by the fp32-parity
 * mode: the tensor core's own accumulator does not round to nearest, so long K chains drift. */
int sea_gemm_bf16_tn_chunked(int num_problems, const sea_gemm_problem* host_problems, int M, int N,
                             int K, int k_chunk, sea_stream_t stream);
/* Stream-K workspace of the calling thread's later GEMM launches (caller-owned device memory, 256-byte
 * aligned; NULL = none).  The first 64 KB are per-tile arrival counters and MUST be zero when handed
 * over (the kernels leave them zero); the rest parks fp32 partial tiles: 2 x 128 x BN x 4 bytes per CTA,
 * about 39 MB makes every shape eligible.  With a workspace, launches whose tile count leaves SMs idle or
 * whose last wave is ragged split the flattened (tile, k-block) space evenly over the SMs; the CTA that
 * arrives last at a shared tile sums the partials in CTA order (deterministic) and runs the epilogue. */
int sea_gemm_set_workspace(void* workspace, size_t bytes);
/* Tuning hook: 1 = launch 256-wide-tile GEMMs as two-CTA clusters over vertically adjacent tiles, each CTA
 * fetching half of the shared B tile and TMA-multicasting it to both (halves B's L2 traffic); default 0. */
void sea_gemm_cluster(int on);
/* Tuning hook: 0 = never stream-K, 1 = when the cost model prefers it (default), 2 = whenever legal. */
void sea_gemm_stream_k(int mode);
/* Test / tuning hook: force the N-tile (64, 128, 256) of the next launches; 0 = heuristic. */
void sea_gemm_force_tile_n(int bn);
/* Tuning probe (results are garbage): 1 = skip the TMA traffic, 2 = skip the MMAs; 0 = normal. */
void sea_gemm_debug_probe(int mode);
/* Tuning probe: non-NULL = CTA 0 of every later GEMM launch records %globaltimer (ns) at 8 hand-off points into
 * dev_buf[0..7]: kernel entry, prologue done, griddepcontrol.wait passed (TMA warp), first stage landed (MMA warp),
 * last MMA committed, accumulator seen by the epilogue, epilogue stores issued, kernel exit.  NULL = off.  The probe
 * points are compiled in only when gemm.cu is built with -DSEA_GEMM_TRACE (they cost ~6 % otherwise). */
void sea_gemm_debug_trace(void* dev_buf);
/* Tuning probe: {tile width, grid, stream-K units per CTA (0 = data-parallel), tiles} of the last launch. */
void sea_gemm_last_config(int* out4);

/* ------------------------------------------------------------------ K4: row norms -----------
 * LayerNorm (custom, weight only: models/base_blocks.py:80-88) or AdaLN (:343-350) over the last
 * dim of fp32 rows; optional fused TIPI add in front (models/temporal.py:111-116, 140-142):
 *   x' = x + W3 g[m] + b3   (written to x_out),   y = norm(x')
 * cond (AdaLN) = output of cond_mlp, [M, 2d]: first half scales, second half shifts.
 * stats (optional) receives (mean, rstd) per row for the backward pass.  d % 4 == 0, d <= 2048. */
typedef struct sea_norm_args {
  const float* x;
  int64_t ldx;
  int32_t M, d, kind; /* SEA_NORM_* */
  const float* weight;
  const float* bias; /* AdaLN's own bias or NULL */
  const float* cond;
  int64_t ldc;
  int32_t cond_div;    /* row m uses cond row m / cond_div (0/1: one cond row per row; T: one per trajectory) */
  const float* add_rows; /* optional [ceil(M/add_div), d]: x' = x + add_rows[m / add_div] (written to x_out) */
  int64_t ld_add;
  int32_t add_div;
  const float* tipi_g; /* [M, tipi_hid] from sea_tipi_hidden, or NULL */
  int32_t tipi_hid;
  const float* tipi_w; /* [d, tipi_hid] */
  const float* tipi_b; /* [d] */
  float* x_out;
  int64_t ldxo;
  float* y_f32;
  int64_t ldy_f32;
  void* y_bf16;
  int64_t ldy_bf16;
  float* stats;
  int32_t cond_folded;       /* AdaLN: the cond rows already hold the effective gamma | beta
                                (sea_adaln_fold); weight / bias are then not read */
  int32_t x_rows_per_batch;  /* > 0: x is a [B, *, ...] buffer with more rows per trajectory than this call
                                covers: row m lives at x + (m / x_rows_per_batch) * x_batch_stride +
                                (m % x_rows_per_batch) * ldx (a prefix of a longer sequence buffer) */
  int64_t x_batch_stride;
  float tipi_dropout_p;      /* > 0: nn.Dropout on the TIPI term (W3 g + b3) before it is added to x
                                (the `ib` MLP's own dropout, models/base_blocks.py:47); mask index m * d + col */
  uint32_t tipi_dropout_site;
  uint64_t tipi_dropout_seed;
} sea_norm_args;
int sea_norm_fwd(const sea_norm_args* args, sea_stream_t stream);
/* `n` (1..SEA_MAX_STREAMS) norms of equal (M, d, kind) in ONE launch (the V field streams). */
int sea_norm_fwd_group(int n, const sea_norm_args* host_args, sea_stream_t stream);

/* In place: cond[r, :d] = weight + cond[r, :d] + 1, cond[r, d:] = bias + cond[r, d:] for R condition rows:
 * AdaLN's effective gamma | beta (models/base_blocks.py:345-350), folded once per trajectory when the
 * condition is time-invariant; consumed with sea_norm_args.cond_folded = 1. */
int sea_adaln_fold(float* cond, int64_t ldc, int R, int d, const float* weight, const float* bias,
                   sea_stream_t stream);

/* AdaLN cond_mlp[0] + SiLU on the scalar condition: h[m,j] = SiLU(w1[j,:]·ib[m,:] + b1[j]), j < n
 * (models/base_blocks.py:337-339, 344).  Either output may be NULL. */
int sea_adaln_hidden(const float* ib, int64_t ld_ib /* row pitch of ib, 0 = ib_num */, int M, int ib_num,
                     const float* w1, const float* b1, int n, void* out_bf16, float* out_f32,
                     sea_stream_t stream);

/* Same for up to 8 modules of equal width in ONE launch (out_bf16 / out_f32: arrays, entries may be NULL). */
int sea_adaln_hidden_group(int items, const float* const* w1, const float* const* b1,
                           void* const* out_bf16, float* const* out_f32, const float* ib, int64_t ld_ib,
                           int M, int ib_num, int n, sea_stream_t stream);

/* rows[r, :] = W3 g[r, :] + b3 for R distinct conditions (TIPI output per trajectory when the
 * condition is time-invariant; consumed through sea_norm_args.add_rows). */
int sea_tipi_rows(const float* g, int64_t ldg, int R, int E, int hid, const float* w3, const float* b3,
                  float* out, sea_stream_t stream);

/* TIPI hidden layer g = GELU(LayerNorm(W0 ib + b0)), hid <= 64 (models/base_blocks.py:22-25 as
 * instantiated by models/temporal.py:108).  pre_out / stats_out (optional) are saved for backward. */
int sea_tipi_hidden(const float* ib, int64_t ld_ib, int M, int ib_num, const float* w0, const float* b0,
                    const float* ln_w, const float* ln_b, int hid, float* g_out, float* pre_out,
                    float* stats_out, sea_stream_t stream);

/* Dropout helpers.  sea_dropout_mask writes the keep decisions (1 / 0) of elements [0, n) of `site` —
 * what the fused kernels apply on the fly — so a test can feed the identical masks to a reference
 * implementation.  sea_dropout_apply: dst[m, n] = src[m, n] * mask(m * N + n) / (1 - p) for an [M, N] fp32
 * matrix (row pitch ld_src), fp32 and / or bf16 destination (used by the backward pass for the gradients that
 * enter a dropped branch).  Site ids of the temporal executor: ((layer * 4 + kind) * 4 + i) * 4 + j with
 * kind 0 = self-attention of stream i, 1 = exchange attention (i, j), 2 = stream-MLP output, 3 = TIPI output. */
int sea_dropout_mask(uint64_t seed, uint32_t site, int64_t n, float p, uint8_t* out, sea_stream_t stream);
int sea_dropout_apply(const float* src, int64_t ld_src, int M, int N, uint64_t seed, uint32_t site, float p,
                      float* dst_f32, int64_t ld_f32, void* dst_bf16, int64_t ld_bf16, sea_stream_t stream);

/* ------------------------------------------------------------------ K6: LayerNorm(H) + GELU --
 * The MLP's inner nn.LayerNorm(H) (affine, eps 1e-5) followed by exact-erf GELU
 * (models/base_blocks.py:23-25).  bf16->bf16 or fp32->fp32.  H % 8 == 0, H <= 16384. */
typedef struct sea_ln_gelu_args {
  const void* h_bf16;
  const float* h_f32;
  int64_t ldh;
  int32_t M, H;
  const float* weight;
  const float* bias;
  void* g_bf16;
  float* g_f32;
  int64_t ldg;
  float* stats;
} sea_ln_gelu_args;
int sea_ln_gelu_fwd(const sea_ln_gelu_args* args, sea_stream_t stream);
/* `n` (1..SEA_MAX_STREAMS) problems of equal (M, H, ldh, ldg, dtype) in ONE launch. */
int sea_ln_gelu_fwd_group(int n, const sea_ln_gelu_args* host_args, sea_stream_t stream);

/* ------------------------------------------------------------------ operand packing ----------
 * src [R,C] (fp32 or bf16) -> bf16 dst, optionally transposed, optionally GELU'd on load,
 * optionally expanded to the 3-way bf16 split (split=1: A-side pattern, 2: B-side pattern; the
 * destination then has 6x the columns) that makes sea_gemm_bf16_tn fp32-accurate. */
typedef struct sea_pack_args {
  const float* src_f32;
  const void* src_bf16;
  int64_t ld;
  int32_t R, C;
  int32_t transpose, split, act;
  int32_t split_inner; /* 0: default (row length of dst / 6); else distance between split segments */
  void* dst;
  int64_t ld_dst;
} sea_pack_args;
int sea_pack_operand(const sea_pack_args* args, sea_stream_t stream);

/* out[n] += sum_m src[m,n]  (bias gradients). */
int sea_colsum_accumulate(const float* src_f32, const void* src_bf16, int64_t ld, int M, int N,
                          float* out, sea_stream_t stream);
/* n (<= SEA_MAX_STREAMS) column sums of equal shape in one launch; out[i] == NULL skips item i. */
int sea_colsum_accumulate_group(int n, const float* const* src_f32, const void* const* src_bf16, int64_t ld,
                                int M, int N, float* const* out, sea_stream_t stream);

/* ------------------------------------------------------------------ backward of K4/K5/K6 ----
 * The reference differentiates these with autograd; here they are explicit kernels.  Parameter
 * gradients ACCUMULATE (+=) into the given fp32 buffers (atomics). */
typedef struct sea_norm_bwd_args {
  const float* dy;      /* [M,d] upstream gradient */
  int64_t lddy;
  const float* x;       /* forward input rows */
  int64_t ldx;
  const float* stats;   /* (mean, rstd) from forward */
  int32_t M, d, kind;
  const float* weight;
  const float* cond;    /* AdaLN cond used in forward */
  int64_t ldc;
  const float* dres;    /* optional gradient of the skip connection, added to dx */
  int64_t lddres;
  float* dx;            /* optional fp32 output */
  int64_t lddx;
  void* dx_bf16;        /* optional bf16 copy (operand of the following GEMMs) */
  int64_t lddx_bf16;
  float* dweight;       /* [d] +=  (may be NULL) */
  float* dbias;         /* [d] +=  (AdaLN's own bias; may be NULL) */
  float* dcond;         /* AdaLN: [M,2d] gradient wrt cond (scale | shift) */
  int64_t lddcond;
  int32_t dcond_accumulate; /* 1: += into dcond (module applied twice in the exchange) */
  void* dcond_bf16;     /* AdaLN, optional: [M,2d] bf16 copy of the cond gradient (row pitch 2d), the operand of the
                           cond_mlp[2] backward GEMMs; with it `dcond` may be NULL.  Not with dcond_accumulate. */
  float* dweight2;      /* optional second destinations of the column sums (+=): for AdaLN, sum_m dcond[m, :] equals */
  float* dbias2;        /* (dweight | dbias), so cond_mlp[2].bias.grad = (dweight2 | dbias2) needs no extra pass */
} sea_norm_bwd_args;
int sea_norm_bwd(const sea_norm_bwd_args* args, sea_stream_t stream);
int sea_norm_bwd_group(int n, const sea_norm_bwd_args* host_args, sea_stream_t stream);  /* equal (M, d, kind) */

typedef struct sea_ln_gelu_bwd_args {
  const void* dg;   /* bf16 [M,H] gradient wrt GELU output */
  int64_t lddg;
  const void* h;    /* bf16 [M,H] pre-LayerNorm activations saved by forward */
  int64_t ldh;
  const float* stats;
  int32_t M, H;
  const float* weight;
  const float* bias;
  void* dh;         /* bf16 [M,H] */
  int64_t lddh;
  float* dweight;   /* [H] += */
  float* dbias;     /* [H] += */
  int32_t prec;     /* SEA_PREC_BF16 (0, default): dg / h / dh are bf16.  SEA_PREC_FP32: all three are fp32 and the
                       kernel uses erff / expf (the reference trains in fp32, train/train_temporal.py:252-258) */
} sea_ln_gelu_bwd_args;
int sea_ln_gelu_bwd(const sea_ln_gelu_bwd_args* args, sea_stream_t stream);
int sea_ln_gelu_bwd_group(int n, const sea_ln_gelu_bwd_args* host_args, sea_stream_t stream); /* equal shapes */

/* fp32-parity mode: d[m,n] *= gelu'(pre[m,n]) (exact erf form) — the GELU between the exchange attention's output
 * projection and cross_up (models/temporal.py:185); the bf16 path fuses this into the dgrad GEMM epilogue. */
int sea_gelu_grad_mul_f32(float* d, int64_t ldd, const float* pre, int64_t ldp, int M, int N, sea_stream_t stream);

/* Backward of sea_adaln_hidden: dh [M,n] -> dw1 [n, ib_num] +=, db1 [n] +=  (ib_num <= 4). */
int sea_adaln_hidden_bwd(const float* dh, int64_t lddh, const float* ib, int M, int ib_num,
                         const float* w1, const float* b1, int n, float* dw1, float* db1,
                         sea_stream_t stream);

/* Backward of the TIPI branch x_i += W3 GELU(LN(W0 ib + b0)) + b3 shared by all streams:
 * dx[i] are the per-stream gradients at the add; everything accumulates. hid <= 8, ib_num <= 4. */
typedef struct sea_tipi_bwd_args {
  const float* dx[SEA_MAX_STREAMS];
  int64_t lddx;
  int32_t n_streams, M, E, hid, ib_num;
  const float* g;     /* [M,hid] forward hidden (after GELU) */
  const float* u;     /* [M,hid] pre-LayerNorm */
  const float* stats; /* [M,2] */
  const float* ib;
  const float* w3;    /* [E,hid] */
  const float* ln_w;
  const float* ln_b;
  float *dw3, *db3, *dlnw, *dlnb, *dw0, *db0;
} sea_tipi_bwd_args;
int sea_tipi_bwd(const sea_tipi_bwd_args* args, sea_stream_t stream);

/* ------------------------------------------------------------------ K2: attention ------------
 * softmax(mask(Q K^T * scale)) V per (batch, head), causal with offset: key k is visible to
 * query q iff k <= q + src_len  (models/base_blocks.py:191-197 and 283-289; Q and K/V come from
 * different tensors for the state-exchange cross-attention).  Rows of q/k/v/o are (b*T + t),
 * head h occupies columns [h*head_dim, (h+1)*head_dim).  q,k are expected to be RoPE'd already
 * (fused into the projection GEMM epilogue).  lse (optional) = log-sum-exp of the scaled scores,
 * [B, n_heads, T], for the backward pass. */
typedef struct sea_attn_args {
  const void* q;
  const void* k;
  const void* v;
  int64_t ldq, ldk, ldv;
  void* o;
  int64_t ldo;
  float* lse;
  int32_t B, T, n_heads, head_dim;
  int32_t src_len;
  float scale;
  int32_t prec; /* SEA_PREC_BF16: bf16 in/out; SEA_PREC_FP32: fp32 in/out */
  float dropout_p;        /* > 0 (bf16 tensor-core path only): nn.Dropout on the attention probabilities after
                             the softmax (models/base_blocks.py:194, 286); element (b, h, q, k) uses mask index
                             ((b * n_heads + h) * T + q) * T2 + k, T2 = T rounded up to even */
  uint32_t dropout_site;
  uint64_t dropout_seed;
} sea_attn_args;
int sea_attention_fwd(const sea_attn_args* args, sea_stream_t stream);
/* Tuning switch: 1 (default) = bf16 problems with T <= 24 (head_dim 128) / 40 (head_dim 64) and no probability
 * dropout run on the short-sequence kernel (one small CTA per (batch, head), warp-level mma.sync); 2 = every
 * T <= 128 (tests); 0 = always the tcgen05 kernel. */
void sea_attention_small(int on);
/* `n` (1..SEA_MAX_STREAMS) problems of identical shape (B, T, n_heads, head_dim, src_len, scale,
 * prec, ldo) in ONE launch: the self-attention of the V independent field streams
 * (models/temporal.py:135-136 runs them one after the other). */
int sea_attention_fwd_group(int n, const sea_attn_args* host_args, sea_stream_t stream);
/* Decode attention for the KV-cached incremental step: ONE new query per (trajectory, head) against
 * the cached keys / values 0 .. n_keys-1 (the last row of the causal attention above).  q, o: one row
 * per trajectory (pitch ldq / ldo); k, v: [B, *, cols] caches with position pitch ldk / ldv and
 * trajectory pitch k/v_batch_stride.  `n` same-shape problems per launch. */
typedef struct sea_attn_decode_args {
  const void* q;
  const void* k;
  const void* v;
  void* o;
  int64_t ldq, ldo, ldk, ldv, k_batch_stride, v_batch_stride;
  int32_t B, n_keys, n_heads, head_dim;
  float scale;
  int32_t prec;
} sea_attn_decode_args;
int sea_attention_decode_group(int n, const sea_attn_decode_args* host_args, sea_stream_t stream);
/* K3: backward of sea_attention_fwd.  d_o is the gradient of o; delta [B,n_heads,T] is scratch.
 * dq/dk/dv use the same row/head layout.  If rope_table is given, dq and dk are rotated back by
 * -theta (the forward RoPE lives in the projection GEMM epilogue), so they are gradients with
 * respect to the PRE-RoPE projections. */
typedef struct sea_attn_bwd_args {
  const void* q;
  const void* k;
  const void* v;
  const void* o;
  const void* d_o;
  int64_t ldq, ldk, ldv, ldo, lddo;
  const float* lse;
  float* delta;
  void* dq;
  void* dk;
  void* dv;
  int64_t lddq, lddk, lddv;
  int32_t B, T, n_heads, head_dim;
  int32_t src_len;
  float scale;
  int32_t prec;
  const float* rope_table; /* pair-major [head_dim/2, rope_ld, 2] or NULL */
  int32_t rope_ld;
  float dropout_p;         /* must repeat the forward's dropout_p / site / seed (the mask is regenerated) */
  uint32_t dropout_site;
  uint64_t dropout_seed;
} sea_attn_bwd_args;
int sea_attention_bwd(const sea_attn_bwd_args* args, sea_stream_t stream);
/* Test hook: 1 forces the CUDA-core kernel even where the tcgen05 kernel applies. */
void sea_attention_force_simt(int on);
/* Tuning hook: 0 = one 128-query tile per CTA for every T (default 1: two tiles per CTA, overlapped, when T > 128). */
void sea_attention_two_tiles(int on);
/* tuning probe for the tensor-core backward (results are WRONG when non-zero): 1 = the compute warps only hand the
 * barriers over, 2 = they move S | dP through TMEM but skip the arithmetic; 0 = normal. */
void sea_attention_bwd_probe(int mode);
/* tuning hook: 1 (default) = 128-wide streamed tiles with a two-half hand-over of P | dS (head dims 64 / 128), 0 = the
 * 64-wide double-buffered plan of the same kernel family; results are identical up to summation order. */
void sea_attention_bwd_wide(int on);
/* Tuning hook: non-NULL = the two-tile forward kernel (head_dim 128, no dropout) records clock64 probes of CTA (0,0,0)
 * into dev_buf (3 x 64 x 8 int64: softmax group 0 / 1 and the MMA warp, per key tile); NULL switches it off. */
void sea_attention_debug_trace(void* dev_buf);

/* ------------------------------------------------------------------ fused AdamW --------------
 * torch.optim.AdamW.step as the reference configures it (utils/train_utils.py:33-39; called at
 * train/train_temporal.py:258), for a list of fp32 tensors in ONE launch:
 *   g' = g * grad_scale ; p *= 1 - lr*wd ; m = b1 m + (1-b1) g' ; v = b2 v + (1-b2) g'^2 ;
 *   p -= lr/bias_corr1 * m / (sqrt(v)/bias_corr2_sqrt + eps)
 * `chunks_dev` is a DEVICE array (built once by the caller while pointers are stable): every entry
 * covers <= 2^31 consecutive elements of one tensor; all four pointers 16-byte aligned.  p_bf16
 * (optional) receives the bf16 rounding of the updated values (the tensor-core copy of a weight,
 * see sea_temporal_cache_slot). */
typedef struct sea_adamw_chunk {
  float* p;
  const float* g;
  float* m;
  float* v;
  void* p_bf16;
  int32_t n;
  int32_t g_is_bf16; /* 1: `g` points at bf16 values (an averaged bf16 gradient bucket); 0: fp32 */
} sea_adamw_chunk;
typedef struct sea_adamw_hyper {
  float lr, beta1, beta2, eps, weight_decay;
  float bias_corr1;      /* 1 - beta1^t */
  float bias_corr2_sqrt; /* sqrt(1 - beta2^t) */
  float grad_scale;      /* 1, or 1/world to fold the data-parallel mean into the step */
  float one_minus_beta1; /* (1 - beta) evaluated in double by the caller, as torch does */
  float one_minus_beta2;
} sea_adamw_hyper;
int sea_adamw_step(const sea_adamw_chunk* chunks_dev, int num_chunks, const sea_adamw_hyper* hp,
                   sea_stream_t stream);
/* Same update with the step count in DEVICE memory (CUDA-graph capturable: nothing step-dependent is baked into the
 * launch): *step_dev (fp32, as torch keeps it) is incremented first, then bias_corr1 = 1 - beta1^t and
 * bias_corr2_sqrt = sqrt(1 - beta2^t) are evaluated on the device; hp->bias_corr* are ignored. */
int sea_adamw_step_dev(const sea_adamw_chunk* chunks_dev, int num_chunks, const sea_adamw_hyper* hp, float* step_dev,
                       sea_stream_t stream);
/* dst[i] = bf16(src[i]) / dst[i] = fp32(src[i]) over n contiguous elements (gradient-bucket compression of the
 * small, reduction-produced gradients; decompression of an averaged bf16 bucket back into param.grad). */
int sea_cast_f32_bf16(const float* src, void* dst_bf16, int64_t n, sea_stream_t stream);
int sea_cast_bf16_f32(const void* src_bf16, float* dst, int64_t n, sea_stream_t stream);

/* ------------------------------------------------------------------ temporal model -----------
 * Whole-model executor for TemporalModel with exchange_mode='sea', ib_scale_mode='mlp',
 * ib_addition_mode='add', add_info_after_cross=True (the mode both reference configs select):
 *   TemporalModel.forward                models/temporal.py:405-416
 *   BaseBlockTemporal.forward            models/temporal.py:126-148
 *   SEABlockTemporal._apply_exchange     models/temporal.py:176-192
 * The descriptor mirrors the reference module tree: every field is the fp32 master parameter
 * (`p`, the nn.Parameter's storage) and, for training, where its gradient goes (`g`, may be NULL).
 * Dead parameters of the reference do not appear. */
#define SEA_BWD_GROUPS 5
typedef struct sea_param {
  const float* p;
  float* g;
} sea_param;

typedef struct sea_norm_params { /* LayerNorm(weight) or AdaLN(weight, bias, cond_mlp) */
  sea_param weight, bias;
  sea_param c0_w, c0_b; /* cond_mlp.0: [2d, ib_num], [2d] */
  sea_param c2_w, c2_b; /* cond_mlp.2: [2d, 2d],    [2d] */
} sea_norm_params;

typedef struct sea_attn_params { /* q,k,v with bias; projection without */
  sea_param q_w, q_b, k_w, k_b, v_w, v_b, proj_w;
} sea_attn_params;

typedef struct sea_stream_params {
  sea_norm_params ln0, ln2, ln_cross;            /* ln.exp.{i}.0, ln.exp.{i}.2, ln_cross.{i} */
  sea_attn_params self_attn;                     /* attn.self.{i}  */
  sea_attn_params cross_attn[SEA_MAX_STREAMS];   /* cross_attn.{i}.{j}, j != i */
  sea_param down_w, down_b, up_w, up_b;          /* cross_down.{i}, cross_up.{i} */
  sea_param mlp0_w, mlp0_b, mlp_ln_w, mlp_ln_b, mlp3_w, mlp3_b; /* mlp.{i}.layers.{0,1,3} */
  sea_param proj_w, proj_b;                      /* proj.{i} */
} sea_stream_params;

typedef struct sea_block_params {
  sea_stream_params s[SEA_MAX_STREAMS];
  sea_param ib0_w, ib0_b, ib_ln_w, ib_ln_b, ib3_w, ib3_b; /* ib.layers.{0,1,3} (TIPI) */
} sea_block_params;

typedef struct sea_temporal_desc {
  int32_t num_layers, num_streams;
  int32_t embed_dim, n_heads, hidden_dim /* scale_ratio*E */, down_dim /* E/down_proj */;
  int32_t ib_num, ib_hidden /* max(1, scale_ratio*ib_num) */;
  int32_t norm_kind /* SEA_NORM_* */, src_len, max_len;
  int32_t precision /* SEA_PREC_BF16: bf16 tensor-core operands; SEA_PREC_FP32: 3-way split */;
  int32_t ib_time_invariant /* caller asserts ib[b,t,:] == ib[b,0,:] for all t (inference only):
                               AdaLN cond_mlp and the TIPI MLP are evaluated once per trajectory */;
  int32_t cond_cache_valid /* 1: cond_cache already holds the results for these B trajectories and the
                              current weights (set by the caller after a first call) -> skip that path */;
  void* cond_cache;        /* optional persistent device buffer (sea_temporal_cond_cache_bytes) used
                              only with ib_time_invariant; NULL = keep everything in the workspace */
  size_t cond_cache_bytes;
  const sea_block_params* blocks; /* host array [num_layers] */
  sea_norm_params final_ln[SEA_MAX_STREAMS]; /* ln.{i} */
  const float* rope_self;  /* device, pair-major [(E/n_heads)/2, max_len, 2] (cos, sin) */
  const float* rope_cross; /* device, pair-major [(down_dim/n_heads)/2, max_len, 2] */
  float dropout_p;         /* train mode (training = 1 forward and its backward): dropout probability of every
                              nn.Dropout on the path (attention probabilities, stream-MLP output, TIPI MLP output,
                              models/base_blocks.py:47, 194, 286); 0 in eval mode.  Masks are counter-based
                              functions of (dropout_seed, site, element), regenerated by the backward */
  uint32_t reserved2;
  uint64_t dropout_seed;   /* drawn by the caller once per forward; the backward must see the same value */
  int32_t grads_fresh;     /* sea_temporal_backward only.  1 = the caller asserts that the gradient buffers of
                              the nn.Linear WEIGHT matrices (everything the weight-gradient GEMMs write) hold no
                              value worth keeping (optimizer.zero_grad()): their first contribution overwrites
                              instead of accumulating, which saves zero-filling and re-reading 4 B/parameter.
                              Every other gradient (biases, norm / TIPI / cond_mlp.0 parameters) still
                              accumulates and must have been zeroed by the caller.  0 = accumulate everywhere. */
  int32_t splitk_slot;     /* 0..3: which of the cache's four stream-K workspaces this call's GEMMs use.  Calls that may
                              run concurrently on different streams (the micro-batches of a rollout plan) must use
                              different slots; the two-stream schedule's auxiliary stream takes (slot + 2) % 4. */
  /* ---- data-parallel training (sea_temporal_backward only; all optional, zero = off) ----
   * Every `g` of this descriptor points into ONE flat fp32 buffer starting at grad_f32_base (the all-reduce
   * bucket).  With grad_bf16 set, each weight-gradient GEMM also stores the bf16 rounding of the value it leaves
   * in its fp32 destination at the same element offset of grad_bf16 (no extra pass: second store of the epilogue),
   * so the gradient exchange can move half the bytes; gradients produced by reductions (biases, norm / TIPI /
   * cond_mlp.0 parameters) are NOT mirrored — the caller casts that (small) region itself (sea_cast_f32_bf16).
   * bwd_events[k] (cudaEvent_t or NULL) is recorded on the stream once the parameter gradients of group k are
   * final, so the exchange of that bucket can start while the rest of the backward runs:
   *   0 final norms ln.{i} | 1 stream MLP + proj | 2 ln.exp.{i}.2 + TIPI | 3 exchange (cross_*, ln_cross) |
   *   4 self-attention + ln.exp.{i}.0 (= end of the backward).  Groups 1-4 are recorded in the LAST processed layer. */
  const float* grad_f32_base;
  void* grad_bf16;
  void* bwd_events[SEA_BWD_GROUPS];
  /* ---- two-stream inference schedule (sea_temporal_forward / _strided / _step, bf16 mode; optional, NULL = off) ----
   * The exchange is sequential over the streams (Gauss-Seidel, models/temporal.py:187-192) and its steps are small,
   * latency-bound launches.  Once stream i has been exchanged, its TIPI / MLP / proj tail (models/temporal.py:140-146)
   * does not depend on the later streams: the executor enqueues it on aux_stream (a cudaStream_t), released by
   * fork_events[i] (cudaEvent_t, recorded on the caller's stream), and the caller's stream waits for join_event
   * (recorded on aux_stream) at the end of the block.  Capturable: the auxiliary stream joins a stream capture
   * through the events.  Same kernels, same arithmetic, same results. */
  void* aux_stream;
  void* fork_events[SEA_MAX_STREAMS];
  void* join_event;
} sea_temporal_desc;

/* Packed low-precision copies of the weights (bf16, fused QKV / KV, optional transposes for the
 * backward pass) live in a caller-owned cache; refresh after every parameter update. */
size_t sea_temporal_cache_bytes(const sea_temporal_desc* d, int training);
size_t sea_temporal_cond_cache_bytes(const sea_temporal_desc* d, int B);
int sea_temporal_refresh(const sea_temporal_desc* d, void* cache, size_t cache_bytes, int training,
                         sea_stream_t stream);
/* Partial refresh after an optimizer step that already wrote the straight bf16 copies itself
 * (sea_adamw_step with p_bf16 from sea_temporal_cache_slot): `what` is a mask of SEA_REFRESH_*. */
enum { SEA_REFRESH_STRAIGHT = 1, SEA_REFRESH_TRANSPOSED = 2, SEA_REFRESH_BIASES = 4, SEA_REFRESH_ALL = 7 };
int sea_temporal_refresh_ex(const sea_temporal_desc* d, void* cache, size_t cache_bytes, int training,
                            int what, sea_stream_t stream);
/* Where the straight bf16 copy [N,K] (contiguous) of the fp32 master `master` lives inside `cache`
 * (SEA_PREC_BF16 only); *dst = NULL when that parameter has no tensor-core copy (biases, norms). */
int sea_temporal_cache_slot(const sea_temporal_desc* d, void* cache, int training, const float* master,
                            void** dst);
/* Activations / saved-for-backward tape live in a caller-owned workspace. */
size_t sea_temporal_workspace_bytes(const sea_temporal_desc* d, int B, int T, int training);
/* x [B,T,V,E] fp32 contiguous, ib [B,T,ib_num] fp32, y [B,T,V,E] fp32. */
int sea_temporal_forward(const sea_temporal_desc* d, const void* cache, const float* x,
                         const float* ib, float* y, int B, int T, void* workspace,
                         size_t workspace_bytes, int training, sea_stream_t stream);
/* KV-cached incremental step (the rollout loop of utils/train_utils.py:202-209 without the O(n^2)
 * prefix recompute): feeds ONE new token per trajectory, x_t [B,V,E] at absolute position `pos`, and
 * returns y_t = TemporalModel.forward(x[:, :pos+1])[:, pos].  Exact for a causal model: positions
 * 0..pos-1 must have been fed before with the same kv_cache (sea_temporal_kv_cache_bytes(d, B, max_len)
 * bytes, caller-owned; holds the RoPE'd q|k|v rows of the self-attention and the k|v rows of every
 * state-exchange pair).  x_t / ib_t / y_t are addressed with a pitch between trajectories, so they can
 * be slices of a [B, T, ...] sequence buffer.  Workspace: sea_temporal_workspace_bytes(d, B, 1, 0).
 * With d->ib_time_invariant and a cond_cache the AdaLN / TIPI condition path is evaluated once per
 * trajectory (cond_cache_valid as for sea_temporal_forward). */
size_t sea_temporal_kv_cache_bytes(const sea_temporal_desc* d, int B, int max_len);
int sea_temporal_step(const sea_temporal_desc* d, const void* cache, void* kv_cache, size_t kv_cache_bytes,
                      int max_len, const float* x_t, int64_t x_batch_stride, const float* ib_t,
                      int64_t ib_batch_stride, float* y_t, int64_t y_batch_stride, int B, int pos,
                      void* workspace, size_t workspace_bytes, sea_stream_t stream);
/* sea_temporal_forward on a prefix of a longer sequence buffer: x is [B, >= T, V, E] with
 * `x_batch_stride` elements between trajectories (the rollout loop keeps ONE [B, steps+1, V, E] buffer and
 * runs the model on x[:, :T] without gathering the prefix first).  Inference only. */
int sea_temporal_forward_strided(const sea_temporal_desc* d, const void* cache, const float* x,
                                 int64_t x_batch_stride, const float* ib, float* y, int B, int T,
                                 void* workspace, size_t workspace_bytes, sea_stream_t stream);
/* Backward of the last sea_temporal_forward(..., training=1) that used this `workspace`
 * (autograd in the reference, train/train_temporal.py:257).  dy [B,T,V,E] fp32 contiguous.
 * Parameter gradients ACCUMULATE (+=) into the `g` pointers of the descriptor (NULL = frozen);
 * dx (optional) receives dL/dx.  SEA_PREC_BF16 only; the cache must have been refreshed with
 * training=1 (it then holds the transposed weights for dgrad). */
int sea_temporal_backward(const sea_temporal_desc* d, const void* cache, const float* x,
                          const float* ib, const float* dy, float* dx, int B, int T,
                          void* workspace, size_t workspace_bytes, sea_stream_t stream);
/* Number of kernels the last forward / backward call on this thread launched. */
int sea_last_launch_count(void);
/* cudaEvent_t plumbing for hosts without a runtime binding: create (timing disabled) / destroy / make `stream`
 * wait for the event.  Return 0 or a cudaError_t. */
int sea_event_create(void** out_event);
int sea_event_destroy(void* cuda_event);
int sea_stream_wait_event(sea_stream_t stream, void* cuda_event);
/* Strided device -> host copy for hosts without a runtime binding: `rows` rows of `row_bytes` bytes from a device buffer
 * with row pitch `src_pitch` into PINNED host memory with row pitch `dst_pitch`, asynchronously on `stream` (one
 * cudaMemcpy2DAsync).  The graphed rollout hands every finished group of steps to the host with it while the next group
 * still runs (the results of utils/train_utils.py:209 `torch.cat(preds)` reach the host without a trailing copy).
 * Returns 0 or a cudaError_t. */
int sea_copy_rows_to_host(void* dst_host, size_t dst_pitch, const void* src_dev, size_t src_pitch, size_t row_bytes,
                          size_t rows, sea_stream_t stream);

/* Optional per-launch timing (CUDA events on the launching stream) for roofline accounting.
 * Between begin and end every executor launch is bracketed by an event pair; end synchronises on
 * the events and returns, per category, total device milliseconds, algorithmic work (FLOPs for
 * GEMM / attention, bytes for the HBM-bound kernels) and launch counts. */
enum { SEA_PROF_GEMM = 0, SEA_PROF_ATTN = 1, SEA_PROF_ELEMWISE = 2, SEA_PROF_NUM = 3 };
typedef struct sea_profile_summary {
  double ms[SEA_PROF_NUM];
  double work[SEA_PROF_NUM];
  int64_t launches[SEA_PROF_NUM];
} sea_profile_summary;
void sea_profile_begin(void);
int sea_profile_end(sea_profile_summary* out);

/* ------------------------------------------------------------------ K7 / K8: ViT-mesh codec ---
 * Fused SpatialModel halves, one CTA per snapshot (models/encoder_decoder.py):
 *   sea_spatial_encode = generate_padding_mask (:173-176, in place on x) + PointwiseEncode.forward
 *                        (:105-123): per-group patch MLP, + positional encoding, num_layers x
 *                        EncoderBlock (base_blocks.py:123-138), final nn.LayerNorm;
 *   sea_spatial_decode = Decode.forward (:137-146).
 * x / out: [B, 64, n_fields, n_inp] fp32; z: [B, 64, G, D] (latent_layout 0) or the temporal
 * model's [B, G, 64*D] (latent_layout 1 = transform_processed_data, utils/train_utils.py:315-337,
 * fused into the store / load).  All parameters fp32, reference layouts.  n_patches must be 64,
 * embed_dim / mlp_hidden multiples of 4 (n_inp arbitrary; multiples of 4 take the vector path), head_dim = G*D/n_heads in {2,4,8,16}, <= 16 layers. */
typedef struct sea_spatial_layer {
  const float *ln1_w, *q_w, *q_b, *k_w, *k_b, *v_w, *v_b, *proj_w, *ln2_w;
  const float *mlp0_w, *mlp0_b, *mlp_ln_w, *mlp_ln_b, *mlp3_w, *mlp3_b;
} sea_spatial_layer;

typedef struct sea_spatial_desc {
  int32_t n_groups, n_fields, n_inp, n_patches, mlp_hidden, embed_dim, n_heads, num_layers;
  int32_t group_first_field[4], group_num_fields[4]; /* field groups must be contiguous runs */
  const float* enc_w1[4]; /* encode.encoders.{g}.layer1.weight [Hs, n_inp*|g|] */
  const float* enc_w2[4]; /* .layer2.weight [D, Hs] */
  const float* enc_b2[4]; /* .layer2.bias   [D]     */
  const float* dec_w1[4]; /* decode.decoders.{g}.layer1.weight [Hs, D] */
  const float* dec_w2[4]; /* .layer2.weight [n_inp*|g|, Hs] */
  const float* dec_b2[4];
  const sea_spatial_layer* layers; /* host array [num_layers]: encode.blocks.{l}.* */
  const float* ln_w;               /* encode.ln */
  const float* ln_b;
  const float* pe;                 /* encode.spatial_pos_encoder.pe[0, :64, :]  [64, G*D] */
} sea_spatial_desc;

int sea_spatial_encode(const sea_spatial_desc* d, float* x, float* z, int B, int latent_layout,
                       float pad_idx, int fix_pad, sea_stream_t stream);
int sea_spatial_decode(const sea_spatial_desc* d, const float* z, float* out, int B, int latent_layout,
                       sea_stream_t stream);
/* The same two passes with every contraction (patch MLPs, q|k|v / projection / MLP of the encoder blocks, QK^T and P.V
 * of the 8-head attention) on the tensor cores: warp-level bf16 mma.sync.m16n8k16, fp32 accumulation; residual state,
 * LayerNorm statistics, softmax and GELU in fp32 (the bf16 parity mode: 2e-2 bar; the fp32 kernels above keep the 1e-4
 * bar).  Weights are rounded to bf16 once into a caller-owned cache (256-byte aligned, sea_spatial_cache_bytes):
 * sea_spatial_pack must run after every parameter update.  Needs embed_dim and mlp_hidden multiples of 16 (n_inp is
 * arbitrary: its axis is zero-padded to a multiple of 16 at pack time); SEA_ERR_UNSUPPORTED when a snapshot's working set exceeds the 227 KB of shared memory. */
size_t sea_spatial_cache_bytes(const sea_spatial_desc* d);
int sea_spatial_pack(const sea_spatial_desc* d, void* cache, size_t cache_bytes, sea_stream_t stream);
int sea_spatial_encode_tc(const sea_spatial_desc* d, const void* cache, float* x, float* z, int B, int latent_layout,
                          float pad_idx, int fix_pad, sea_stream_t stream);
int sea_spatial_decode_tc(const sea_spatial_desc* d, const void* cache, const float* z, float* out, int B,
                          int latent_layout, sea_stream_t stream);

/* ------------------------------------------------------------------ mesh patchify / unpatch ----
 * DataPartitioner2D of the reference (utils/data_processors.py:9-111; called from patchify_and_scale
 * :484-542 and inverse_scale_and_unpatch :553-573) on the device.  All index work is exact:
 *   sea_patch_bucketize   patch_id[c] = (ix-1)*(n-1) + (iy-1) with ix = clamp(bucketize(x[c], x_boundary,
 *                         right=True), 1, m-1) (:34-38), and the cell count of every patch (counts is
 *                         zeroed by the call; max(counts) is the padded patch length `capacity`);
 *   sea_patch_index_map   index_map[p, :] = ascending cell indices of patch p (mask.nonzero(), :43-45),
 *                         padded to `capacity` with pad_id (< 0) (:61-88);
 *   sea_patch_gather      out[s, p, c, f] = fields[f][s, index_map[p, c]], pad -> pad_value  (layout_pfc = 0:
 *                         the reference's stacked [S, P, C, F], :529; layout_pfc = 1: [S, P, F, C], the
 *                         SpatialModel input); fields[f] is [S, n_cells] with row pitch ld_field; F <= 8;
 *   sea_patch_scatter     inverse_partition (:90-111): out[s, index_map[p, c], f] = part[s, p, c, f]. */
int sea_patch_bucketize(const float* x, const float* y, int n_cells, const float* x_boundary, int m,
                        const float* y_boundary, int n, int32_t* patch_id, int32_t* counts, sea_stream_t stream);
/* DataPartitioner3D (utils/data_processors.py:114-165): the same with a third axis, patch_id = ((ix-1)*(n-1) + (iy-1))*(k-1)
 * + (iz-1), (m-1)(n-1)(k-1) <= 65535 patches; index_map / gather / scatter are shared with the 2-D partitioner. */
int sea_patch_bucketize3d(const float* x, const float* y, const float* z, int n_cells, const float* x_boundary, int m,
                          const float* y_boundary, int n, const float* z_boundary, int k, int32_t* patch_id,
                          int32_t* counts, sea_stream_t stream);
int sea_patch_index_map(const int32_t* patch_id, int n_cells, int n_patches, int capacity, int64_t pad_id,
                        int64_t* index_map, sea_stream_t stream);
int sea_patch_gather(const float* const* host_field_ptrs, int n_fields, int64_t ld_field, const int64_t* index_map,
                     int n_snapshots, int n_patches, int capacity, float pad_value, int layout_pfc, float* out,
                     sea_stream_t stream);
int sea_patch_scatter(const float* part, const int64_t* index_map, int n_snapshots, int n_patches, int capacity,
                      int n_fields, int n_cells, int layout_pfc, float* out, sea_stream_t stream);
/* The same two passes fused with the per-field-group MinMaxScaler of MeshProcessor (utils/data_processors.py:225-272,
 * _scale_fields :544-551, inverse_scale_and_unpatch :553-573) and reading / writing the reference's own interleaved
 * field tensor [S, n_cells, F] (no per-field copies): the resident fields -> latents -> fields pipeline.
 *   gather_scaled   out[s,p,..] = ((x - min) / (max - min)) * range + lo     (transform, :245-252; pad -> pad_value)
 *   scatter_scaled  out[s,idx,f] = ((v - lo) / range) * (max - min) + min     (inverse_transform, :258-272)
 * evaluated as the same sequence of IEEE fp32 operations torch performs (bit-identical).  enabled = 0: identity. */
typedef struct sea_field_scaler {
  float min_val, max_val; /* fitted on the training data (MinMaxScaler.fit) */
  float lo, range;        /* feature_range[0], feature_range[1] - feature_range[0] */
  int32_t enabled;
  int32_t reserved;
} sea_field_scaler;
int sea_patch_gather_scaled(const float* fields, int n_cells, int n_fields, const sea_field_scaler* host_scalers,
                            const int64_t* index_map, int n_snapshots, int n_patches, int capacity, float pad_value,
                            int layout_pfc, float* out, sea_stream_t stream);
int sea_patch_scatter_scaled(const float* part, const int64_t* index_map, int n_snapshots, int n_patches, int capacity,
                             int n_fields, int n_cells, int layout_pfc, const sea_field_scaler* host_scalers,
                             float* out, sea_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SEA_B200_H_ */

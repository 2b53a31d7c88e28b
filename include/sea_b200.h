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
enum { SEA_PREC_BF16 = 0, SEA_PREC_FP32 = 1 };
enum { SEA_NORM_LN = 0, SEA_NORM_ADALN = 1 };

/* ------------------------------------------------------------------ library ---------------- */
int sea_init(int device);              /* caches SM count + driver entry points for `device`  */
const char* sea_strerror(int code);    /* static string for SEA_ERR_* / cudaError_t           */
int sea_version(void);
int sea_num_sms(void);

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
  const float* rope_table;  /* [>=seq_len, head_dim/2, 2] fp32 (cos, sin) */
  float* out_f32;
  int64_t ld_out_f32;
  void* out_pre_bf16;
  int64_t ld_out_pre_bf16;
  void* out_bf16;
  int64_t ld_out_bf16;
} sea_gemm_epilogue;

typedef struct sea_gemm_problem {
  const void* a; /* bf16 [M,K], lda */
  int64_t lda;
  const void* b; /* bf16 [N,K], ldb  (nn.Linear weight layout) */
  int64_t ldb;
  sea_gemm_epilogue epi;
} sea_gemm_problem;

/* `num_problems` (1..4) same-shape problems in ONE launch (the V field streams are independent). */
int sea_gemm_bf16_tn(int num_problems, const sea_gemm_problem* host_problems, int M, int N, int K,
                     sea_stream_t stream);
/* Test / tuning hook: force the N-tile (64, 128, 256) of the next launches; 0 = heuristic. */
void sea_gemm_force_tile_n(int bn);

#ifdef __cplusplus
}
#endif
#endif /* SEA_B200_H_ */

// Internal structures of the temporal executor (shared by forward and backward).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <vector>

#include "../../include/sea_b200.h"

namespace sea {

using bf16 = __nv_bfloat16;

// Bump allocator over a caller-owned buffer; base == nullptr only measures.
struct Arena {
  char* base;
  size_t off = 0;
  void* take(size_t bytes) {
    const size_t a = (off + 255) & ~static_cast<size_t>(255);
    off = a + bytes;
    return base ? base + a : nullptr;
  }
};

// One nn.Linear (or a fused group of them) in tensor-core layout.
struct PackedLinear {
  const bf16* w;   // [N, kf*K]  (kf = 6 in the fp32 split mode); dgrad reads it as an MN-major operand
  long long ldw;
  int N, K;
};

struct StreamCache {
  PackedLinear qkv, sproj, down, up, mlp0, mlp3, proj;
  float* qkv_bias;
  PackedLinear cq[SEA_MAX_STREAMS], ckv[SEA_MAX_STREAMS], cproj[SEA_MAX_STREAMS];
  float* ckv_bias[SEA_MAX_STREAMS];
  PackedLinear c2_ln0, c2_ln2, c2_lnc;
};
struct BlockCache { StreamCache s[SEA_MAX_STREAMS]; };
constexpr size_t kSplitKBytes = 65536 + 148ull * 2 * 128 * 256 * 4;  // counters + two 128x256 fp32 slots per SM
struct CacheLayout {
  std::vector<BlockCache> blocks;
  PackedLinear c2_final[SEA_MAX_STREAMS];
  // stream-K workspaces (sea_gemm_set_workspace): GEMMs that may run CONCURRENTLY — the micro-batches of a rollout
  // plan on their own streams (desc->splitk_slot), the auxiliary stream of the two-stream schedule — need their own
  void* splitk[4];
  size_t splitk_bytes;
};

// "act" buffers are bf16 in SEA_PREC_BF16 and fp32 in SEA_PREC_FP32.
struct StreamTape {
  void *hid0, *hid2, *hidc;        // AdaLN SiLU(cond_mlp.0(ib))           [M,2E] [M,2E] [M,2Dd]
  float *cond0, *cond2, *condc;    // AdaLN cond_mlp outputs (scale|shift)  fp32
  void* n0; float* st0;            // Norm_{i,0}(x_i), (mean, rstd)
  void* qkv;                       // [M,3E]  RoPE'd q | RoPE'd k | v
  void* ao; float* lse;            // attention output [M,E], log-sum-exp [B,nh,T]
  float* x1; bf16* x1b;            // x_i after self-attention (fp32 + bf16 operand copy)
  float* dpre; float* stc_pre; void* npre;     // cross_down(x1), stats, ln_cross(...)
  float* dpost; float* stc_post; void* npost;  // same on the exchanged stream (i < V-1)
  void* q[SEA_MAX_STREAMS]; void* kv[SEA_MAX_STREAMS]; void* a[SEA_MAX_STREAMS];
  float* lse_c[SEA_MAX_STREAMS];
  void* p[SEA_MAX_STREAMS]; bf16* g[SEA_MAX_STREAMS];  // cross projection pre-GELU / GELU
  float* xp; bf16* xpb;            // x_i after the exchange
  float* x2; void* n2; float* st2; // x_i + TIPI, Norm_{i,2}
  void* h; float* stH; void* gh;   // MLP hidden pre-LN [M,H], stats, GELU(LN(h))
  void* x3;                        // x2 + MLP (operand of proj)
  float* xout;                     // proj output = block output
};
struct LayerTape {
  StreamTape s[SEA_MAX_STREAMS];
  float *tipi_g, *tipi_pre, *tipi_st;
  float* tipi_rows;  // [B,E] TIPI output per trajectory (time-invariant condition)
};
struct Tape {
  std::vector<LayerTape> L;
  void* hidF[SEA_MAX_STREAMS];
  float* condF[SEA_MAX_STREAMS];
  float* stF[SEA_MAX_STREAMS];
  bf16* packA[SEA_MAX_STREAMS];  // fp32 mode: split A operands
};

// Gradient buffers of the backward pass (one set, reused across layers; bf16 mode only).
// "b" twins are the bf16 tensor-core operands (NULL in SEA_PREC_FP32, where the fp32 buffer is the operand);
// void* buffers exist in one copy, in the mode's activation dtype (bf16 / fp32).
struct BwdStream {
  float* dxout; bf16* dxoutb;   // gradient at the block output / previous layer's input
  float* dx3; bf16* dx3b;       // at x3 = x2 + MLP
  void* dg; void* dh;           // at GELU output / at the MLP hidden pre-LN   [M,H]
  float* dn2;                   // at Norm_{i,2} output
  float* dx2; bf16* dx2b;       // at x2 = x_post + TIPI
  float* dxp; bf16* dxpb;       // at x_post when the exchanged stream feeds later streams
  void *dp, *da, *dq, *dkv;     // exchange branch: pre-GELU, attention out, q, k|v
  float *dnpre, *dnpost;        // at ln_cross outputs (accumulated over consumers)
  float* ddn; bf16* ddnb;       // at cross_down outputs
  float* dx1; bf16* dx1b;       // at x1 (after self-attention)
  void* dao; void* dqkv;        // self-attention: at attention output, at q|k|v (RoPE undone)
  float* dn0;                   // at Norm_{i,0} output
  float *dcond0, *dcond2, *dcondc, *dcondF;  // AdaLN: at the cond_mlp outputs
};
struct BwdTape {
  BwdStream s[SEA_MAX_STREAMS];
  bf16* dcb[SEA_MAX_STREAMS];    // bf16 copy of dcond
  float* dhid[SEA_MAX_STREAMS];  // gradient at SiLU output
  float* delta;                  // attention backward scratch
  bf16 *pack1, *pack2;           // SEA_PREC_FP32: split-operand scratch of the 3x-bf16 backward GEMMs
};
void layout_bwd_tape(const sea_temporal_desc* d, int B, int T, Arena& ar, BwdTape& t);

struct Ctx {
  const sea_temporal_desc* d;
  const CacheLayout* cache;
  Tape* tape;
  cudaStream_t s;
  bool fp32;
  int B, T, M;
  int Mc;            // rows of the condition path (M, or B when ib is time-invariant)
  long long ld_ib;   // row pitch of ib for the condition path
  int cond_div;      // token row m uses condition row m / cond_div
  bool inv;          // condition path evaluated once per trajectory (time-invariant ib)
  float drop_p;      // train-mode dropout probability of this call (0 in eval / inference)
  int pos0;          // absolute position of row 0 of every trajectory (KV-cached step: the new token's)
  void* splitk;      // stream-K workspace of the GEMMs this context launches (NULL: leave the current one)
};

struct LinIn {
  const void* a;  // act dtype [M, K]
  long long lda;
  int act_on_load;  // fp32 mode only: GELU applied while packing
};
struct LinOut {
  const float* bias;
  const float* residual; long long ld_res;
  int res_rows; long long res_bs;   // residual is a prefix of a longer per-trajectory buffer (see sea_gemm_epilogue)
  float* f32; long long ld_f32;   // fp32 copy of v
  void* pre; long long ld_pre;    // act-dtype copy of v
  void* post; long long ld_post;  // act-dtype copy of act(v)
  int act;
  int rope_cols, head_dim; const float* rope_table;
  float drop_p; unsigned drop_site;   // nn.Dropout on (acc + bias) before the residual (train mode)
};

void layout_cache(const sea_temporal_desc* d, bool training, Arena& ar, CacheLayout& c);
void layout_tape(const sea_temporal_desc* d, int B, int T, bool training, Arena& ar, Tape& t);
void layout_cond_cache(const sea_temporal_desc* d, int B, Arena& ar, Tape& t);
int linear_group(Ctx& c, int n, const LinIn* in, const PackedLinear* const* W, const LinOut* out, int Mrows);

extern thread_local int g_launches;

// The stream-K workspace of the GEMM launches is thread-local state of the library (sea_gemm_set_workspace): an executor
// call points it into the caller's weight cache and MUST take it back before returning — the cache may be freed right
// after the call, and a later stand-alone sea_gemm_bf16_tn on this thread would park partial tiles in freed memory.
struct SplitKGuard {
  ~SplitKGuard() { sea_gemm_set_workspace(nullptr, 0); }
};

// dropout site ids of the temporal executor (see sea_dropout_mask)
inline unsigned drop_site(int layer, int kind, int i, int j) { return static_cast<unsigned>(((layer * 4 + kind) * 4 + i) * 4 + j); }
enum { SEA_SITE_SELF = 0, SEA_SITE_CROSS = 1, SEA_SITE_MLP = 2, SEA_SITE_TIPI = 3 };

// RAII CUDA-event pair around one launch when profiling is on (no-op otherwise).
// `work` is the launch's algorithmic work: FLOPs for GEMM / attention, bytes for HBM-bound kernels.
class ProfScope {
 public:
  ProfScope(cudaStream_t s, int cat, double work);
  ~ProfScope();
 private:
  cudaStream_t s_;
  int idx_;
};

}  // namespace sea

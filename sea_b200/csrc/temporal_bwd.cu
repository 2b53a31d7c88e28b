// Backward pass of the temporal executor (the reference gets it from autograd,
// train/train_temporal.py:257).  Reads the tape the forward left in `workspace`, accumulates (+=)
// parameter gradients into the `g` pointers of the descriptor and optionally writes dL/dx.
//
// Every nn.Linear y = a W^T + b contributes three tensor-core GEMMs / reductions:
//   dgrad  da = dy · W          gemm_tn(A = dy [M,N],     B = W^T [K,N])   (W^T cached at refresh)
//   wgrad  dW += dy^T · a       gemm_tn(A = dy^T [N,M],   B = a^T [K,M])   (operands transposed
//                                                                            by the packing kernel)
//   dbias  db += colsum(dy)
// GELU' of the exchange branch is fused into the dgrad epilogue; RoPE is undone inside the
// attention backward; LN / AdaLN / LN+GELU / TIPI have dedicated backward kernels.
//
// precision = SEA_PREC_FP32 (the reference trains in fp32, no AMP: train/train_temporal.py:252-258): the tape and
// every gradient buffer are fp32, attention / norms / LN+GELU run their fp32 kernels, and each of the three GEMMs
// above is made fp32-accurate the way the forward's are: both operands are split into three bf16 terms, the six
// significant cross products are concatenated along the contraction dimension and summed by the SAME tcgen05
// kernel, 512 contraction columns per fresh TMEM accumulator (linear_bwd_fp32 below).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <atomic>

#include <vector>

#include "../../include/sea_b200.h"
#include "internal.h"
#include "temporal_internal.h"

namespace sea {
namespace {

#define SEA_TRY(expr)              \
  do {                             \
    int _rc = (expr);              \
    if (_rc != SEA_OK) return _rc; \
  } while (0)

struct BCtx {
  Ctx c;
  const float* ib;
  BwdTape* bt;
  bool fp32 = false;                  // SEA_PREC_FP32: fp32 operands / gradients, split GEMMs
  // the GEMM operand copy of a gradient / activation: its bf16 twin, or (fp32 mode) the fp32 buffer itself
  const void* op(const float* f, const bf16* h) const { return fp32 ? static_cast<const void*>(f) : static_cast<const void*>(h); }
  bf16* gb16 = nullptr;               // desc->grad_bf16: bf16 twin of the flat gradient buffer (data-parallel buckets)
  const float* gf32 = nullptr;
  bf16* twin_of(const float* g) const {   // where the bf16 copy of gradient element g lives (NULL: not mirrored)
    return (gb16 != nullptr && g != nullptr && g >= gf32) ? gb16 + (g - gf32) : nullptr;
  }
  cudaError_t mark(int group) const { // "parameter gradients of `group` are final" (desc->bwd_events)
    void* ev = c.d->bwd_events[group];
    if (ev == nullptr) return cudaSuccess;
    return cudaEventRecord(static_cast<cudaEvent_t>(ev), c.s);
  }
  bool fresh = false;                 // desc->grads_fresh: first wgrad contribution overwrites
  std::vector<const float*> touched;  // weight gradients already written by this call
  sea_stream_t st() const { return reinterpret_cast<sea_stream_t>(c.s); }
};

int cast_f32(BCtx& b, const float* src, long long ld, int R, int Ccols, bf16* dst, long long ldd, int transpose) {
  sea_pack_args a{};
  a.src_f32 = src; a.ld = ld; a.R = R; a.C = Ccols; a.transpose = transpose;
  a.dst = dst; a.ld_dst = ldd;
  ++g_launches;
  ProfScope prof(b.c.s, SEA_PROF_ELEMWISE, 6.0 * R * Ccols);
  return sea_pack_operand(&a, b.st());
}

// bias gradients of the n grouped Linears in one launch (NULL destinations are skipped)
int colsum_group(BCtx& b, int n, const float* const* f32, const void* const* b16, long long ld, int M, int N,
                 float* const* out) {
  bool any = false;
  for (int g = 0; g < n; ++g) any = any || out[g] != nullptr;
  if (!any) return SEA_OK;
  const void* bp[SEA_MAX_STREAMS];
  for (int g = 0; g < n; ++g) bp[g] = b16 ? b16[g] : nullptr;
  ++g_launches;
  ProfScope prof(b.c.s, SEA_PROF_ELEMWISE, (f32 ? 4.0 : 2.0) * M * N * n);
  return sea_colsum_accumulate_group(n, f32, b16 ? bp : nullptr, ld, M, N, out, b.st());
}

int gemm(BCtx& b, int n, sea_gemm_problem* probs, int M, int N, int K) {
  ++g_launches;
  ProfScope prof(b.c.s, SEA_PROF_GEMM, 2.0 * M * static_cast<double>(N) * K * n);
  return sea_gemm_bf16_tn(n, probs, M, N, K, b.st());
}

// One Linear's backward for up to V streams at once (same shapes).
struct LinB {
  const void* dy; long long lddy;   // [M,N] gradient wrt the Linear's output   (act dtype: bf16, or fp32 in fp32 mode)
  const void* a; long long lda;     // [M,K] saved input                          (act dtype)
  int a_act;                        // fp32 mode: activation applied to `a` on load (the forward's GELU-on-pack)
  const PackedLinear* W;
  const float* Wm[3];               // fp32 masters of the (n_split fused) Linear(s): operands of the fp32-mode dgrad
  float* dW; float* db;             // accumulate; may be NULL (frozen)
  int n_split;                      // fused Linear: N is n_split blocks with separate dW/db
  float* dW_split[3]; float* db_split[3];
  bool dgrad;
  float* da_f32; long long ld_da;
  const float* da_res; long long ld_res;   // added to da (skip connections / accumulation)
  void* da_b16; long long ld_dab;   // act-dtype copy of da (fp32 mode: THE fp32 destination when da_f32 is NULL)
  const void* gelu_of; long long ld_gelu;  // da *= gelu'(gelu_of)  (bf16: dgrad epilogue; fp32: separate pass)
};

inline const void* col_off(const void* p, long long cols, bool fp32) {
  return static_cast<const char*>(p) + cols * (fp32 ? 4 : 2);
}

// fp32-accurate backward of one Linear: every product is a 3x-bf16 split GEMM on the tensor cores
// (x = x1 + x2 + x3; A' = [x3 x2 x1 x2 x1 x1], B' = [y1 y2 y3 y1 y2 y1] along the contraction dimension,
// small products first; 512 contraction columns per fresh TMEM accumulator, chunk sums added in fp32 RN).
//   wgrad  dW[N,K] (+)= dy^T a : A' = split(dy^T) [N, 6 Mp],  B' = split(a^T)  [K, 6 Mp]   (Mp = M rounded up to 8)
//   dgrad  da[M,K]  =  dy W    : A' = split(dy)   [M, 6 N],   B' = split(W^T)  [K, 6 N]    (from the fp32 masters)
int pack32(BCtx& b, const float* src, long long ld, int R, int Ccols, int transpose, int split, int act, int inner,
           bf16* dst, long long ldd) {
  sea_pack_args a{};
  a.src_f32 = src; a.ld = ld; a.R = R; a.C = Ccols; a.transpose = transpose; a.split = split; a.act = act;
  a.split_inner = inner; a.dst = dst; a.ld_dst = ldd;
  ++g_launches;
  ProfScope prof(b.c.s, SEA_PROF_ELEMWISE, 16.0 * R * Ccols);
  return sea_pack_operand(&a, b.st());
}

int gemm_split(BCtx& b, const bf16* A, long long lda, const bf16* Bm, long long ldb, int M, int N, int K6,
               float* out, long long ldo, const float* res, long long ldres) {
  sea_gemm_problem p{};
  p.a = A; p.lda = lda; p.b = Bm; p.ldb = ldb;
  p.epi.out_f32 = out; p.epi.ld_out_f32 = ldo;
  p.epi.residual = res; p.epi.ld_residual = ldres;
  ++g_launches;
  ProfScope prof(b.c.s, SEA_PROF_GEMM, 2.0 * M * static_cast<double>(N) * K6);
  return sea_gemm_bf16_tn_chunked(1, &p, M, N, K6, 512, b.st());
}

int linear_bwd_fp32(BCtx& b, int n, LinB* L) {
  const int M = b.c.M, Mp = (M + 7) & ~7;
  const int N = L[0].W->N, K = L[0].W->K;
  bf16* P1 = b.bt->pack1; bf16* P2 = b.bt->pack2;
  for (int g = 0; g < n; ++g) {
    const float* dy = static_cast<const float*>(L[g].dy);
    const float* a = static_cast<const float*>(L[g].a);
    const int parts = L[g].n_split > 0 ? L[g].n_split : 1;
    const int Np = N / parts;
    bool any_w = false;
    for (int part = 0; part < parts; ++part) any_w |= (parts > 1 ? L[g].dW_split[part] : L[g].dW) != nullptr;
    if (any_w) {
      if (Mp != M) {   // the padded contraction columns must read as zero
        SEA_CUDA_OK(cudaMemsetAsync(P1, 0, sizeof(bf16) * 6ull * Mp * N, b.c.s));
        SEA_CUDA_OK(cudaMemsetAsync(P2, 0, sizeof(bf16) * 6ull * Mp * K, b.c.s));
      }
      SEA_TRY(pack32(b, a, L[g].lda, M, K, 1, 2, L[g].a_act, Mp, P2, 6LL * Mp));          // a^T  -> [K, 6 Mp]
      SEA_TRY(pack32(b, dy, L[g].lddy, M, N, 1, 1, 0, Mp, P1, 6LL * Mp));                  // dy^T -> [N, 6 Mp]
      for (int part = 0; part < parts; ++part) {
        float* dW = parts > 1 ? L[g].dW_split[part] : L[g].dW;
        if (!dW) continue;
        bool first_touch = b.fresh;
        for (const float* t : b.touched) first_touch = first_touch && t != dW;
        if (first_touch) b.touched.push_back(dW);
        SEA_TRY(gemm_split(b, P1 + static_cast<long long>(part) * Np * 6LL * Mp, 6LL * Mp, P2, 6LL * Mp, Np, K, 6 * Mp,
                           dW, K, first_touch ? nullptr : dW, K));
      }
    }
    for (int part = 0; part < parts; ++part) {
      float* db = parts > 1 ? L[g].db_split[part] : L[g].db;
      const float* src = dy + static_cast<long long>(part) * Np;
      SEA_TRY(colsum_group(b, 1, &src, nullptr, L[g].lddy, M, Np, &db));
    }
    if (L[g].dgrad) {
      float* da = L[g].da_f32 ? L[g].da_f32 : static_cast<float*>(L[g].da_b16);
      const long long ldda = L[g].da_f32 ? L[g].ld_da : L[g].ld_dab;
      if (!da) return SEA_ERR_INVALID;
      SEA_TRY(pack32(b, dy, L[g].lddy, M, N, 0, 1, 0, N, P1, 6LL * N));                    // dy   -> [M, 6 N]
      for (int part = 0; part < parts; ++part) {                                          // W^T  -> [K, 6 N]
        if (!L[g].Wm[part]) return SEA_ERR_INVALID;
        SEA_TRY(pack32(b, L[g].Wm[part], K, Np, K, 1, 2, 0, N, P2 + static_cast<long long>(part) * Np, 6LL * N));
      }
      SEA_TRY(gemm_split(b, P1, 6LL * N, P2, 6LL * N, M, K, 6 * N, da, ldda, L[g].da_res, L[g].ld_res));
      if (L[g].gelu_of) {
        ++g_launches;
        SEA_TRY(sea_gelu_grad_mul_f32(da, ldda, static_cast<const float*>(L[g].gelu_of), L[g].ld_gelu, M, K, b.st()));
      }
    }
  }
  return SEA_OK;
}

int linear_bwd(BCtx& b, int n, LinB* L) {
  if (b.fp32) return linear_bwd_fp32(b, n, L);
  const int M = b.c.M;
  const int N = L[0].W->N, K = L[0].W->K;
  sea_gemm_problem probs[SEA_MAX_STREAMS];
  // ---- wgrad + dbias
  bool any_w = false;
  for (int g = 0; g < n; ++g) any_w |= (L[g].dW != nullptr) || (L[g].n_split > 0 && L[g].dW_split[0] != nullptr);
  if (any_w) {
    // dW[N,K] += dY[M,N]^T X[M,K]: both operands are read as they lie in memory (MN-major UMMA
    // descriptors) — no transpose pass.
    const int parts = L[0].n_split > 0 ? L[0].n_split : 1;
    const int Np = N / parts;
    for (int part = 0; part < parts; ++part) {
      for (int g = 0; g < n; ++g) {
        sea_gemm_problem& p = probs[g];
        p = sea_gemm_problem{};
        p.a = col_off(L[g].dy, static_cast<long long>(part) * Np, false); p.lda = L[g].lddy;
        p.b = L[g].a; p.ldb = L[g].lda;
        p.mn_major = SEA_GEMM_A_MN | SEA_GEMM_B_MN;
        float* dW = parts > 1 ? L[g].dW_split[part] : L[g].dW;
        p.epi.out_f32 = dW; p.epi.ld_out_f32 = K;
        // fresh gradients: the first contribution to a weight overwrites (no zero-fill, no read-back)
        bool first_touch = b.fresh;
        for (const float* t : b.touched) first_touch = first_touch && t != dW;
        if (first_touch) { b.touched.push_back(dW); }
        else { p.epi.residual = dW; p.epi.ld_residual = K; }
        // data-parallel bf16 bucket: the value this launch leaves in dW, rounded, at the same offset of the twin
        // (the last contribution to a weight writes its final value)
        if (bf16* tw = b.twin_of(dW)) { p.epi.out_pre_bf16 = tw; p.epi.ld_out_pre_bf16 = K; }
      }
      SEA_TRY(gemm(b, n, probs, Np, K, M));
      const void* src[SEA_MAX_STREAMS]; float* dbs[SEA_MAX_STREAMS];
      for (int g = 0; g < n; ++g) {
        src[g] = col_off(L[g].dy, static_cast<long long>(part) * Np, false);
        dbs[g] = parts > 1 ? L[g].db_split[part] : L[g].db;
      }
      SEA_TRY(colsum_group(b, n, nullptr, src, L[0].lddy, M, Np, dbs));
    }
  }
  // ---- dgrad
  if (L[0].dgrad) {
    for (int g = 0; g < n; ++g) {
      sea_gemm_problem& p = probs[g];
      p = sea_gemm_problem{};
      p.a = L[g].dy; p.lda = L[g].lddy;
      p.b = L[g].W->w; p.ldb = L[g].W->ldw;   // W [N_out, K_in] read as the MN-major B operand: no transposed copy
      p.mn_major = SEA_GEMM_B_MN;
      p.b_is_static = 1;  // packed weights are older than this call's launch fence
      p.epi.residual = L[g].da_res; p.epi.ld_residual = L[g].ld_res;
      p.epi.gelu_grad_of = L[g].gelu_of; p.epi.ld_gelu = L[g].ld_gelu;
      p.epi.out_f32 = L[g].da_f32; p.epi.ld_out_f32 = L[g].ld_da;
      p.epi.out_pre_bf16 = L[g].da_b16; p.epi.ld_out_pre_bf16 = L[g].ld_dab;
    }
    SEA_TRY(gemm(b, n, probs, M, K, N));
  }
  return SEA_OK;
}

sea_norm_bwd_args norm_bwd_args(BCtx& b, int kind, const sea_norm_params& np, const float* cond, const float* dy,
                                long long lddy, const float* x, long long ldx, const float* stats, int dim,
                                const float* dres, long long lddres, float* dx, long long lddx, bf16* dxb,
                                float* dcond, int dcond_acc) {
  sea_norm_bwd_args a{};
  a.dy = dy; a.lddy = lddy; a.x = x; a.ldx = ldx; a.stats = stats;
  a.M = b.c.M; a.d = dim; a.kind = kind;
  a.weight = np.weight.p; a.cond = cond; a.ldc = 2LL * dim;
  a.dres = dres; a.lddres = lddres;
  a.dx = dx; a.lddx = lddx; a.dx_bf16 = dxb; a.lddx_bf16 = dim;
  a.dweight = np.weight.g;
  a.dbias = kind == SEA_NORM_ADALN ? np.bias.g : nullptr;
  a.dcond = kind == SEA_NORM_ADALN ? dcond : nullptr; a.lddcond = 2LL * dim; a.dcond_accumulate = dcond_acc;
  return a;
}

// E-wide AdaLN norms (ln0, ln2, final): the cond gradient goes straight to the bf16 operand buffer of the
// cond_mlp[2] backward GEMMs, and cond_mlp[2].bias.grad (= the column sums of dcond = this norm's d(weight) | d(bias))
// comes out of the norm kernel's own column reduction: no fp32 dcond, no cast pass, no column-sum pass.
void dcond_direct(BCtx& b, sea_norm_bwd_args& a, const sea_norm_params& np, int g) {
  if (a.kind != SEA_NORM_ADALN || b.fp32) return;   // fp32 mode keeps the fp32 dcond (operand of the split GEMMs)
  a.dcond = nullptr;
  a.dcond_bf16 = b.bt->dcb[g];
  if (np.c2_b.g) { a.dweight2 = np.c2_b.g; a.dbias2 = np.c2_b.g + a.d; }
}

int norm_bwd_group(BCtx& b, int n, const sea_norm_bwd_args* a) {
  ++g_launches;
  ProfScope prof(b.c.s, SEA_PROF_ELEMWISE, 16.0 * b.c.M * a[0].d * n);
  return sea_norm_bwd_group(n, a, b.st());
}

int norm_bwd(BCtx& b, int kind, const sea_norm_params& np, const float* cond, const float* dy,
             long long lddy, const float* x, long long ldx, const float* stats, int dim,
             const float* dres, long long lddres, float* dx, long long lddx, bf16* dxb, float* dcond,
             int dcond_acc) {
  sea_norm_bwd_args a{};
  a.dy = dy; a.lddy = lddy; a.x = x; a.ldx = ldx; a.stats = stats;
  a.M = b.c.M; a.d = dim; a.kind = kind;
  a.weight = np.weight.p; a.cond = cond; a.ldc = 2LL * dim;
  a.dres = dres; a.lddres = lddres;
  a.dx = dx; a.lddx = lddx; a.dx_bf16 = dxb; a.lddx_bf16 = dim;
  a.dweight = np.weight.g;
  a.dbias = kind == SEA_NORM_ADALN ? np.bias.g : nullptr;
  a.dcond = kind == SEA_NORM_ADALN ? dcond : nullptr; a.lddcond = 2LL * dim; a.dcond_accumulate = dcond_acc;
  ++g_launches;
  ProfScope prof(b.c.s, SEA_PROF_ELEMWISE, 16.0 * b.c.M * dim);
  return sea_norm_bwd(&a, b.st());
}

// AdaLN cond_mlp backward for n modules of width d2 = 2*dim (grouped across streams).
int cond_bwd(BCtx& b, int n, const sea_norm_params* const* np, void* const* hid, float* const* dcond,
             const PackedLinear* const* W, int d2, bool direct = false) {
  const int M = b.c.M;
  LinB L[SEA_MAX_STREAMS];
  if (b.fp32) direct = false;
  for (int g = 0; g < n; ++g) {
    if (!direct && !b.fp32) SEA_TRY(cast_f32(b, dcond[g], d2, M, d2, b.bt->dcb[g], d2, 0));
    LinB& l = L[g];
    l = LinB{};
    l.dy = b.op(dcond[g], b.bt->dcb[g]); l.lddy = d2;
    l.a = hid[g]; l.lda = d2;
    l.W = W[g]; l.Wm[0] = np[g]->c2_w.p;
    l.dW = np[g]->c2_w.g; l.db = nullptr;  // bias from the fp32 dcond below
    l.dgrad = true;
    l.da_f32 = b.bt->dhid[g]; l.ld_da = d2;
  }
  SEA_TRY(linear_bwd(b, n, L));
  if (!direct) {
    const float* src[SEA_MAX_STREAMS]; float* dbs[SEA_MAX_STREAMS];
    for (int g = 0; g < n; ++g) { src[g] = dcond[g]; dbs[g] = np[g]->c2_b.g; }
    SEA_TRY(colsum_group(b, n, src, nullptr, d2, M, d2, dbs));
  }
  for (int g = 0; g < n; ++g) {
    if (np[g]->c0_w.g && np[g]->c0_b.g) {
      ++g_launches;
      SEA_TRY(sea_adaln_hidden_bwd(b.bt->dhid[g], d2, b.ib, M, b.c.d->ib_num, np[g]->c0_w.p, np[g]->c0_b.p,
                                   d2, np[g]->c0_w.g, np[g]->c0_b.g, b.st()));
    }
  }
  return SEA_OK;
}

int attention_bwd(BCtx& b, const void* q, long long ldq, const void* k, const void* v, long long ldkv,
                  const void* o, const void* d_o, long long ldo, const float* lse, void* dq, long long lddq,
                  void* dk, void* dv, long long lddkv, int hd, const float* rope, unsigned site) {
  sea_attn_bwd_args a{};
  a.dropout_p = b.c.drop_p; a.dropout_site = site; a.dropout_seed = b.c.d->dropout_seed;
  a.q = q; a.k = k; a.v = v; a.o = o; a.d_o = d_o;
  a.ldq = ldq; a.ldk = ldkv; a.ldv = ldkv; a.ldo = ldo; a.lddo = ldo;
  a.lse = lse; a.delta = b.bt->delta;
  a.dq = dq; a.dk = dk; a.dv = dv; a.lddq = lddq; a.lddk = lddkv; a.lddv = lddkv;
  a.B = b.c.B; a.T = b.c.T; a.n_heads = b.c.d->n_heads; a.head_dim = hd; a.src_len = b.c.d->src_len;
  a.scale = 1.0f / sqrtf(static_cast<float>(hd));
  a.prec = b.fp32 ? SEA_PREC_FP32 : SEA_PREC_BF16;
  a.rope_table = rope; a.rope_ld = b.c.d->max_len;
  g_launches += 3;
  ProfScope prof(b.c.s, SEA_PROF_ATTN, 5.0 * b.c.B * b.c.d->n_heads * static_cast<double>(b.c.T) * b.c.T * hd);
  return sea_attention_bwd(&a, b.st());
}

}  // namespace

void layout_bwd_tape(const sea_temporal_desc* d, int B, int T, Arena& ar, BwdTape& t) {
  const size_t M = static_cast<size_t>(B) * T;
  const size_t E = d->embed_dim, Dd = d->down_dim, H = d->hidden_dim;
  const int V = d->num_streams;
  const bool ada = d->norm_kind == SEA_NORM_ADALN;
  const bool fp32 = d->precision == SEA_PREC_FP32;
  auto f32 = [&](size_t n) { return static_cast<float*>(ar.take(n * 4)); };
  // bf16 twin of an fp32 gradient (the tensor-core operand); fp32 mode reads the fp32 buffer itself
  auto b16 = [&](size_t n) { return fp32 ? nullptr : static_cast<bf16*>(ar.take(n * 2)); };
  // gradients that exist in ONE copy, in the activation dtype of the mode
  auto act = [&](size_t n) { return ar.take(n * (fp32 ? 4 : 2)); };
  for (int i = 0; i < V; ++i) {
    BwdStream& s = t.s[i];
    s.dxout = f32(M * E); s.dxoutb = b16(M * E);
    s.dx3 = f32(M * E); s.dx3b = b16(M * E);
    s.dg = act(M * H); s.dh = act(M * H);
    s.dn2 = f32(M * E);
    s.dx2 = f32(M * E); s.dx2b = b16(M * E);
    s.dxp = f32(M * E); s.dxpb = b16(M * E);
    s.dp = act(M * Dd); s.da = act(M * Dd); s.dq = act(M * Dd); s.dkv = act(M * 2 * Dd);
    s.dnpre = f32(M * Dd); s.dnpost = f32(M * Dd);
    s.ddn = f32(M * Dd); s.ddnb = b16(M * Dd);
    s.dx1 = f32(M * E); s.dx1b = b16(M * E);
    s.dao = act(M * E); s.dqkv = act(M * 3 * E);
    s.dn0 = f32(M * E);
    if (ada) {
      s.dcond0 = f32(M * 2 * E); s.dcond2 = f32(M * 2 * E); s.dcondc = f32(M * 2 * Dd);
      s.dcondF = f32(M * 2 * E);
    }
    if (ada) { t.dcb[i] = b16(M * 2 * E); t.dhid[i] = f32(M * 2 * E); }
  }
  t.delta = f32(static_cast<size_t>(B) * d->n_heads * T);
  t.pack1 = t.pack2 = nullptr;
  if (fp32) {
    // split-operand scratch of linear_bwd_fp32: pack1 = max(N x 6Mp, M x 6N), pack2 = max(K x 6Mp, K x 6N) over
    // every Linear of the model (widest: the stream MLP, and AdaLN's [2E, 2E] cond_mlp[2])
    const size_t Mp = (M + 7) & ~static_cast<size_t>(7);
    size_t wide = H > 3 * E ? H : 3 * E;
    if (ada && 2 * E > wide) wide = 2 * E;
    size_t kn = H * E;
    if (ada && 4 * E * E > kn) kn = 4 * E * E;
    if (3 * E * E > kn) kn = 3 * E * E;
    const size_t p2 = wide * Mp > kn ? wide * Mp : kn;
    t.pack1 = static_cast<bf16*>(ar.take(6 * wide * Mp * 2));
    t.pack2 = static_cast<bf16*>(ar.take(6 * p2 * 2));
  }
}

}  // namespace sea

using namespace sea;

extern "C" int sea_temporal_backward(const sea_temporal_desc* d, const void* cache, const float* x,
                                     const float* ib, const float* dy, float* dx, int B, int T,
                                     void* workspace, size_t workspace_bytes, sea_stream_t stream) {
  if (!d || !d->blocks || !cache || !x || !ib || !dy || !workspace || B <= 0 || T <= 0) return SEA_ERR_INVALID;
  const bool fp32 = d->precision == SEA_PREC_FP32;
  if (fp32 && d->dropout_p > 0.f) return SEA_ERR_UNSUPPORTED;   // train-mode dropout runs in the bf16 mode only
  if (d->ib_hidden > 8 || d->ib_num > 4) return SEA_ERR_UNSUPPORTED;
  if (workspace_bytes < sea_temporal_workspace_bytes(d, B, T, 1)) return SEA_ERR_WORKSPACE;
  SEA_TRY(ensure_init());
  g_launches = 0;
  pdl_fence_next();

  Arena car{const_cast<char*>(static_cast<const char*>(cache))};
  CacheLayout cl;
  layout_cache(d, true, car, cl);
  SplitKGuard splitk_guard;
  SEA_TRY(sea_gemm_set_workspace(cl.splitk[0], cl.splitk_bytes));
  Arena war{static_cast<char*>(workspace)};
  Tape tape;
  layout_tape(d, B, T, true, war, tape);
  BwdTape bt;
  layout_bwd_tape(d, B, T, war, bt);

  BCtx b{};
  b.c.d = d; b.c.cache = &cl; b.c.tape = &tape;
  b.c.s = reinterpret_cast<cudaStream_t>(stream);
  b.c.fp32 = fp32; b.fp32 = fp32; b.c.B = B; b.c.T = T; b.c.M = B * T;
  b.c.Mc = b.c.M; b.c.ld_ib = d->ib_num; b.c.cond_div = 1;
  b.ib = ib; b.bt = &bt;
  b.fresh = d->grads_fresh != 0;
  if (d->grad_bf16 != nullptr && !fp32) { b.gb16 = static_cast<bf16*>(d->grad_bf16); b.gf32 = d->grad_f32_base; }
  if (d->grad_bf16 != nullptr && d->grad_f32_base == nullptr) return SEA_ERR_INVALID;
  if (d->dropout_p < 0.f || d->dropout_p >= 1.f) return SEA_ERR_INVALID;
  b.c.drop_p = d->dropout_p;
  const int M = b.c.M, V = d->num_streams, E = d->embed_dim, Dd = d->down_dim, H = d->hidden_dim;
  const int hd = E / d->n_heads, hdc = Dd / d->n_heads;
  const int kind = d->norm_kind;
  const bool ada = kind == SEA_NORM_ADALN;
  const long long ldY = static_cast<long long>(V) * E;
  LinB L[SEA_MAX_STREAMS];

  // ---- final norm --------------------------------------------------------------------------
  {
    const LayerTape& last = tape.L[d->num_layers - 1];
    sea_norm_bwd_args na[SEA_MAX_STREAMS];
    for (int i = 0; i < V; ++i)
      na[i] = norm_bwd_args(b, kind, d->final_ln[i], tape.condF[i], dy + static_cast<long long>(i) * E, ldY,
                            last.s[i].xout, E, tape.stF[i], E, nullptr, 0, bt.s[i].dxout, E, bt.s[i].dxoutb,
                            bt.s[i].dcondF, 0);
    for (int i = 0; i < V; ++i) dcond_direct(b, na[i], d->final_ln[i], i);
    SEA_TRY(norm_bwd_group(b, V, na));
    if (ada) {
      const sea_norm_params* np[SEA_MAX_STREAMS]; void* hid[SEA_MAX_STREAMS];
      float* dc[SEA_MAX_STREAMS]; const PackedLinear* W[SEA_MAX_STREAMS];
      for (int i = 0; i < V; ++i) { np[i] = &d->final_ln[i]; hid[i] = tape.hidF[i]; dc[i] = bt.s[i].dcondF; W[i] = &cl.c2_final[i]; }
      SEA_TRY(cond_bwd(b, V, np, hid, dc, W, 2 * E, true));
    }
    SEA_CUDA_OK(b.mark(0));
  }

  for (int l = d->num_layers - 1; l >= 0; --l) {
    const sea_block_params& bp = d->blocks[l];
    LayerTape& lt = tape.L[l];
    const BlockCache& bc = cl.blocks[l];
    const float* xin[SEA_MAX_STREAMS];
    long long ldxin;
    if (l == 0) { for (int i = 0; i < V; ++i) xin[i] = x + static_cast<long long>(i) * E; ldxin = ldY; }
    else { for (int i = 0; i < V; ++i) xin[i] = tape.L[l - 1].s[i].xout; ldxin = E; }

    // (5) proj
    for (int i = 0; i < V; ++i) {
      LinB& q = L[i]; q = LinB{};
      q.dy = b.op(bt.s[i].dxout, bt.s[i].dxoutb); q.lddy = E; q.a = lt.s[i].x3; q.lda = E;
      q.W = &bc.s[i].proj; q.Wm[0] = bp.s[i].proj_w.p; q.dW = bp.s[i].proj_w.g; q.db = bp.s[i].proj_b.g;
      q.dgrad = true; q.da_f32 = bt.s[i].dx3; q.ld_da = E; q.da_b16 = bt.s[i].dx3b; q.ld_dab = E;
    }
    SEA_TRY(linear_bwd(b, V, L));
    if (b.c.drop_p > 0.f) {
      // x3 = x2 + Drop(mlp(n2)): the gradient entering the MLP branch carries the forward's mask; the
      // skip path keeps dx3 as it is (fp32), so only the bf16 operand of the mlp3 backward is rewritten
      for (int i = 0; i < V; ++i) {
        ++g_launches;
        SEA_TRY(sea_dropout_apply(bt.s[i].dx3, E, M, E, d->dropout_seed, drop_site(l, SEA_SITE_MLP, i, 0), b.c.drop_p,
                                  nullptr, 0, bt.s[i].dx3b, E, b.st()));
      }
    }
    // (4) MLP: Linear2, LN+GELU, Linear1
    for (int i = 0; i < V; ++i) {
      LinB& q = L[i]; q = LinB{};
      q.dy = b.op(bt.s[i].dx3, bt.s[i].dx3b); q.lddy = E; q.a = lt.s[i].gh; q.lda = H;
      q.W = &bc.s[i].mlp3; q.Wm[0] = bp.s[i].mlp3_w.p; q.dW = bp.s[i].mlp3_w.g; q.db = bp.s[i].mlp3_b.g;
      q.dgrad = true; q.da_b16 = bt.s[i].dg; q.ld_dab = H;
    }
    SEA_TRY(linear_bwd(b, V, L));
    {
      sea_ln_gelu_bwd_args la[SEA_MAX_STREAMS];
      for (int i = 0; i < V; ++i) {
        sea_ln_gelu_bwd_args& a = la[i];
        a = sea_ln_gelu_bwd_args{};
        a.dg = bt.s[i].dg; a.lddg = H; a.h = lt.s[i].h; a.ldh = H; a.stats = lt.s[i].stH;
        a.M = M; a.H = H; a.weight = bp.s[i].mlp_ln_w.p; a.bias = bp.s[i].mlp_ln_b.p;
        a.dh = bt.s[i].dh; a.lddh = H; a.dweight = bp.s[i].mlp_ln_w.g; a.dbias = bp.s[i].mlp_ln_b.g;
        a.prec = fp32 ? SEA_PREC_FP32 : SEA_PREC_BF16;
      }
      ++g_launches;
      ProfScope prof(b.c.s, SEA_PROF_ELEMWISE, 6.0 * M * static_cast<double>(H) * V);
      SEA_TRY(sea_ln_gelu_bwd_group(V, la, b.st()));
    }
    for (int i = 0; i < V; ++i) {
      LinB& q = L[i]; q = LinB{};
      q.dy = bt.s[i].dh; q.lddy = H; q.a = lt.s[i].n2; q.lda = E;
      q.W = &bc.s[i].mlp0; q.Wm[0] = bp.s[i].mlp0_w.p; q.dW = bp.s[i].mlp0_w.g; q.db = bp.s[i].mlp0_b.g;
      q.dgrad = true; q.da_f32 = bt.s[i].dn2; q.ld_da = E;
    }
    SEA_TRY(linear_bwd(b, V, L));
    if (l == 0) {
      // every layer's stream-MLP weight gradient (the bulk of the gradient bytes) is final from here on:
      // the data-parallel exchange of that bucket may start while the rest of the backward runs
      SEA_CUDA_OK(b.mark(1));
    }
    // Norm_{i,2} (+ skip) -> gradient at x2 = x_post + TIPI
    {
      sea_norm_bwd_args na[SEA_MAX_STREAMS];
      for (int i = 0; i < V; ++i)
        na[i] = norm_bwd_args(b, kind, bp.s[i].ln2, lt.s[i].cond2, bt.s[i].dn2, E, lt.s[i].x2, E, lt.s[i].st2, E,
                              bt.s[i].dx3, E, bt.s[i].dx2, E, bt.s[i].dx2b, bt.s[i].dcond2, 0);
      for (int i = 0; i < V; ++i) dcond_direct(b, na[i], bp.s[i].ln2, i);
      SEA_TRY(norm_bwd_group(b, V, na));
    }
    if (ada) {
      const sea_norm_params* np[SEA_MAX_STREAMS]; void* hid[SEA_MAX_STREAMS];
      float* dc[SEA_MAX_STREAMS]; const PackedLinear* W[SEA_MAX_STREAMS];
      for (int i = 0; i < V; ++i) { np[i] = &bp.s[i].ln2; hid[i] = lt.s[i].hid2; dc[i] = bt.s[i].dcond2; W[i] = &bc.s[i].c2_ln2; }
      SEA_TRY(cond_bwd(b, V, np, hid, dc, W, 2 * E, true));
    }
    // (3) TIPI (shared module: gradients of all streams add up)
    if (bp.ib3_w.g) {
      sea_tipi_bwd_args a{};
      for (int i = 0; i < V; ++i) a.dx[i] = bt.s[i].dx2;
      if (b.c.drop_p > 0.f) {
        // x2 = x_post + Drop_i(TIPI): each stream's branch gradient carries its own mask (dn2 is free here)
        for (int i = 0; i < V; ++i) {
          ++g_launches;
          SEA_TRY(sea_dropout_apply(bt.s[i].dx2, E, M, E, d->dropout_seed, drop_site(l, SEA_SITE_TIPI, i, 0), b.c.drop_p,
                                    bt.s[i].dn2, E, nullptr, 0, b.st()));
          a.dx[i] = bt.s[i].dn2;
        }
      }
      a.lddx = E; a.n_streams = V; a.M = M; a.E = E; a.hid = d->ib_hidden; a.ib_num = d->ib_num;
      a.g = lt.tipi_g; a.u = lt.tipi_pre; a.stats = lt.tipi_st; a.ib = ib;
      a.w3 = bp.ib3_w.p; a.ln_w = bp.ib_ln_w.p; a.ln_b = bp.ib_ln_b.p;
      a.dw3 = bp.ib3_w.g; a.db3 = bp.ib3_b.g; a.dlnw = bp.ib_ln_w.g; a.dlnb = bp.ib_ln_b.g;
      a.dw0 = bp.ib0_w.g; a.db0 = bp.ib0_b.g;
      g_launches += V + 1;
      SEA_TRY(sea_tipi_bwd(&a, b.st()));
    }
    if (l == 0) SEA_CUDA_OK(b.mark(2));

    // (2) state exchange, reverse order
    bool pre_written[SEA_MAX_STREAMS] = {}, post_written[SEA_MAX_STREAMS] = {};
    const float* dxp[SEA_MAX_STREAMS]; const void* dxpb[SEA_MAX_STREAMS];   // fp32 value / GEMM operand of d(x_post)
    for (int i = V - 1; i >= 0; --i) {
      StreamTape& s = lt.s[i];
      BwdStream& g = bt.s[i];
      if (i < V - 1 && post_written[i]) {
        SEA_TRY(norm_bwd(b, kind, bp.s[i].ln_cross, s.condc, g.dnpost, Dd, s.dpost, Dd, s.stc_post, Dd,
                         nullptr, 0, g.ddn, Dd, g.ddnb, g.dcondc, 0));
        LinB& q = L[0]; q = LinB{};
        q.dy = b.op(g.ddn, g.ddnb); q.lddy = Dd; q.a = b.op(s.xp, s.xpb); q.lda = E;
        q.W = &bc.s[i].down; q.Wm[0] = bp.s[i].down_w.p; q.dW = bp.s[i].down_w.g; q.db = bp.s[i].down_b.g;
        q.dgrad = true; q.da_f32 = g.dxp; q.ld_da = E; q.da_res = g.dx2; q.ld_res = E;
        q.da_b16 = g.dxpb; q.ld_dab = E;
        SEA_TRY(linear_bwd(b, 1, L));
        dxp[i] = g.dxp; dxpb[i] = b.op(g.dxp, g.dxpb);
      } else {
        dxp[i] = g.dx2; dxpb[i] = b.op(g.dx2, g.dx2b);
      }
      for (int j = 0; j < V; ++j) {
        if (j == i) continue;
        // cross_up (+ GELU' fused)
        { LinB& q = L[0]; q = LinB{};
          q.dy = dxpb[i]; q.lddy = E; q.lda = Dd;
          if (fp32) { q.a = s.p[j]; q.a_act = SEA_ACT_GELU; }   // the forward applied GELU while packing (no g copy)
          else q.a = s.g[j];
          q.W = &bc.s[i].up; q.Wm[0] = bp.s[i].up_w.p; q.dW = bp.s[i].up_w.g; q.db = bp.s[i].up_b.g;
          q.dgrad = true; q.da_b16 = g.dp; q.ld_dab = Dd;
          q.gelu_of = s.p[j]; q.ld_gelu = Dd;
          SEA_TRY(linear_bwd(b, 1, L)); }
        // attention output projection
        { LinB& q = L[0]; q = LinB{};
          q.dy = g.dp; q.lddy = Dd; q.a = s.a[j]; q.lda = Dd;
          q.W = &bc.s[i].cproj[j]; q.Wm[0] = bp.s[i].cross_attn[j].proj_w.p; q.dW = bp.s[i].cross_attn[j].proj_w.g;
          q.dgrad = true; q.da_b16 = g.da; q.ld_dab = Dd;
          SEA_TRY(linear_bwd(b, 1, L)); }
        const void* kv = s.kv[j];
        SEA_TRY(attention_bwd(b, s.q[j], Dd, kv, col_off(kv, Dd, fp32), 2 * Dd, s.a[j], g.da, Dd, s.lse_c[j], g.dq, Dd,
                              g.dkv, const_cast<void*>(col_off(g.dkv, Dd, fp32)), 2 * Dd, hdc, d->rope_cross,
                              drop_site(l, SEA_SITE_CROSS, i, j)));
        // q projection (input: ln_cross_i(down_i(x1_i)))
        { LinB& q = L[0]; q = LinB{};
          q.dy = g.dq; q.lddy = Dd; q.a = s.npre; q.lda = Dd;
          q.W = &bc.s[i].cq[j]; q.Wm[0] = bp.s[i].cross_attn[j].q_w.p;
          q.dW = bp.s[i].cross_attn[j].q_w.g; q.db = bp.s[i].cross_attn[j].q_b.g;
          q.dgrad = true; q.da_f32 = g.dnpre; q.ld_da = Dd;
          if (pre_written[i]) { q.da_res = g.dnpre; q.ld_res = Dd; }
          SEA_TRY(linear_bwd(b, 1, L));
          pre_written[i] = true; }
        // fused k|v projection (input: stream j, exchanged if j < i)
        { LinB& q = L[0]; q = LinB{};
          const bool from_post = j < i;
          float* dst = from_post ? bt.s[j].dnpost : bt.s[j].dnpre;
          bool& written = from_post ? post_written[j] : pre_written[j];
          q.dy = g.dkv; q.lddy = 2 * Dd;
          q.a = from_post ? lt.s[j].npost : lt.s[j].npre; q.lda = Dd;
          q.W = &bc.s[i].ckv[j]; q.Wm[0] = bp.s[i].cross_attn[j].k_w.p; q.Wm[1] = bp.s[i].cross_attn[j].v_w.p;
          q.n_split = 2;
          q.dW_split[0] = bp.s[i].cross_attn[j].k_w.g; q.dW_split[1] = bp.s[i].cross_attn[j].v_w.g;
          q.db_split[0] = bp.s[i].cross_attn[j].k_b.g; q.db_split[1] = bp.s[i].cross_attn[j].v_b.g;
          q.dgrad = true; q.da_f32 = dst; q.ld_da = Dd;
          if (written) { q.da_res = dst; q.ld_res = Dd; }
          SEA_TRY(linear_bwd(b, 1, L));
          written = true; }
      }
    }
    // pre-exchange branch of every stream: ln_cross + cross_down on x1 (grouped over the streams)
    if (V == 1) {  // the exchange is the identity
      BwdStream& g = bt.s[0];
      SEA_CUDA_OK(cudaMemcpyAsync(g.dx1, dxp[0], sizeof(float) * M * E, cudaMemcpyDeviceToDevice, b.c.s));
      if (!fp32) SEA_CUDA_OK(cudaMemcpyAsync(g.dx1b, dxpb[0], sizeof(bf16) * M * E, cudaMemcpyDeviceToDevice, b.c.s));
    } else {
      sea_norm_bwd_args na[SEA_MAX_STREAMS];
      for (int i = 0; i < V; ++i) {
        StreamTape& s = lt.s[i];
        BwdStream& g = bt.s[i];
        na[i] = norm_bwd_args(b, kind, bp.s[i].ln_cross, s.condc, g.dnpre, Dd, s.dpre, Dd, s.stc_pre, Dd, nullptr, 0,
                              g.ddn, Dd, g.ddnb, g.dcondc, (i < V - 1 && post_written[i]) ? 1 : 0);
      }
      SEA_TRY(norm_bwd_group(b, V, na));
      for (int i = 0; i < V; ++i) {
        StreamTape& s = lt.s[i];
        BwdStream& g = bt.s[i];
        LinB& q = L[i]; q = LinB{};
        q.dy = b.op(g.ddn, g.ddnb); q.lddy = Dd; q.a = b.op(s.x1, s.x1b); q.lda = E;
        q.W = &bc.s[i].down; q.Wm[0] = bp.s[i].down_w.p; q.dW = bp.s[i].down_w.g; q.db = bp.s[i].down_b.g;
        q.dgrad = true; q.da_f32 = g.dx1; q.ld_da = E; q.da_res = dxp[i]; q.ld_res = E;
        q.da_b16 = g.dx1b; q.ld_dab = E;
      }
      SEA_TRY(linear_bwd(b, V, L));
    }
    if (ada && V > 1) {
      const sea_norm_params* np[SEA_MAX_STREAMS]; void* hid[SEA_MAX_STREAMS];
      float* dc[SEA_MAX_STREAMS]; const PackedLinear* W[SEA_MAX_STREAMS];
      for (int i = 0; i < V; ++i) { np[i] = &bp.s[i].ln_cross; hid[i] = lt.s[i].hidc; dc[i] = bt.s[i].dcondc; W[i] = &bc.s[i].c2_lnc; }
      SEA_TRY(cond_bwd(b, V, np, hid, dc, W, 2 * Dd));
    }

    if (l == 0) SEA_CUDA_OK(b.mark(3));

    // (1) self-attention
    for (int i = 0; i < V; ++i) {
      LinB& q = L[i]; q = LinB{};
      q.dy = b.op(bt.s[i].dx1, bt.s[i].dx1b); q.lddy = E; q.a = lt.s[i].ao; q.lda = E;
      q.W = &bc.s[i].sproj; q.Wm[0] = bp.s[i].self_attn.proj_w.p; q.dW = bp.s[i].self_attn.proj_w.g;
      q.dgrad = true; q.da_b16 = bt.s[i].dao; q.ld_dab = E;
    }
    SEA_TRY(linear_bwd(b, V, L));
    for (int i = 0; i < V; ++i) {
      const void* qkv = lt.s[i].qkv;
      void* dqkv = bt.s[i].dqkv;
      SEA_TRY(attention_bwd(b, qkv, 3 * E, col_off(qkv, E, fp32), col_off(qkv, 2 * E, fp32), 3 * E, lt.s[i].ao,
                            bt.s[i].dao, E, lt.s[i].lse, dqkv, 3 * E, const_cast<void*>(col_off(dqkv, E, fp32)),
                            const_cast<void*>(col_off(dqkv, 2 * E, fp32)), 3 * E, hd,
                            d->rope_self, drop_site(l, SEA_SITE_SELF, i, 0)));
    }
    for (int i = 0; i < V; ++i) {
      LinB& q = L[i]; q = LinB{};
      q.dy = bt.s[i].dqkv; q.lddy = 3 * E; q.a = lt.s[i].n0; q.lda = E;
      q.W = &bc.s[i].qkv; q.n_split = 3;
      q.Wm[0] = bp.s[i].self_attn.q_w.p; q.Wm[1] = bp.s[i].self_attn.k_w.p; q.Wm[2] = bp.s[i].self_attn.v_w.p;
      q.dW_split[0] = bp.s[i].self_attn.q_w.g; q.dW_split[1] = bp.s[i].self_attn.k_w.g; q.dW_split[2] = bp.s[i].self_attn.v_w.g;
      q.db_split[0] = bp.s[i].self_attn.q_b.g; q.db_split[1] = bp.s[i].self_attn.k_b.g; q.db_split[2] = bp.s[i].self_attn.v_b.g;
      q.dgrad = true; q.da_f32 = bt.s[i].dn0; q.ld_da = E;
    }
    SEA_TRY(linear_bwd(b, V, L));
    // Norm_{i,0} (+ skip) -> gradient at the layer input
    {
      sea_norm_bwd_args na[SEA_MAX_STREAMS];
      for (int i = 0; i < V; ++i) {
        float* dst = nullptr; long long ldd = E; bf16* dstb = nullptr;
        if (l > 0) { dst = bt.s[i].dxout; dstb = bt.s[i].dxoutb; }
        else if (dx) { dst = dx + static_cast<long long>(i) * E; ldd = ldY; }
        else { dst = bt.s[i].dxout; }  // not requested: still needed for the parameter gradients
        na[i] = norm_bwd_args(b, kind, bp.s[i].ln0, lt.s[i].cond0, bt.s[i].dn0, E, xin[i], ldxin, lt.s[i].st0, E,
                              bt.s[i].dx1, E, dst, ldd, dstb, bt.s[i].dcond0, 0);
        dcond_direct(b, na[i], bp.s[i].ln0, i);
      }
      SEA_TRY(norm_bwd_group(b, V, na));
    }
    if (ada) {
      const sea_norm_params* np[SEA_MAX_STREAMS]; void* hid[SEA_MAX_STREAMS];
      float* dc[SEA_MAX_STREAMS]; const PackedLinear* W[SEA_MAX_STREAMS];
      for (int i = 0; i < V; ++i) { np[i] = &bp.s[i].ln0; hid[i] = lt.s[i].hid0; dc[i] = bt.s[i].dcond0; W[i] = &bc.s[i].c2_ln0; }
      SEA_TRY(cond_bwd(b, V, np, hid, dc, W, 2 * E, true));
    }
  }
  SEA_CUDA_OK(b.mark(4));
  return SEA_OK;
}

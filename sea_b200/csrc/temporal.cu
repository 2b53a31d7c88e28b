// Whole-model executor for the temporal transformer (host-side orchestration, C++).
//
// One call = the complete TemporalModel.forward of the reference (models/temporal.py:405-416):
// every kernel of the pass is enqueued on the caller's stream from here, so the Python side
// makes ONE FFI call per forward (and the sequence is CUDA-graph capturable: no allocation, no
// host sync, descriptors built on the host and passed by value).
//
// Data layout in HBM
//   weights   fp32 masters stay where torch put them; bf16 (or 3-way-split bf16) packed copies,
//             with q|k|v and k|v fused along N, live in the caller-owned `cache`;
//   residual  stream x_i is fp32 [M, E] (M = B*T) throughout (bf16 budget: SURVEY.md hard part 2);
//   operands  every GEMM A-operand is written in bf16 by the kernel that produces it
//             (norm / attention / GEMM epilogue), never re-read in fp32;
//   tape      all intermediates sit at fixed offsets of the caller-owned `workspace`, which is
//             also what the backward pass reads.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <vector>

#include "../../include/sea_b200.h"
#include "internal.h"
#include "temporal_internal.h"

namespace sea {

thread_local int g_launches = 0;

// ---------------------------------------------------------------------------------- profiler
// Optional per-launch CUDA-event timing on the launching stream (bench.py's roofline leg).
namespace {
struct ProfRec { cudaEvent_t a, b; int cat; double work; };
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
std::vector<cudaEvent_t> g_event_pool;
cudaEvent_t prof_event() {
  if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
}  // namespace

ProfScope::ProfScope(cudaStream_t s, int cat, double work) : s_(s), idx_(-1) {
  if (!g_prof_on) return;
  ProfRec r{prof_event(), prof_event(), cat, work};
  cudaEventRecord(r.a, s_);
  g_prof.push_back(r);
  idx_ = static_cast<int>(g_prof.size()) - 1;
}
ProfScope::~ProfScope() {
  if (idx_ >= 0) cudaEventRecord(g_prof[idx_].b, s_);
}

// ------------------------------------------------------------------------------ cache layout
static PackedLinear take_linear(Arena& ar, int N, int K, int kf, bool training) {
  PackedLinear p{};
  (void)training;  // one layout for inference and training: dgrad needs no transposed copies
  p.N = N; p.K = K;
  p.ldw = static_cast<long long>(kf) * K;
  p.w = static_cast<bf16*>(ar.take(sizeof(bf16) * N * p.ldw));
  return p;
}

void layout_cache(const sea_temporal_desc* d, bool training, Arena& ar, CacheLayout& c) {
  const int kf = d->precision == SEA_PREC_FP32 ? 6 : 1;
  const int E = d->embed_dim, Dd = d->down_dim, H = d->hidden_dim, V = d->num_streams;
  const bool ada = d->norm_kind == SEA_NORM_ADALN;
  c.blocks.assign(d->num_layers, BlockCache{});
  for (int l = 0; l < d->num_layers; ++l) {
    for (int i = 0; i < V; ++i) {
      StreamCache& s = c.blocks[l].s[i];
      s.qkv = take_linear(ar, 3 * E, E, kf, training);
      s.qkv_bias = static_cast<float*>(ar.take(sizeof(float) * 3 * E));
      s.sproj = take_linear(ar, E, E, kf, training);
      s.down = take_linear(ar, Dd, E, kf, training);
      s.up = take_linear(ar, E, Dd, kf, training);
      s.mlp0 = take_linear(ar, H, E, kf, training);
      s.mlp3 = take_linear(ar, E, H, kf, training);
      s.proj = take_linear(ar, E, E, kf, training);
      for (int j = 0; j < V; ++j) {
        if (j == i) continue;
        s.cq[j] = take_linear(ar, Dd, Dd, kf, training);
        s.ckv[j] = take_linear(ar, 2 * Dd, Dd, kf, training);
        s.ckv_bias[j] = static_cast<float*>(ar.take(sizeof(float) * 2 * Dd));
        s.cproj[j] = take_linear(ar, Dd, Dd, kf, training);
      }
      if (ada) {
        s.c2_ln0 = take_linear(ar, 2 * E, 2 * E, kf, training);
        s.c2_ln2 = take_linear(ar, 2 * E, 2 * E, kf, training);
        s.c2_lnc = take_linear(ar, 2 * Dd, 2 * Dd, kf, training);
      }
    }
  }
  for (int i = 0; i < V; ++i)
    if (ada) c.c2_final[i] = take_linear(ar, 2 * E, 2 * E, kf, training);
  // stream-K workspace of the GEMM launches (arrival counters + parked partial tiles), see sea_gemm_set_workspace
  c.splitk_bytes = kSplitKBytes;
  for (int k = 0; k < 4; ++k) c.splitk[k] = ar.take(c.splitk_bytes);
}

static int pack_weight(const float* src, int N, int K, const PackedLinear& dst, int row_off,
                       int n_total, bool fp32, int what, cudaStream_t s) {
  // rows [row_off, row_off+N) of the (possibly fused) packed matrix
  sea_pack_args a{};
  a.src_f32 = src; a.ld = K; a.R = N; a.C = K;
  a.split = fp32 ? 2 : 0;
  int rc = SEA_OK;
  if (what & SEA_REFRESH_STRAIGHT) {
    a.dst = const_cast<bf16*>(dst.w) + static_cast<long long>(row_off) * dst.ldw;
    a.ld_dst = dst.ldw;
    rc = sea_pack_operand(&a, reinterpret_cast<sea_stream_t>(s));
    if (rc) return rc;
    ++g_launches;
  }
  (void)n_total;
  return rc;
}

#define SEA_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != SEA_OK) return _rc; \
  } while (0)

// Every (fp32 master -> rows of a packed matrix) pair of the model, in one place: the refresh
// walks it to pack, sea_temporal_cache_slot walks it to answer "where does this master's copy live".
template <typename F>
static int for_each_weight(const sea_temporal_desc* d, CacheLayout& c, F&& f) {
  const int E = d->embed_dim, Dd = d->down_dim, H = d->hidden_dim, V = d->num_streams;
  const bool ada = d->norm_kind == SEA_NORM_ADALN;
  for (int l = 0; l < d->num_layers; ++l) {
    for (int i = 0; i < V; ++i) {
      const sea_stream_params& p = d->blocks[l].s[i];
      StreamCache& sc = c.blocks[l].s[i];
      SEA_TRY(f(p.self_attn.q_w.p, E, E, sc.qkv, 0, 3 * E));
      SEA_TRY(f(p.self_attn.k_w.p, E, E, sc.qkv, E, 3 * E));
      SEA_TRY(f(p.self_attn.v_w.p, E, E, sc.qkv, 2 * E, 3 * E));
      SEA_TRY(f(p.self_attn.proj_w.p, E, E, sc.sproj, 0, E));
      SEA_TRY(f(p.down_w.p, Dd, E, sc.down, 0, Dd));
      SEA_TRY(f(p.up_w.p, E, Dd, sc.up, 0, E));
      SEA_TRY(f(p.mlp0_w.p, H, E, sc.mlp0, 0, H));
      SEA_TRY(f(p.mlp3_w.p, E, H, sc.mlp3, 0, E));
      SEA_TRY(f(p.proj_w.p, E, E, sc.proj, 0, E));
      for (int j = 0; j < V; ++j) {
        if (j == i) continue;
        const sea_attn_params& ca = p.cross_attn[j];
        SEA_TRY(f(ca.q_w.p, Dd, Dd, sc.cq[j], 0, Dd));
        SEA_TRY(f(ca.k_w.p, Dd, Dd, sc.ckv[j], 0, 2 * Dd));
        SEA_TRY(f(ca.v_w.p, Dd, Dd, sc.ckv[j], Dd, 2 * Dd));
        SEA_TRY(f(ca.proj_w.p, Dd, Dd, sc.cproj[j], 0, Dd));
      }
      if (ada) {
        SEA_TRY(f(p.ln0.c2_w.p, 2 * E, 2 * E, sc.c2_ln0, 0, 2 * E));
        SEA_TRY(f(p.ln2.c2_w.p, 2 * E, 2 * E, sc.c2_ln2, 0, 2 * E));
        SEA_TRY(f(p.ln_cross.c2_w.p, 2 * Dd, 2 * Dd, sc.c2_lnc, 0, 2 * Dd));
      }
    }
  }
  if (ada)
    for (int i = 0; i < V; ++i) SEA_TRY(f(d->final_ln[i].c2_w.p, 2 * E, 2 * E, c.c2_final[i], 0, 2 * E));
  return SEA_OK;
}

}  // namespace sea

using namespace sea;

extern "C" int sea_last_launch_count(void) { return g_launches; }

extern "C" void sea_profile_begin(void) {
  for (auto& r : g_prof) { g_event_pool.push_back(r.a); g_event_pool.push_back(r.b); }
  g_prof.clear();
  g_prof_on = true;
}

extern "C" int sea_profile_end(sea_profile_summary* out) {
  g_prof_on = false;
  if (!out) return SEA_ERR_INVALID;
  for (int c = 0; c < SEA_PROF_NUM; ++c) { out->ms[c] = 0; out->work[c] = 0; out->launches[c] = 0; }
  for (auto& r : g_prof) {
    cudaError_t e = cudaEventSynchronize(r.b);
    if (e != cudaSuccess) return static_cast<int>(e);
    float ms = 0.f;
    e = cudaEventElapsedTime(&ms, r.a, r.b);
    if (e != cudaSuccess) return static_cast<int>(e);
    out->ms[r.cat] += ms; out->work[r.cat] += r.work; out->launches[r.cat] += 1;
  }
  return SEA_OK;
}

extern "C" size_t sea_temporal_cache_bytes(const sea_temporal_desc* d, int training) {
  if (!d) return 0;
  Arena ar{nullptr};
  CacheLayout c;
  layout_cache(d, training != 0, ar, c);
  return ar.off + 256;
}

static int validate_desc(const sea_temporal_desc* d) {
  if (!d || !d->blocks) return SEA_ERR_INVALID;
  if (d->num_layers < 1 || d->num_streams < 1 || d->num_streams > SEA_MAX_STREAMS) return SEA_ERR_UNSUPPORTED;
  if (d->n_heads < 1 || d->embed_dim % d->n_heads || d->down_dim % d->n_heads) return SEA_ERR_UNSUPPORTED;
  const int hd = d->embed_dim / d->n_heads, hdc = d->down_dim / d->n_heads;
  if (hd % 32 || hdc % 32 || hd > 256) return SEA_ERR_UNSUPPORTED;
  if (d->embed_dim > 2048 || d->hidden_dim > 16384 || d->hidden_dim % 8) return SEA_ERR_UNSUPPORTED;
  if (d->ib_hidden < 1 || d->ib_hidden > 64 || d->ib_num < 1) return SEA_ERR_UNSUPPORTED;
  if (!d->rope_self || !d->rope_cross) return SEA_ERR_INVALID;
  return SEA_OK;
}

extern "C" int sea_temporal_refresh_ex(const sea_temporal_desc* d, void* cache, size_t cache_bytes,
                                       int training, int what, sea_stream_t stream) {
  SEA_TRY(validate_desc(d));
  if (!cache) return SEA_ERR_INVALID;
  if (cache_bytes < sea_temporal_cache_bytes(d, training)) return SEA_ERR_WORKSPACE;
  g_launches = 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  Arena ar{static_cast<char*>(cache)};
  CacheLayout c;
  layout_cache(d, training != 0, ar, c);
  const bool fp32 = d->precision == SEA_PREC_FP32;
  const int E = d->embed_dim, Dd = d->down_dim, V = d->num_streams;
  for (int k = 0; k < 4; ++k) SEA_CUDA_OK(cudaMemsetAsync(c.splitk[k], 0, 65536, s));   // stream-K arrival counters start at zero
  SEA_TRY(for_each_weight(d, c, [&](const float* src, int N, int K, const PackedLinear& dst, int row_off, int n_total) {
    return pack_weight(src, N, K, dst, row_off, n_total, fp32, what, s);
  }));
  if (what & SEA_REFRESH_BIASES) {
    for (int l = 0; l < d->num_layers; ++l) {
      for (int i = 0; i < V; ++i) {
        const sea_stream_params& p = d->blocks[l].s[i];
        StreamCache& sc = c.blocks[l].s[i];
        SEA_CUDA_OK(cudaMemcpyAsync(sc.qkv_bias, p.self_attn.q_b.p, sizeof(float) * E, cudaMemcpyDeviceToDevice, s));
        SEA_CUDA_OK(cudaMemcpyAsync(sc.qkv_bias + E, p.self_attn.k_b.p, sizeof(float) * E, cudaMemcpyDeviceToDevice, s));
        SEA_CUDA_OK(cudaMemcpyAsync(sc.qkv_bias + 2 * E, p.self_attn.v_b.p, sizeof(float) * E, cudaMemcpyDeviceToDevice, s));
        for (int j = 0; j < V; ++j) {
          if (j == i) continue;
          const sea_attn_params& ca = p.cross_attn[j];
          SEA_CUDA_OK(cudaMemcpyAsync(sc.ckv_bias[j], ca.k_b.p, sizeof(float) * Dd, cudaMemcpyDeviceToDevice, s));
          SEA_CUDA_OK(cudaMemcpyAsync(sc.ckv_bias[j] + Dd, ca.v_b.p, sizeof(float) * Dd, cudaMemcpyDeviceToDevice, s));
        }
      }
    }
  }
  return SEA_OK;
}

extern "C" int sea_temporal_refresh(const sea_temporal_desc* d, void* cache, size_t cache_bytes,
                                    int training, sea_stream_t stream) {
  return sea_temporal_refresh_ex(d, cache, cache_bytes, training, SEA_REFRESH_ALL, stream);
}

extern "C" int sea_temporal_cache_slot(const sea_temporal_desc* d, void* cache, int training,
                                       const float* master, void** dst) {
  SEA_TRY(validate_desc(d));
  if (!cache || !master || !dst) return SEA_ERR_INVALID;
  *dst = nullptr;
  if (d->precision != SEA_PREC_BF16) return SEA_OK;  // split copies are not a plain rounding of the master
  Arena ar{static_cast<char*>(cache)};
  CacheLayout c;
  layout_cache(d, training != 0, ar, c);
  return for_each_weight(d, c, [&](const float* src, int, int, const PackedLinear& pl, int row_off, int) {
    if (src == master) *dst = const_cast<bf16*>(pl.w) + static_cast<long long>(row_off) * pl.ldw;
    return static_cast<int>(SEA_OK);
  });
}

// ---------------------------------------------------------------------------------- the tape
namespace sea {

void layout_tape(const sea_temporal_desc* d, int B, int T, bool training, Arena& ar, Tape& t) {
  (void)training;
  const bool fp32 = d->precision == SEA_PREC_FP32;
  const size_t esz = fp32 ? 4 : 2;
  const size_t M = static_cast<size_t>(B) * T;
  const size_t E = d->embed_dim, Dd = d->down_dim, H = d->hidden_dim;
  const int V = d->num_streams;
  const bool ada = d->norm_kind == SEA_NORM_ADALN;
  auto act = [&](size_t elems) { return ar.take(elems * esz); };
  auto f32 = [&](size_t elems) { return static_cast<float*>(ar.take(elems * 4)); };
  auto b16 = [&](size_t elems) { return fp32 ? nullptr : static_cast<bf16*>(ar.take(elems * 2)); };
  t.L.assign(d->num_layers, LayerTape{});
  for (int l = 0; l < d->num_layers; ++l) {
    LayerTape& lt = t.L[l];
    lt.tipi_g = f32(M * d->ib_hidden);
    lt.tipi_pre = f32(M * d->ib_hidden);
    lt.tipi_st = f32(M * 2);
    lt.tipi_rows = f32(static_cast<size_t>(B) * E);
    for (int i = 0; i < V; ++i) {
      StreamTape& s = lt.s[i];
      if (ada) {
        s.hid0 = act(M * 2 * E); s.hid2 = act(M * 2 * E); s.hidc = act(M * 2 * Dd);
        s.cond0 = f32(M * 2 * E); s.cond2 = f32(M * 2 * E); s.condc = f32(M * 2 * Dd);
      }
      s.n0 = act(M * E); s.st0 = f32(M * 2);
      s.qkv = act(M * 3 * E); s.ao = act(M * E);
      s.lse = f32(static_cast<size_t>(B) * d->n_heads * T);
      s.x1 = f32(M * E); s.x1b = b16(M * E);
      s.dpre = f32(M * Dd); s.stc_pre = f32(M * 2); s.npre = act(M * Dd);
      if (i < V - 1) { s.dpost = f32(M * Dd); s.stc_post = f32(M * 2); s.npost = act(M * Dd); }
      for (int j = 0; j < V; ++j) {
        if (j == i) continue;
        s.q[j] = act(M * Dd); s.kv[j] = act(M * 2 * Dd); s.a[j] = act(M * Dd);
        s.lse_c[j] = f32(static_cast<size_t>(B) * d->n_heads * T);
        s.p[j] = act(M * Dd); s.g[j] = b16(M * Dd);
      }
      s.xp = f32(M * E); s.xpb = b16(M * E);
      s.x2 = f32(M * E); s.n2 = act(M * E); s.st2 = f32(M * 2);
      s.h = act(M * H); s.stH = f32(M * 2); s.gh = act(M * H);
      s.x3 = act(M * E);
      s.xout = f32(M * E);
    }
  }
  for (int i = 0; i < V; ++i) {
    if (ada) { t.hidF[i] = act(M * 2 * E); t.condF[i] = f32(M * 2 * E); }
    t.stF[i] = f32(M * 2);
  }
  if (fp32) {
    const size_t kmax = H > 2 * E ? H : 2 * E;
    for (int g = 0; g < SEA_MAX_STREAMS; ++g) t.packA[g] = static_cast<bf16*>(ar.take(M * 6 * kmax * 2));
  }
}

// Outputs of the condition path that later kernels read, placed in a persistent caller-owned
// buffer (one row per trajectory) so a rollout can reuse them across forward calls.
void layout_cond_cache(const sea_temporal_desc* d, int B, Arena& ar, Tape& t) {
  const size_t E = d->embed_dim, Dd = d->down_dim, R = static_cast<size_t>(B);
  const bool ada = d->norm_kind == SEA_NORM_ADALN;
  auto f32 = [&](size_t n) { return static_cast<float*>(ar.take(n * 4)); };
  for (int l = 0; l < d->num_layers; ++l) {
    float* rows = f32(R * E);
    if (!t.L.empty()) t.L[l].tipi_rows = rows;
    for (int i = 0; i < d->num_streams; ++i) {
      if (!ada) continue;
      float* c0 = f32(R * 2 * E); float* c2 = f32(R * 2 * E); float* cc = f32(R * 2 * Dd);
      if (!t.L.empty()) { t.L[l].s[i].cond0 = c0; t.L[l].s[i].cond2 = c2; t.L[l].s[i].condc = cc; }
    }
  }
  for (int i = 0; i < d->num_streams; ++i)
    if (ada) { float* cf = f32(R * 2 * E); t.condF[i] = cf; }
}

// ------------------------------------------------------------------------------ op helpers
int linear_group(Ctx& c, int n, const LinIn* in, const PackedLinear* const* W, const LinOut* out,
                 int Mrows) {
  sea_gemm_problem probs[SEA_MAX_STREAMS];
  const int N = W[0]->N, K = W[0]->K;
  for (int g = 0; g < n; ++g) {
    sea_gemm_problem& p = probs[g];
    p = sea_gemm_problem{};
    const LinOut& o = out[g];
    if (c.fp32) {
      sea_pack_args pa{};
      pa.src_f32 = static_cast<const float*>(in[g].a);
      pa.ld = in[g].lda; pa.R = Mrows; pa.C = K; pa.split = 1; pa.act = in[g].act_on_load;
      pa.dst = c.tape->packA[g]; pa.ld_dst = 6LL * K;
      {
        ProfScope prof(c.s, SEA_PROF_ELEMWISE, 4.0 * Mrows * K + 12.0 * Mrows * K);
        SEA_TRY(sea_pack_operand(&pa, reinterpret_cast<sea_stream_t>(c.s)));
      }
      ++g_launches;
      p.a = c.tape->packA[g]; p.lda = 6LL * K;
    } else {
      p.a = in[g].a; p.lda = in[g].lda;
    }
    p.b = W[g]->w; p.ldb = W[g]->ldw;
    p.b_is_static = 1;  // packed weights: written before this call's launch fence
    sea_gemm_epilogue& e = p.epi;
    e.bias = o.bias;
    e.residual = o.residual; e.ld_residual = o.ld_res;
    e.res_rows_per_batch = o.res_rows; e.res_batch_stride = o.res_bs;
    e.dropout_p = o.drop_p; e.dropout_site = o.drop_site; e.dropout_seed = c.d->dropout_seed;
    e.act = o.act;
    if (o.rope_cols > 0) {
      e.rope_cols = o.rope_cols; e.head_dim = o.head_dim; e.seq_len = c.T;
      e.rope_table = o.rope_table; e.rope_sign = 1.f; e.rope_ld = c.d->max_len;
      e.rope_pos0 = c.pos0;
    }
    if (c.fp32) {
      // exactly one fp32 destination in the parity mode
      if (o.f32) { e.out_f32 = o.f32; e.ld_out_f32 = o.ld_f32; }
      else if (o.pre) { e.out_f32 = static_cast<float*>(o.pre); e.ld_out_f32 = o.ld_pre; }
      else { e.out_f32 = static_cast<float*>(o.post); e.ld_out_f32 = o.ld_post; }
    } else {
      e.out_f32 = o.f32; e.ld_out_f32 = o.ld_f32;
      e.out_pre_bf16 = o.pre; e.ld_out_pre_bf16 = o.ld_pre;
      e.out_bf16 = o.post; e.ld_out_bf16 = o.ld_post;
    }
  }
  ++g_launches;
  if (c.splitk != nullptr) SEA_TRY(sea_gemm_set_workspace(c.splitk, kSplitKBytes));   // host-side selection, per launch
  ProfScope prof(c.s, SEA_PROF_GEMM, 2.0 * Mrows * static_cast<double>(N) * K * n);
  if (c.fp32) {
    // fresh accumulator every 512 columns of K' so the tensor core's non-RN accumulation cannot drift
    bool can_chunk = true;
    for (int g = 0; g < n; ++g)
      if (probs[g].epi.out_f32 == nullptr || probs[g].epi.out_f32 == probs[g].epi.residual) can_chunk = false;
    return sea_gemm_bf16_tn_chunked(n, probs, Mrows, N, 6 * K, can_chunk ? 512 : 0,
                                    reinterpret_cast<sea_stream_t>(c.s));
  }
  return sea_gemm_bf16_tn(n, probs, Mrows, N, K, reinterpret_cast<sea_stream_t>(c.s));
}

// One row-norm of the model as launch arguments (+ its algorithmic bytes for the profiler).
struct NormCall { sea_norm_args a; double bytes; };

NormCall norm_call(Ctx& c, int kind, const sea_norm_params& np, const float* cond, const float* x,
                   long long ldx, int dim, void* y_act, float* y_f32, long long ldy_f32, float* stats,
                   const sea_block_params* tipi, const float* tipi_g, float* x_out) {
  NormCall nc{};
  sea_norm_args& a = nc.a;
  a.x = x; a.ldx = ldx; a.M = c.M; a.d = dim; a.kind = kind;
  a.weight = np.weight.p;
  a.bias = kind == SEA_NORM_ADALN ? np.bias.p : nullptr;
  a.cond = cond; a.ldc = 2LL * dim; a.cond_div = c.cond_div;
  a.cond_folded = (kind == SEA_NORM_ADALN && c.inv) ? 1 : 0;   // per-trajectory rows hold gamma | beta already
  if (tipi && c.inv) {
    a.add_rows = tipi_g; a.ld_add = dim; a.add_div = c.cond_div;  // tipi_g = per-trajectory TIPI rows
    a.x_out = x_out; a.ldxo = dim;
  } else if (tipi) {
    a.tipi_g = tipi_g; a.tipi_hid = c.d->ib_hidden;
    a.tipi_w = tipi->ib3_w.p; a.tipi_b = tipi->ib3_b.p;
    a.x_out = x_out; a.ldxo = dim;
  }
  if (y_f32) { a.y_f32 = y_f32; a.ldy_f32 = ldy_f32; }
  else if (c.fp32) { a.y_f32 = static_cast<float*>(y_act); a.ldy_f32 = dim; }
  else { a.y_bf16 = y_act; a.ldy_bf16 = dim; }
  a.stats = stats;
  const double esz = c.fp32 ? 4.0 : 2.0;
  nc.bytes = static_cast<double>(c.M) * dim * (4.0 + (y_f32 ? 4.0 : esz) + (tipi ? 4.0 : 0.0) +
                                               (kind == SEA_NORM_ADALN ? 8.0 : 0.0));
  return nc;
}

// The V field streams' norms are independent: one launch for all of them.
int norm_group(Ctx& c, int n, const NormCall* calls) {
  sea_norm_args args[SEA_MAX_STREAMS];
  double bytes = 0;
  for (int i = 0; i < n; ++i) { args[i] = calls[i].a; bytes += calls[i].bytes; }
  ++g_launches;
  ProfScope prof(c.s, SEA_PROF_ELEMWISE, bytes);
  return sea_norm_fwd_group(n, args, reinterpret_cast<sea_stream_t>(c.s));
}

sea_attn_args attn_call(Ctx& c, const void* q, long long ldq, const void* k, const void* v, long long ldkv,
                        void* o, long long ldo, float* lse, int head_dim, unsigned site = 0) {
  sea_attn_args a{};
  a.dropout_p = c.drop_p; a.dropout_site = site; a.dropout_seed = c.d->dropout_seed;
  a.q = q; a.k = k; a.v = v; a.ldq = ldq; a.ldk = ldkv; a.ldv = ldkv;
  a.o = o; a.ldo = ldo; a.lse = lse;
  a.B = c.B; a.T = c.T; a.n_heads = c.d->n_heads; a.head_dim = head_dim;
  a.src_len = c.d->src_len;
  a.scale = 1.0f / sqrtf(static_cast<float>(head_dim));
  a.prec = c.fp32 ? SEA_PREC_FP32 : SEA_PREC_BF16;
  return a;
}

sea_attn_decode_args decode_call(Ctx& c, const void* q, long long ldq, const void* k, const void* v, long long ldkv,
                                 long long kv_bs, void* o, long long ldo, int n_keys, int head_dim) {
  sea_attn_decode_args a{};
  a.q = q; a.k = k; a.v = v; a.o = o;
  a.ldq = ldq; a.ldo = ldo; a.ldk = ldkv; a.ldv = ldkv; a.k_batch_stride = kv_bs; a.v_batch_stride = kv_bs;
  a.B = c.B; a.n_keys = n_keys; a.n_heads = c.d->n_heads; a.head_dim = head_dim;
  a.scale = 1.0f / sqrtf(static_cast<float>(head_dim));
  a.prec = c.fp32 ? SEA_PREC_FP32 : SEA_PREC_BF16;
  return a;
}

int decode_group(Ctx& c, int n, const sea_attn_decode_args* a) {
  ++g_launches;
  ProfScope prof(c.s, SEA_PROF_ATTN, 4.0 * n * c.B * c.d->n_heads * static_cast<double>(a[0].n_keys) * a[0].head_dim);
  return sea_attention_decode_group(n, a, reinterpret_cast<sea_stream_t>(c.s));
}

int attention_group(Ctx& c, int n, const sea_attn_args* a) {
  ++g_launches;
  ProfScope prof(c.s, SEA_PROF_ATTN,
                 2.0 * n * c.B * c.d->n_heads * static_cast<double>(c.T) * c.T * a[0].head_dim);
  return sea_attention_fwd_group(n, a, reinterpret_cast<sea_stream_t>(c.s));
}

static int adaln_cond(Ctx& c, int n, const sea_norm_params* const* np, void* const* hid,
                      const PackedLinear* const* W, float* const* cond, int d2, const float* ib) {
  // cond = Linear(2d,2d)(SiLU(Linear(ib_num,2d)(ib)))   models/base_blocks.py:337-344
  LinIn in[SEA_MAX_STREAMS];
  LinOut out[SEA_MAX_STREAMS];
  const float* w1[SEA_MAX_STREAMS]; const float* b1[SEA_MAX_STREAMS];
  void* ob[SEA_MAX_STREAMS]; float* of[SEA_MAX_STREAMS];
  for (int g = 0; g < n; ++g) {
    w1[g] = np[g]->c0_w.p; b1[g] = np[g]->c0_b.p;
    ob[g] = c.fp32 ? nullptr : hid[g];
    of[g] = c.fp32 ? static_cast<float*>(hid[g]) : nullptr;
  }
  SEA_TRY(sea_adaln_hidden_group(n, w1, b1, ob, of, ib, c.ld_ib, c.Mc, c.d->ib_num, d2,
                                 reinterpret_cast<sea_stream_t>(c.s)));
  ++g_launches;
  for (int g = 0; g < n; ++g) {
    in[g] = LinIn{hid[g], d2, 0};
    out[g] = LinOut{};
    out[g].bias = np[g]->c2_b.p;
    out[g].f32 = cond[g]; out[g].ld_f32 = d2;
  }
  return linear_group(c, n, in, W, out, c.Mc);
}

}  // namespace sea

extern "C" size_t sea_temporal_cond_cache_bytes(const sea_temporal_desc* d, int B) {
  if (!d || B <= 0) return 0;
  Arena ar{nullptr};
  Tape t;
  layout_cond_cache(d, B, ar, t);
  return ar.off + 256;
}

extern "C" size_t sea_temporal_workspace_bytes(const sea_temporal_desc* d, int B, int T, int training) {
  if (!d || B <= 0 || T <= 0) return 0;
  Arena ar{nullptr};
  Tape t;
  layout_tape(d, B, T, training != 0, ar, t);
  if (training) {
    BwdTape bt;
    layout_bwd_tape(d, B, T, ar, bt);
  }
  return ar.off + 256;
}

// ------------------------------------------------------------------ KV-cached incremental step
// What a later position needs from earlier ones (everything else on the path is per token):
//   self-attention   q|k|v rows of every stream           [B, max_len, 3E]   (RoPE'd q, k)
//   exchange         k|v rows of every ordered pair (i,j)  [B, max_len, 2Dd]
// in the activation dtype of the precision mode.  Position t of trajectory b is row b*max_len + t.
namespace sea {
struct KvLayout {
  struct S { void* qkv; void* kvc[SEA_MAX_STREAMS]; };
  std::vector<S> L;  // [num_layers * num_streams]
};
static void layout_kv(const sea_temporal_desc* d, int B, int max_len, Arena& ar, KvLayout& kv) {
  const size_t esz = d->precision == SEA_PREC_FP32 ? 4 : 2;
  const size_t rows = static_cast<size_t>(B) * max_len;
  const int V = d->num_streams;
  kv.L.assign(static_cast<size_t>(d->num_layers) * V, KvLayout::S{});
  for (int l = 0; l < d->num_layers; ++l)
    for (int i = 0; i < V; ++i) {
      KvLayout::S& s = kv.L[static_cast<size_t>(l) * V + i];
      s.qkv = ar.take(rows * 3 * d->embed_dim * esz);
      for (int j = 0; j < V; ++j)
        if (j != i) s.kvc[j] = ar.take(rows * 2 * d->down_dim * esz);
    }
}
struct StepCtx {
  const KvLayout* kv;
  int pos;       // absolute position of the single new token
  int max_len;   // positions per trajectory in the caches
  long long x_bs, ib_bs, y_bs;  // pitches between trajectories of x_t / ib_t / y_t (elements)
};
}  // namespace sea

extern "C" size_t sea_temporal_kv_cache_bytes(const sea_temporal_desc* d, int B, int max_len) {
  if (!d || B <= 0 || max_len <= 0) return 0;
  Arena ar{nullptr};
  KvLayout kv;
  layout_kv(d, B, max_len, ar, kv);
  return ar.off + 256;
}

static int forward_impl(const sea_temporal_desc* d, const void* cache, const float* x, const float* ib, float* y,
                        int B, int T, void* workspace, size_t workspace_bytes, int training,
                        const StepCtx* step, sea_stream_t stream, long long x_seq_bs = 0);

extern "C" int sea_temporal_forward(const sea_temporal_desc* d, const void* cache, const float* x,
                                    const float* ib, float* y, int B, int T, void* workspace,
                                    size_t workspace_bytes, int training, sea_stream_t stream) {
  return forward_impl(d, cache, x, ib, y, B, T, workspace, workspace_bytes, training, nullptr, stream);
}

extern "C" int sea_temporal_step(const sea_temporal_desc* d, const void* cache, void* kv_cache,
                                 size_t kv_cache_bytes, int max_len, const float* x_t, int64_t x_batch_stride,
                                 const float* ib_t, int64_t ib_batch_stride, float* y_t, int64_t y_batch_stride,
                                 int B, int pos, void* workspace, size_t workspace_bytes, sea_stream_t stream) {
  if (!d || !kv_cache || max_len <= 0 || pos < 0 || pos >= max_len) return SEA_ERR_INVALID;
  if (max_len > d->max_len) return SEA_ERR_UNSUPPORTED;
  // tril(diagonal = src_len > 0) lets a query see src_len FUTURE keys (models/base_blocks.py:173,192): the prefix
  // loop's outputs at earlier positions then change as the sequence grows, and no key/value cache reproduces it
  if (d->src_len != 0) return SEA_ERR_UNSUPPORTED;
  if (kv_cache_bytes < sea_temporal_kv_cache_bytes(d, B, max_len)) return SEA_ERR_WORKSPACE;
  const int V = d->num_streams;
  if (x_batch_stride < static_cast<int64_t>(V) * d->embed_dim || (x_batch_stride % 4) ||
      y_batch_stride < static_cast<int64_t>(V) * d->embed_dim || (y_batch_stride % 4) || ib_batch_stride < d->ib_num)
    return SEA_ERR_INVALID;
  Arena ar{static_cast<char*>(kv_cache)};
  KvLayout kv;
  layout_kv(d, B, max_len, ar, kv);
  StepCtx sc{&kv, pos, max_len, x_batch_stride, ib_batch_stride, y_batch_stride};
  return forward_impl(d, cache, x_t, ib_t, y_t, B, 1, workspace, workspace_bytes, 0, &sc, stream);
}

extern "C" int sea_temporal_forward_strided(const sea_temporal_desc* d, const void* cache, const float* x,
                                            int64_t x_batch_stride, const float* ib, float* y, int B, int T,
                                            void* workspace, size_t workspace_bytes, sea_stream_t stream) {
  if (!d || x_batch_stride < static_cast<int64_t>(T) * d->num_streams * d->embed_dim || (x_batch_stride % 8))
    return SEA_ERR_INVALID;
  return forward_impl(d, cache, x, ib, y, B, T, workspace, workspace_bytes, 0, nullptr, stream, x_batch_stride);
}

static int forward_impl(const sea_temporal_desc* d, const void* cache, const float* x, const float* ib, float* y,
                        int B, int T, void* workspace, size_t workspace_bytes, int training,
                        const StepCtx* step, sea_stream_t stream, long long x_seq_bs) {
  SEA_TRY(validate_desc(d));
  if (!cache || !x || !ib || !y || !workspace || B <= 0 || T <= 0) return SEA_ERR_INVALID;
  if (T > d->max_len) return SEA_ERR_UNSUPPORTED;
  if (workspace_bytes < sea_temporal_workspace_bytes(d, B, T, training)) return SEA_ERR_WORKSPACE;
  SEA_TRY(ensure_init());
  g_launches = 0;
  pdl_fence_next();

  Arena car{const_cast<char*>(static_cast<const char*>(cache))};
  CacheLayout cl;
  layout_cache(d, training != 0, car, cl);
  Arena war{static_cast<char*>(workspace)};
  Tape tape;
  layout_tape(d, B, T, training != 0, war, tape);
  if (d->splitk_slot < 0 || d->splitk_slot > 3) return SEA_ERR_INVALID;
  SplitKGuard splitk_guard;
  SEA_TRY(sea_gemm_set_workspace(cl.splitk[d->splitk_slot], cl.splitk_bytes));

  Ctx c{};
  c.d = d; c.cache = &cl; c.tape = &tape;
  c.s = reinterpret_cast<cudaStream_t>(stream);
  c.fp32 = d->precision == SEA_PREC_FP32;
  c.B = B; c.T = T; c.M = B * T;
  c.pos0 = step ? step->pos : 0;
  c.drop_p = (training != 0 && d->dropout_p > 0.f) ? d->dropout_p : 0.f;
  if (d->dropout_p < 0.f || d->dropout_p >= 1.f) return SEA_ERR_INVALID;
  if (c.drop_p > 0.f && c.fp32) return SEA_ERR_UNSUPPORTED;   // training runs in bf16 mode
  // time-invariant condition (inference): one cond / TIPI row per trajectory instead of per token
  const bool inv = d->ib_time_invariant != 0 && training == 0 && (T > 1 || step != nullptr);
  c.inv = inv;
  c.Mc = inv ? B : c.M;
  c.ld_ib = step ? step->ib_bs : (inv ? static_cast<long long>(T) * d->ib_num : d->ib_num);
  c.cond_div = inv ? T : 1;
  const int M = c.M, V = d->num_streams, E = d->embed_dim, Dd = d->down_dim, H = d->hidden_dim;
  const int hd = E / d->n_heads, hdc = Dd / d->n_heads;
  const bool ada = d->norm_kind == SEA_NORM_ADALN;
  const int kind = d->norm_kind;
  const size_t esz = c.fp32 ? 4 : 2;
  sea_stream_t st = stream;
  // two-stream schedule (inference, bf16): see the fork inside the exchange loop.  Needs the caller's auxiliary stream and
  // events (desc->aux_stream, fork_events, join_event); the auxiliary GEMMs get their own stream-K workspace.
  bool fork = d->aux_stream != nullptr && d->join_event != nullptr && training == 0 && !c.fp32 && V >= 2 && !g_prof_on;
  for (int i = 0; fork && i < V - 1; ++i) fork = d->fork_events[i] != nullptr;
  c.splitk = cl.splitk[d->splitk_slot];
  Ctx caux = c;
  caux.s = reinterpret_cast<cudaStream_t>(d->aux_stream);
  caux.splitk = cl.splitk[(d->splitk_slot + 2) & 3];

  // ---- everything that depends on ib only: TIPI hidden + all AdaLN conditions (hoisted) ----
  // With a time-invariant condition the results live in the caller's persistent cond cache and
  // are reused by later calls on the same trajectories (the rollout loop) while it stays valid.
  bool reuse = false;
  if (inv && d->cond_cache != nullptr) {
    if (d->cond_cache_bytes < sea_temporal_cond_cache_bytes(d, B)) return SEA_ERR_WORKSPACE;
    Arena cca{static_cast<char*>(d->cond_cache)};
    layout_cond_cache(d, B, cca, tape);
    reuse = d->cond_cache_valid != 0;
  }
  if (!reuse) {
    for (int l = 0; l < d->num_layers; ++l) {
      const sea_block_params& bp = d->blocks[l];
      LayerTape& lt = tape.L[l];
      SEA_TRY(sea_tipi_hidden(ib, c.ld_ib, c.Mc, d->ib_num, bp.ib0_w.p, bp.ib0_b.p, bp.ib_ln_w.p, bp.ib_ln_b.p,
                              d->ib_hidden, lt.tipi_g, lt.tipi_pre, lt.tipi_st, st));
      ++g_launches;
      if (inv) {
        SEA_TRY(sea_tipi_rows(lt.tipi_g, d->ib_hidden, B, E, d->ib_hidden, bp.ib3_w.p, bp.ib3_b.p, lt.tipi_rows, st));
        ++g_launches;
      }
      if (ada) {
        const sea_norm_params* np[SEA_MAX_STREAMS]; void* hid[SEA_MAX_STREAMS];
        const PackedLinear* W[SEA_MAX_STREAMS]; float* cond[SEA_MAX_STREAMS];
        if (2 * V <= SEA_MAX_STREAMS) {  // ln0 and ln2 of every stream in one grouped launch
          for (int i = 0; i < V; ++i) {
            np[i] = &bp.s[i].ln0; hid[i] = lt.s[i].hid0; W[i] = &cl.blocks[l].s[i].c2_ln0; cond[i] = lt.s[i].cond0;
            np[V + i] = &bp.s[i].ln2; hid[V + i] = lt.s[i].hid2; W[V + i] = &cl.blocks[l].s[i].c2_ln2; cond[V + i] = lt.s[i].cond2;
          }
          SEA_TRY(adaln_cond(c, 2 * V, np, hid, W, cond, 2 * E, ib));
        } else {
          for (int which = 0; which < 2; ++which) {
            for (int i = 0; i < V; ++i) {
              np[i] = which ? &bp.s[i].ln2 : &bp.s[i].ln0;
              hid[i] = which ? lt.s[i].hid2 : lt.s[i].hid0;
              W[i] = which ? &cl.blocks[l].s[i].c2_ln2 : &cl.blocks[l].s[i].c2_ln0;
              cond[i] = which ? lt.s[i].cond2 : lt.s[i].cond0;
            }
            SEA_TRY(adaln_cond(c, V, np, hid, W, cond, 2 * E, ib));
          }
        }
        for (int i = 0; i < V; ++i) {
          np[i] = &bp.s[i].ln_cross; hid[i] = lt.s[i].hidc;
          W[i] = &cl.blocks[l].s[i].c2_lnc; cond[i] = lt.s[i].condc;
        }
        SEA_TRY(adaln_cond(c, V, np, hid, W, cond, 2 * Dd, ib));
      }
    }
    if (ada) {
      const sea_norm_params* np[SEA_MAX_STREAMS]; void* hid[SEA_MAX_STREAMS];
      const PackedLinear* W[SEA_MAX_STREAMS]; float* cond[SEA_MAX_STREAMS];
      for (int i = 0; i < V; ++i) {
        np[i] = &d->final_ln[i]; hid[i] = tape.hidF[i]; W[i] = &cl.c2_final[i]; cond[i] = tape.condF[i];
      }
      SEA_TRY(adaln_cond(c, V, np, hid, W, cond, 2 * E, ib));
    }
    if (ada && inv) {
      // one condition row per trajectory: fold AdaLN's own weight / bias into it once, so the norm
      // kernels read two vectors per row instead of four (models/base_blocks.py:345-350)
      for (int l = 0; l < d->num_layers; ++l)
        for (int i = 0; i < V; ++i) {
          const sea_stream_params& sp = d->blocks[l].s[i];
          StreamTape& stp = tape.L[l].s[i];
          SEA_TRY(sea_adaln_fold(stp.cond0, 2LL * E, B, E, sp.ln0.weight.p, sp.ln0.bias.p, st));
          SEA_TRY(sea_adaln_fold(stp.cond2, 2LL * E, B, E, sp.ln2.weight.p, sp.ln2.bias.p, st));
          SEA_TRY(sea_adaln_fold(stp.condc, 2LL * Dd, B, Dd, sp.ln_cross.weight.p, sp.ln_cross.bias.p, st));
          g_launches += 3;
        }
      for (int i = 0; i < V; ++i) {
        SEA_TRY(sea_adaln_fold(tape.condF[i], 2LL * E, B, E, d->final_ln[i].weight.p, d->final_ln[i].bias.p, st));
        ++g_launches;
      }
    }
  }

  // ---- layers ----
  const float* xin[SEA_MAX_STREAMS];
  long long ldxin = step ? step->x_bs : static_cast<long long>(V) * E;   // step mode: one row per trajectory
  for (int i = 0; i < V; ++i) xin[i] = x + static_cast<long long>(i) * E;  // x[:, :, i, :]

  for (int l = 0; l < d->num_layers; ++l) {
    const sea_block_params& bp = d->blocks[l];
    LayerTape& lt = tape.L[l];
    const BlockCache& bc = cl.blocks[l];
    LinIn in[SEA_MAX_STREAMS];
    LinOut out[SEA_MAX_STREAMS];
    const PackedLinear* W[SEA_MAX_STREAMS];

    // (1) x_i += SelfAttn_i(Norm_{i,0}(x_i))          models/temporal.py:135-136
    NormCall ncall[SEA_MAX_STREAMS];
    sea_attn_args acall[SEA_MAX_STREAMS];
    // layer 0 may read a prefix of a longer [B, steps+1, V, E] sequence buffer in place
    const bool xs = (l == 0 && x_seq_bs > 0);
    for (int i = 0; i < V; ++i) {
      ncall[i] = norm_call(c, kind, bp.s[i].ln0, lt.s[i].cond0, xin[i], ldxin, E, lt.s[i].n0, nullptr, 0,
                           lt.s[i].st0, nullptr, nullptr, nullptr);
      if (xs) { ncall[i].a.x_rows_per_batch = T; ncall[i].a.x_batch_stride = x_seq_bs; }
    }
    SEA_TRY(norm_group(c, V, ncall));
    for (int i = 0; i < V; ++i) {
      in[i] = LinIn{lt.s[i].n0, E, 0};
      W[i] = &bc.s[i].qkv;
      out[i] = LinOut{};
      out[i].bias = bc.s[i].qkv_bias;
      if (step) {  // the new row goes straight into the cache: [b, pos, :]
        char* row = static_cast<char*>(step->kv->L[static_cast<size_t>(l) * V + i].qkv) + esz * 3 * E * step->pos;
        out[i].post = row; out[i].ld_post = 3LL * E * step->max_len;
      } else {
        out[i].post = lt.s[i].qkv; out[i].ld_post = 3 * E;
      }
      out[i].rope_cols = 2 * E; out[i].head_dim = hd; out[i].rope_table = d->rope_self;
    }
    SEA_TRY(linear_group(c, V, in, W, out, M));
    if (step) {
      sea_attn_decode_args dc[SEA_MAX_STREAMS];
      for (int i = 0; i < V; ++i) {
        char* base = static_cast<char*>(step->kv->L[static_cast<size_t>(l) * V + i].qkv);
        dc[i] = decode_call(c, base + esz * 3 * E * step->pos, 3LL * E * step->max_len, base + esz * E,
                            base + esz * 2 * E, 3 * E, 3LL * E * step->max_len, lt.s[i].ao, E, step->pos + 1, hd);
      }
      SEA_TRY(decode_group(c, V, dc));
    } else {
      for (int i = 0; i < V; ++i) {
        char* base = static_cast<char*>(lt.s[i].qkv);
        acall[i] = attn_call(c, base, 3 * E, base + esz * E, base + esz * 2 * E, 3 * E, lt.s[i].ao, E,
                             lt.s[i].lse, hd, drop_site(l, SEA_SITE_SELF, i, 0));
      }
      SEA_TRY(attention_group(c, V, acall));
    }
    for (int i = 0; i < V; ++i) {
      in[i] = LinIn{lt.s[i].ao, E, 0};
      W[i] = &bc.s[i].sproj;
      out[i] = LinOut{};
      out[i].residual = xin[i]; out[i].ld_res = ldxin;
      if (xs) { out[i].res_rows = T; out[i].res_bs = x_seq_bs; }
      out[i].f32 = lt.s[i].x1; out[i].ld_f32 = E;
      out[i].pre = lt.s[i].x1b; out[i].ld_pre = E;
    }
    SEA_TRY(linear_group(c, V, in, W, out, M));

    // (3) x_i += TIPI(ib) fused with Norm_{i,2}; (4) MLP; (5) proj    models/temporal.py:140-146
    // for the streams [i0, i1), enqueued through `cx` (the caller's stream, or the auxiliary one: see the fork below)
    auto block_tail = [&](Ctx& cx, int i0, int i1) -> int {
      const int nn = i1 - i0;
      NormCall nc2[SEA_MAX_STREAMS];
      LinIn tin[SEA_MAX_STREAMS];
      LinOut tout[SEA_MAX_STREAMS];
      const PackedLinear* tW[SEA_MAX_STREAMS];
      for (int i = i0; i < i1; ++i) {
        nc2[i - i0] = norm_call(cx, kind, bp.s[i].ln2, lt.s[i].cond2, lt.s[i].xp, E, E, lt.s[i].n2, nullptr, 0,
                             lt.s[i].st2, &bp, inv ? lt.tipi_rows : lt.tipi_g, lt.s[i].x2);
        if (cx.drop_p > 0.f) {   // the ib-MLP is called once per stream: independent masks (temporal.py:140-142)
          nc2[i - i0].a.tipi_dropout_p = cx.drop_p; nc2[i - i0].a.tipi_dropout_site = drop_site(l, SEA_SITE_TIPI, i, 0);
          nc2[i - i0].a.tipi_dropout_seed = d->dropout_seed;
        }
      }
      SEA_TRY(norm_group(cx, nn, nc2));
      for (int i = i0; i < i1; ++i) {
        tin[i - i0] = LinIn{lt.s[i].n2, E, 0};
        tW[i - i0] = &bc.s[i].mlp0;
        tout[i - i0] = LinOut{};
        tout[i - i0].bias = bp.s[i].mlp0_b.p;
        tout[i - i0].post = lt.s[i].h; tout[i - i0].ld_post = H;
      }
      SEA_TRY(linear_group(cx, nn, tin, tW, tout, M));
      {
        sea_ln_gelu_args la[SEA_MAX_STREAMS];
        for (int i = i0; i < i1; ++i) {
          sea_ln_gelu_args& a = la[i - i0];
          a = sea_ln_gelu_args{};
          if (cx.fp32) { a.h_f32 = static_cast<const float*>(lt.s[i].h); a.g_f32 = static_cast<float*>(lt.s[i].gh); }
          else { a.h_bf16 = lt.s[i].h; a.g_bf16 = lt.s[i].gh; }
          a.ldh = H; a.ldg = H; a.M = M; a.H = H;
          a.weight = bp.s[i].mlp_ln_w.p; a.bias = bp.s[i].mlp_ln_b.p; a.stats = lt.s[i].stH;
        }
        ProfScope prof(cx.s, SEA_PROF_ELEMWISE, 2.0 * esz * M * static_cast<double>(H) * nn);
        SEA_TRY(sea_ln_gelu_fwd_group(nn, la, reinterpret_cast<sea_stream_t>(cx.s)));
        ++g_launches;
      }
      for (int i = i0; i < i1; ++i) {
        tin[i - i0] = LinIn{lt.s[i].gh, H, 0};
        tW[i - i0] = &bc.s[i].mlp3;
        tout[i - i0] = LinOut{};
        tout[i - i0].bias = bp.s[i].mlp3_b.p;
        tout[i - i0].residual = lt.s[i].x2; tout[i - i0].ld_res = E;
        tout[i - i0].post = lt.s[i].x3; tout[i - i0].ld_post = E;
        tout[i - i0].drop_p = cx.drop_p; tout[i - i0].drop_site = drop_site(l, SEA_SITE_MLP, i, 0);
      }
      SEA_TRY(linear_group(cx, nn, tin, tW, tout, M));
      for (int i = i0; i < i1; ++i) {
        tin[i - i0] = LinIn{lt.s[i].x3, E, 0};
        tW[i - i0] = &bc.s[i].proj;
        tout[i - i0] = LinOut{};
        tout[i - i0].bias = bp.s[i].proj_b.p;
        tout[i - i0].f32 = lt.s[i].xout; tout[i - i0].ld_f32 = E;
      }
      SEA_TRY(linear_group(cx, nn, tin, tW, tout, M));
      return SEA_OK;
    };
    // (2) State-Exchange Attention, sequential over i   models/temporal.py:176-192
    for (int i = 0; i < V; ++i) {
      in[i] = LinIn{c.fp32 ? static_cast<const void*>(lt.s[i].x1) : lt.s[i].x1b, E, 0};
      W[i] = &bc.s[i].down;
      out[i] = LinOut{};
      out[i].bias = bp.s[i].down_b.p;
      out[i].f32 = lt.s[i].dpre; out[i].ld_f32 = Dd;
    }
    SEA_TRY(linear_group(c, V, in, W, out, M));
    for (int i = 0; i < V; ++i)
      ncall[i] = norm_call(c, kind, bp.s[i].ln_cross, lt.s[i].condc, lt.s[i].dpre, Dd, Dd, lt.s[i].npre,
                           nullptr, 0, lt.s[i].stc_pre, nullptr, nullptr, nullptr);
    SEA_TRY(norm_group(c, V, ncall));
    // Every q projection, and the k / v projections whose source stream has not been exchanged yet
    // (j > i), read only the pre-exchange ln_cross outputs: same shape [M, Dd] x [Dd, Dd], so they
    // go out together, up to four per launch (models/base_blocks.py:271-276).
    {
      LinIn gin[SEA_MAX_STREAMS]; LinOut gout[SEA_MAX_STREAMS];
      PackedLinear gw[SEA_MAX_STREAMS]; const PackedLinear* gW[SEA_MAX_STREAMS];
      int ng = 0;
      auto flush = [&]() -> int {
        if (ng == 0) return SEA_OK;
        for (int g = 0; g < ng; ++g) gW[g] = &gw[g];
        const int rc = linear_group(c, ng, gin, gW, gout, M);
        ng = 0;
        return rc;
      };
      auto push = [&](const void* a_src, const PackedLinear& w, int row_off, const float* bias, void* dst,
                      long long ld_dst, bool rope) -> int {
        gin[ng] = LinIn{a_src, Dd, 0};
        gw[ng] = w;
        gw[ng].N = Dd;
        gw[ng].w = w.w + static_cast<long long>(row_off) * w.ldw;
        gout[ng] = LinOut{};
        gout[ng].bias = bias;
        gout[ng].post = dst; gout[ng].ld_post = ld_dst;
        if (rope) { gout[ng].rope_cols = Dd; gout[ng].head_dim = hdc; gout[ng].rope_table = d->rope_cross; }
        if (++ng == SEA_MAX_STREAMS) return flush();
        return SEA_OK;
      };
      for (int i = 0; i < V; ++i) {
        StreamTape& s = lt.s[i];
        for (int j = 0; j < V; ++j) {
          if (j == i) continue;
          SEA_TRY(push(s.npre, bc.s[i].cq[j], 0, bp.s[i].cross_attn[j].q_b.p, s.q[j], Dd, true));
          if (j > i) {
            char* kvb = static_cast<char*>(s.kv[j]);
            long long ldkv = 2 * Dd;
            if (step) {
              kvb = static_cast<char*>(step->kv->L[static_cast<size_t>(l) * V + i].kvc[j]) + esz * 2 * Dd * step->pos;
              ldkv = 2LL * Dd * step->max_len;
            }
            SEA_TRY(push(lt.s[j].npre, bc.s[i].ckv[j], 0, bc.s[i].ckv_bias[j], kvb, ldkv, true));
            SEA_TRY(push(lt.s[j].npre, bc.s[i].ckv[j], Dd, bc.s[i].ckv_bias[j] + Dd, kvb + esz * Dd, ldkv, false));
          }
        }
      }
      SEA_TRY(flush());
    }
    for (int i = 0; i < V; ++i) {
      StreamTape& s = lt.s[i];
      const float* xcur = s.x1;
      for (int j = 0; j < V; ++j) {
        if (j == i) continue;
        if (j < i) {
          // (k,v) from the already exchanged stream j     models/temporal.py:189-191
          in[0] = LinIn{lt.s[j].npost, Dd, 0};
          W[0] = &bc.s[i].ckv[j];
          out[0] = LinOut{};
          out[0].bias = bc.s[i].ckv_bias[j];
          if (step) {
            out[0].post = static_cast<char*>(step->kv->L[static_cast<size_t>(l) * V + i].kvc[j]) + esz * 2 * Dd * step->pos;
            out[0].ld_post = 2LL * Dd * step->max_len;
          } else {
            out[0].post = s.kv[j]; out[0].ld_post = 2 * Dd;
          }
          out[0].rope_cols = Dd; out[0].head_dim = hdc; out[0].rope_table = d->rope_cross;
          SEA_TRY(linear_group(c, 1, in, W, out, M));
        }
        if (step) {
          char* kvb = static_cast<char*>(step->kv->L[static_cast<size_t>(l) * V + i].kvc[j]);
          sea_attn_decode_args dc = decode_call(c, s.q[j], Dd, kvb, kvb + esz * Dd, 2 * Dd, 2LL * Dd * step->max_len,
                                                s.a[j], Dd, step->pos + 1, hdc);
          SEA_TRY(decode_group(c, 1, &dc));
        } else {
          char* kvb = static_cast<char*>(s.kv[j]);
          acall[0] = attn_call(c, s.q[j], Dd, kvb, kvb + esz * Dd, 2 * Dd, s.a[j], Dd, s.lse_c[j], hdc,
                               drop_site(l, SEA_SITE_CROSS, i, j));
          SEA_TRY(attention_group(c, 1, acall));
        }
        // cross_up(GELU(projection(attn)))             models/base_blocks.py:293, temporal.py:185
        in[0] = LinIn{s.a[j], Dd, 0};
        W[0] = &bc.s[i].cproj[j];
        out[0] = LinOut{};
        out[0].pre = s.p[j]; out[0].ld_pre = Dd;
        out[0].post = s.g[j]; out[0].ld_post = Dd; out[0].act = SEA_ACT_GELU;
        SEA_TRY(linear_group(c, 1, in, W, out, M));
        in[0] = c.fp32 ? LinIn{s.p[j], Dd, SEA_ACT_GELU} : LinIn{s.g[j], Dd, 0};
        W[0] = &bc.s[i].up;
        out[0] = LinOut{};
        out[0].bias = bp.s[i].up_b.p;
        out[0].residual = xcur; out[0].ld_res = E;
        out[0].f32 = s.xp; out[0].ld_f32 = E;
        out[0].pre = s.xpb; out[0].ld_pre = E;
        SEA_TRY(linear_group(c, 1, in, W, out, M));
        xcur = s.xp;
      }
      if (V == 1) {  // no partner stream: exchange is the identity
        SEA_CUDA_OK(cudaMemcpyAsync(s.xp, s.x1, sizeof(float) * M * E, cudaMemcpyDeviceToDevice, c.s));
      }
      if (fork && i < V - 1) {
        // stream i is final for this block: its TIPI / MLP / proj tail (half of the block's FLOPs at V = 2) goes to the
        // auxiliary stream and fills the SMs the remaining, latency-bound exchange steps of the later streams leave idle
        SEA_CUDA_OK(cudaEventRecord(static_cast<cudaEvent_t>(d->fork_events[i]), c.s));
        SEA_CUDA_OK(cudaStreamWaitEvent(caux.s, static_cast<cudaEvent_t>(d->fork_events[i]), 0));
        SEA_TRY(block_tail(caux, i, i + 1));
      }
      if (i < V - 1) {  // later streams see the UPDATED stream i (Gauss–Seidel)
        in[0] = LinIn{c.fp32 ? static_cast<const void*>(s.xp) : s.xpb, E, 0};
        W[0] = &bc.s[i].down;
        out[0] = LinOut{};
        out[0].bias = bp.s[i].down_b.p;
        out[0].f32 = s.dpost; out[0].ld_f32 = Dd;
        SEA_TRY(linear_group(c, 1, in, W, out, M));
        ncall[0] = norm_call(c, kind, bp.s[i].ln_cross, s.condc, s.dpost, Dd, Dd, s.npost, nullptr, 0,
                             s.stc_post, nullptr, nullptr, nullptr);
        SEA_TRY(norm_group(c, 1, ncall));
      }
    }

    if (fork) {   // the last stream's tail on the caller's stream; the others are already running on the auxiliary one
      SEA_TRY(block_tail(c, V - 1, V));
      SEA_CUDA_OK(cudaEventRecord(static_cast<cudaEvent_t>(d->join_event), caux.s));
      SEA_CUDA_OK(cudaStreamWaitEvent(c.s, static_cast<cudaEvent_t>(d->join_event), 0));
    } else {
      SEA_TRY(block_tail(c, 0, V));
    }
    for (int i = 0; i < V; ++i) xin[i] = lt.s[i].xout;
    ldxin = E;
  }

  // ---- final norm per stream, written straight into the strided [B,T,V,E] output ----
  {
    NormCall ncall[SEA_MAX_STREAMS];
    for (int i = 0; i < V; ++i)
      ncall[i] = norm_call(c, kind, d->final_ln[i], tape.condF[i], xin[i], ldxin, E, nullptr,
                           y + static_cast<long long>(i) * E, step ? step->y_bs : static_cast<long long>(V) * E, tape.stF[i],
                           nullptr, nullptr, nullptr);
    SEA_TRY(norm_group(c, V, ncall));
  }
  return SEA_OK;
}

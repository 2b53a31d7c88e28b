// K2 (tcgen05 variant) — fused causal attention forward, bf16 operands, fp32 softmax/accumulate.
// One kernel serves self-attention (models/base_blocks.py:191-197) and the state-exchange
// cross-attention (models/base_blocks.py:283-289): Q and K/V are independent tensor maps.
//
// CTA = one 128-query tile of one (batch, head); keys/values stream in tiles of BKV.
//   warp 0      TMA: Q once, then K_j / V_j tiles into a 2-stage ring (128B-swizzled boxes of
//               64 columns; 3-D tensor maps {col, t, b} so a tile never crosses a batch);
//   warp 1      TMEM allocation + tcgen05.mma issue:
//                 S = Q K_j^T          (SS: both operands K-major in smem, N = BKV)
//                 O (+)= P V_j         (TS: P read from TMEM as packed bf16, V_j MN-major in smem)
//   warps 2..5  online softmax, one query row per thread: tcgen05.ld S -> scale, causal mask
//               (k <= q + src_len, no tril buffer), running max / sum in registers, exp2,
//               P -> bf16 -> tcgen05.st; O is rescaled in TMEM only when the running max grew by
//               more than 2^8 (lazy rescale); final O / l and log-sum-exp written from registers.
// TMEM columns: S [0,BKV), P (packed bf16) overwrites S in place [0,BKV/2) — tcgen05.mma executes
// in issue order, so the next tile's S cannot overtake the P·V that reads P — and O [128,128+HD)
// (256 columns allocated) or, for HD = 256, O [256,512) (512 allocated).
// STAGES = 1 (the whole key range fits one tile: short prefixes of a rollout) halves the shared
// memory, so two CTAs share an SM and B x heads = 256 CTAs run as a single wave.
// Up to SEA_MAX_STREAMS same-shape problems (the V field streams) share one launch (blockIdx.x).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/sea_b200.h"
#include "internal.h"
#include "ptx.cuh"

namespace sea {
namespace {

constexpr int BQ = 128;
constexpr int kThreads = 192;
constexpr uint32_t kColS = 0, kColP = 0;

struct AttnTcItem {
  CUtensorMap tq, tk, tv;
  __nv_bfloat16* o;
  float* lse;
  uint32_t drop_site;
};
struct alignas(64) AttnTcParams {
  AttnTcItem it[SEA_MAX_STREAMS];
  long long ldo;
  int B, T, n_heads, src_len;
  float scale_log2;  // scale * log2(e)
  unsigned long long drop_seed;   // dropout on the probabilities (DROP kernels only)
  uint32_t drop_thresh;
  float drop_scale;
  long long* trace;               // clock64 probe buffer (TRACE kernels only; sea_attention_debug_trace)
};

template <int HD, int BKV, int STAGES_>
struct ACfg {
  static constexpr int ATOMS = HD / 64;
  static constexpr int Q_BYTES = BQ * HD * 2;
  static constexpr int KV_BYTES = BKV * HD * 2;   // one of K or V
  static constexpr int STAGES = STAGES_;
  static constexpr int SMEM = Q_BYTES + STAGES * 2 * KV_BYTES + 1024 + 128;
  static constexpr uint32_t COL_O = HD <= 128 ? 128 : 256;
  static constexpr uint32_t TMEM_COLS = HD <= 128 ? 256 : 512;
  static constexpr int MIN_CTAS = (SMEM <= 110 * 1024 && TMEM_COLS <= 256) ? 2 : 1;
  // HD = 256 (one query tile per CTA: two would not fit TMEM): two S|P buffers of BKV = 64 columns in front of O, so
  // the pipe computes S(j+1) while the softmax warps work on S(j) — the overlap the two-tile kernel gets from its
  // second tile.  Issue order: S0, S1, then per key tile j: PV(j), S(j+2).
  // Measured on B200 (B=4, T=2024): 116.7 us with the two buffers vs 105.5 us with one — the softmax warps, not the
  // pipe, bound this kernel (64 scores per thread against 1024 cycles of MMA per key tile), and issuing S(j+1) first
  // only delays P.V(j) behind it.  Kept as a compile-time switch (SEA_ATTN_PING256), off by default.
#ifdef SEA_ATTN_PING256
  static constexpr bool PING = (HD == 256 && STAGES_ == 2 && 2 * BKV <= static_cast<int>(COL_O));
#else
  static constexpr bool PING = false;
#endif
};

template <int HD, int BKV, int STAGES_, bool DROP>
__global__ void __launch_bounds__(kThreads, (ACfg<HD, BKV, STAGES_>::MIN_CTAS))
attn_fwd_tc_kernel(const __grid_constant__ AttnTcParams pp) {
  using C = ACfg<HD, BKV, STAGES_>;
  constexpr uint32_t kColO = C::COL_O;
  const int bz = blockIdx.x / pp.n_heads;   // (problem, batch); blockIdx.z = query tile: heaviest tiles launch first
  const AttnTcItem& p = pp.it[bz / pp.B];
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + C::Q_BYTES;                        // [STAGES][KV_BYTES]
  uint8_t* sV = sK + C::STAGES * C::KV_BYTES;           // [STAGES][KV_BYTES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + C::STAGES * C::KV_BYTES);
  uint64_t* q_full = bars;          // 1
  uint64_t* kv_full = bars + 1;     // [2]
  uint64_t* kv_empty = bars + 3;    // [2]
  uint64_t* s_full = bars + 5;      // [2] MMA -> softmax (one per S buffer)
  uint64_t* p_full = bars + 7;      // [2] softmax (128 arrivals) -> MMA
  uint64_t* o_done = bars + 9;      // 1   MMA -> softmax
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  constexpr bool PING = C::PING;

  // warp-uniform role index + elect.sync regions: see the note in gemm.cu (no waterfall loops
  // around UTMALDG / UTCHMMA)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int qt = gridDim.z - 1 - blockIdx.z;  // heavy (late) query tiles first, across ALL (batch, head) pairs
  const int h = blockIdx.x % pp.n_heads, b = bz % pp.B;
  const int q0 = qt * BQ;
  const int q_hi = min(pp.T - 1, q0 + BQ - 1);
  const int k_last = min(pp.T - 1, q_hi + pp.src_len);
  const int n_kv = k_last / BKV + 1;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&p.tq);
    ptx::prefetch_tmap(&p.tk);
    ptx::prefetch_tmap(&p.tv);
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&kv_full[s], 1);
      ptx::mbar_init(&kv_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&s_full[s], 1);
      ptx::mbar_init(&p_full[s], 128);
    }
    ptx::mbar_init(o_done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, C::TMEM_COLS);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  ptx::pdl_trigger();
  ptx::pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(q_full, C::Q_BYTES);
#pragma unroll
      for (int a = 0; a < C::ATOMS; ++a)
        ptx::tma_load_3d(sQ + a * (BQ * 128), &p.tq, q_full, h * HD + a * 64, q0, b);
    }
    __syncwarp();
    for (int j = 0; j < n_kv; ++j) {
      const int s = j % C::STAGES;
      const uint32_t ph = (j / C::STAGES) & 1;
      ptx::mbar_wait(&kv_empty[s], ph ^ 1);
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(&kv_full[s], 2 * C::KV_BYTES);
#pragma unroll
        for (int a = 0; a < C::ATOMS; ++a) {
          ptx::tma_load_3d(sK + s * C::KV_BYTES + a * (BKV * 128), &p.tk, &kv_full[s], h * HD + a * 64, j * BKV, b);
          ptx::tma_load_3d(sV + s * C::KV_BYTES + a * (BKV * 128), &p.tv, &kv_full[s], h * HD + a * 64, j * BKV, b);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(BQ, BKV, 0, 0);
    constexpr uint32_t idesc_o = ptx::umma_idesc_bf16(BQ, HD, 0, 1);  // B = V is MN-major
    ptx::mbar_wait(q_full, 0);
    auto issue_s = [&](int j) {
      const int s = j % C::STAGES;
      const int sb = PING ? (j & 1) : 0;
      ptx::mbar_wait(&kv_full[s], (j / C::STAGES) & 1);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t qb = ptx::smem_u32(sQ);
        const uint32_t kb = ptx::smem_u32(sK + s * C::KV_BYTES);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          const uint32_t off = (k >> 2) * (BQ * 128) + (k & 3) * 32;
          const uint32_t koff = (k >> 2) * (BKV * 128) + (k & 3) * 32;
          ptx::umma_f16_ss(tmem + kColS + sb * BKV, ptx::umma_smem_desc(qb + off, 16, 1024),
                           ptx::umma_smem_desc(kb + koff, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&s_full[sb]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int j) {
      const int s = j % C::STAGES;
      const int sb = PING ? (j & 1) : 0;
      ptx::mbar_wait(&p_full[sb], PING ? ((j >> 1) & 1) : (j & 1));
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t vb = ptx::smem_u32(sV + s * C::KV_BYTES);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          // 16 keys per step = two 8-row groups (SBO 1024 B); 64-wide head-dim chunks LBO apart
          const uint64_t vdesc = ptx::umma_smem_desc(vb + k * 2048, BKV * 128, 1024);
          ptx::umma_f16_ts(tmem + kColO, tmem + kColP + sb * BKV + k * 8, vdesc, idesc_o, (j | k) != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&kv_empty[s]);
        ptx::umma_commit(o_done);
      }
      __syncwarp();
    };
    if (PING) {
      issue_s(0);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) issue_s(j + 1);
        issue_pv(j);
      }
    } else {
      for (int j = 0; j < n_kv; ++j) {
        issue_s(j);
        issue_pv(j);
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax warps
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int q = q0 + row;
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    float m_ref = -INFINITY, l_sum = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      const int kv0 = j * BKV;
      const int sb = PING ? (j & 1) : 0;
      const uint32_t colS = kColS + sb * BKV;     // this tile's S buffer; its packed P overwrites it in place
      ptx::mbar_wait(&s_full[sb], PING ? ((j >> 1) & 1) : (j & 1));
      ptx::tc_fence_after();
      const bool need_mask = (kv0 + BKV - 1 > q0 + pp.src_len) || (kv0 + BKV > pp.T);
      // Two sweeps over the score row in 64-column chunks (max, then exponentials): a thread never holds more
      // than 64 scores, so the kernel stays under 128 registers — the per-sub-partition register file
      // (16 K registers, 4 of a CTA pair's 12 warps) is what decides whether two CTAs really share an SM —
      // at the price of one extra ~60-cycle TMEM read per chunk.
      constexpr int CHK = BKV < 64 ? BKV : 64;
      uint32_t r[CHK];
      auto load_chunk = [&](int c) {
#pragma unroll
        for (int i = 0; i < CHK / 32; ++i) ptx::tmem_ld_32x32p(tmem + lane_base + colS + c * CHK + i * 32, r + i * 32);
        ptx::tmem_ld_wait();
        if (need_mask) {
#pragma unroll
          for (int e = 0; e < CHK; ++e) {
            const int kk = kv0 + c * CHK + e;
            if (kk > q + pp.src_len || kk >= pp.T) r[e] = 0xff800000u;  // -inf
          }
        }
      };
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // max of the RAW scores (scale > 0 commutes with max)
#pragma unroll
      for (int c = 0; c < BKV / CHK; ++c) {
        load_chunk(c);
#pragma unroll
        for (int e = 0; e < CHK; e += 8) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            mx4[u] = fmaxf(mx4[u], fmaxf(__uint_as_float(r[e + 2 * u]), __uint_as_float(r[e + 2 * u + 1])));
        }
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      const float m_cand = fmaxf(m_ref, mx * pp.scale_log2);
      const bool grow = (m_cand > m_ref + 8.0f) || (m_ref == -INFINITY && m_cand != -INFINITY);
      const bool any_grow = __any_sync(0xffffffffu, grow);
      // O is owned by the previous P.V until it retires.  With one S buffer so is P's destination; with two, only a
      // rescale of O has to wait (S(j) complete implies P.V(j-2) complete, so the parity wait cannot lag a phase)
      if (j > 0 && (!PING || any_grow)) {
        ptx::mbar_wait(o_done, (j - 1) & 1);
        ptx::tc_fence_after();
      }
      if (any_grow) {
        const float alpha = (m_ref == -INFINITY) ? 0.f : ptx::ex2(m_ref - m_cand);
        l_sum *= alpha;
        m_ref = m_cand;
        if (j > 0) {
#pragma unroll 1
          for (int c = 0; c < HD / 16; ++c) {
            uint32_t o[16];
            ptx::tmem_ld_32x16p(tmem + lane_base + kColO + c * 16, o);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
            ptx::tmem_st_32x16p(tmem + lane_base + kColO + c * 16, o);
          }
        }
      }
      // p = 2^(s * scale_log2 - m): one FFMA + one MUFU per score; packed bf16 pairs overwrite S
      const float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;
      float l0 = 0.f, l1 = 0.f;
      // dropout acts on the NORMALISED probabilities (base_blocks.py:193-194): the row sum stays undropped,
      // only the P that feeds P.V is masked and rescaled
      const unsigned long long drop_row = DROP ? ((static_cast<unsigned long long>(b) * pp.n_heads + h) * pp.T + q) *
                                                     static_cast<unsigned long long>((pp.T + 1) & ~1) + kv0 : 0ull;
#pragma unroll
      for (int c = 0; c < BKV / CHK; ++c) {
        // chunk c of S is still intact: the packed P of chunks < c occupies columns [0, c * CHK / 2)
        load_chunk(c);
#pragma unroll
        for (int e = 0; e < CHK; e += 2) {
          float p0 = ptx::ex2(fmaf(__uint_as_float(r[e]), pp.scale_log2, neg_m));
          float p1 = ptx::ex2(fmaf(__uint_as_float(r[e + 1]), pp.scale_log2, neg_m));
          l0 += p0; l1 += p1;
          if (DROP) {
            const uint2 hsh = ptx::drop_hash(pp.drop_seed, p.drop_site, (drop_row + c * CHK + e) >> 1);
            p0 = hsh.x >= pp.drop_thresh ? p0 * pp.drop_scale : 0.f;
            p1 = hsh.y >= pp.drop_thresh ? p1 * pp.drop_scale : 0.f;
          }
          r[e >> 1] = ptx::pack_bf16(p0, p1);
        }
#pragma unroll
        for (int i = 0; i < CHK / 32; ++i)
          ptx::tmem_st_32x16p(tmem + lane_base + colS + c * (CHK / 2) + i * 16, r + i * 16);
        // the next chunk's load must not overtake these stores in the TMEM pipe (P of chunk c lands in columns
        // below chunk c+1, but the load of chunk c+1 reuses the registers the store is still reading)
        ptx::tmem_st_wait();
      }
      l_sum += l0 + l1;
      ptx::tc_fence_before();
      ptx::mbar_arrive(&p_full[sb]);
    }
    // epilogue: O / l  -> bf16, log-sum-exp
    ptx::mbar_wait(o_done, (n_kv - 1) & 1);
    ptx::tc_fence_after();
    const float inv = 1.f / l_sum;
    __nv_bfloat16* orow = p.o + (static_cast<long long>(b) * pp.T + q) * pp.ldo + h * HD;
#pragma unroll
    for (int c = 0; c < HD / 32; ++c) {
      uint32_t r[32];
      ptx::tmem_ld_32x32(tmem + lane_base + kColO + c * 32, r);
      ptx::tmem_ld_wait();
      if (q < pp.T) {
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
          uint4 o;
          o.x = ptx::pack_bf16(__uint_as_float(r[e]) * inv, __uint_as_float(r[e + 1]) * inv);
          o.y = ptx::pack_bf16(__uint_as_float(r[e + 2]) * inv, __uint_as_float(r[e + 3]) * inv);
          o.z = ptx::pack_bf16(__uint_as_float(r[e + 4]) * inv, __uint_as_float(r[e + 5]) * inv);
          o.w = ptx::pack_bf16(__uint_as_float(r[e + 6]) * inv, __uint_as_float(r[e + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c * 32 + e) = o;
        }
      }
    }
    if (p.lse != nullptr && q < pp.T)
      p.lse[(static_cast<long long>(b) * pp.n_heads + h) * pp.T + q] =
          (m_ref + log2f(l_sum)) * 0.69314718055994530942f;
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// Two query tiles per CTA (T > 128, HD <= 128): the tensor pipe and the softmax warps overlap.
//   warp 0      TMA (Q for 256 query rows once; K_j and V_j rings of 2 stages each, released separately:
//               K_j is free as soon as S(j) is issued, V_j after P.V(j), so both loads run a tile ahead)
//   warp 1      tcgen05.mma issue, interleaved so each softmax group always has its next S in flight:
//                 S0_0, S1_0, then per key tile j:  PV0_j (two halves), S0_{j+1},  PV1_j (two halves), S1_{j+1}
//   warps 2..5  softmax group 0 (query rows q0 .. q0+127),  warps 6..9  group 1 (q0+128 .. q0+255)
// While group 0 exponentiates tile j the pipe computes PV1_{j-1} / S1_j, and vice versa.  A group hands
// its P over in two halves of 64 keys: the pipe starts P.V on the first half while the second half is
// still being exponentiated.  Because the MMAs execute in issue order, "S_g(j) complete" implies
// "PV_g(j-1) complete": a group may rescale its O accumulator as soon as it sees its next S.
// TMEM columns: S0|P0 [0,128)  S1|P1 [128,256)  O0 [256,256+HD)  O1 [384,384+HD).
// The exponentials stay on the SFU: scripts/micro/sfu_probe.cu (profiles/r1d_attention.md) measured 8 cycles
// per score and sub-partition for FFMA + MUFU.EX2 + FADD + pack with four warps (14 with one), and a
// polynomial 2^x on the FMA pipe ADDS to that instead of overlapping (10 cycles with 3 of 8 scores moved).
constexpr int kThreads2 = 320;

template <int HD>
struct ACfg2 {
  static constexpr int BKV = 128;
  static constexpr int HB = 64;     // keys per P hand-off
  static constexpr int ATOMS = HD / 64;
  static constexpr int Q_BYTES = 2 * BQ * HD * 2;
  static constexpr int KV_BYTES = BKV * HD * 2;
  static constexpr int SMEM = Q_BYTES + 2 * 2 * KV_BYTES + 1024 + 256;
};

// TRACE: CTA (0,0,0) records clock64 at the hand-off points: trace[(role * 64 + j) * 8 + k], role 0/1 = softmax
// group, 2 = MMA warp (tuning aid behind sea_attention_debug_trace; never on the product path).
template <int HD, bool DROP, bool TRACE = false>
__global__ void __launch_bounds__(kThreads2, 1) attn_fwd_tc2_kernel(const __grid_constant__ AttnTcParams pp) {
  using C = ACfg2<HD>;
  constexpr int BKV = C::BKV, HB = C::HB;
  const bool tr_on = TRACE && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (threadIdx.x & 31) == 0;
  auto probe = [&](int role, int j, int k) {
    if (TRACE && tr_on && j < 64) pp.trace[(role * 64 + j) * 8 + k] = clock64();
  };
  const int bz = blockIdx.x / pp.n_heads;   // (problem, batch); blockIdx.z = query tile: heaviest tiles launch first
  const AttnTcItem& p = pp.it[bz / pp.B];
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                                   // [ATOMS][256 rows][128 B]
  uint8_t* sK = sQ + C::Q_BYTES;                        // [2][KV_BYTES]
  uint8_t* sV = sK + 2 * C::KV_BYTES;                   // [2][KV_BYTES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * C::KV_BYTES);
  uint64_t* q_full = bars;          // 1
  uint64_t* k_full = bars + 1;      // [2]
  uint64_t* v_full = bars + 3;      // [2]
  uint64_t* k_empty = bars + 5;     // [2]
  uint64_t* v_empty = bars + 7;     // [2]
  uint64_t* s_full = bars + 9;      // [group]: MMA -> softmax
  uint64_t* p_full = bars + 11;     // [group][half]: softmax (128 arrivals) -> MMA
  uint64_t* o_done = bars + 15;     // [group]: MMA -> epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int qb = gridDim.z - 1 - blockIdx.z;  // heavy (late) query blocks first, across ALL (batch, head) pairs
  const int h = blockIdx.x % pp.n_heads, b = bz % pp.B;
  const int q0 = qb * 2 * BQ;
  int n_kv[2];
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const int qg0 = q0 + g * BQ;
    n_kv[g] = qg0 < pp.T ? min(pp.T - 1, min(pp.T - 1, qg0 + BQ - 1) + pp.src_len) / BKV + 1 : 0;
  }
  const int n_tot = max(n_kv[0], n_kv[1]);

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&p.tq);
    ptx::prefetch_tmap(&p.tk);
    ptx::prefetch_tmap(&p.tv);
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&k_full[s], 1);
      ptx::mbar_init(&v_full[s], 1);
      ptx::mbar_init(&k_empty[s], 1);
      ptx::mbar_init(&v_empty[s], 1);
      ptx::mbar_init(&s_full[s], 1);
      ptx::mbar_init(&o_done[s], 1);
    }
    for (int s = 0; s < 4; ++s) ptx::mbar_init(&p_full[s], 128);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  ptx::pdl_trigger();
  ptx::pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(q_full, C::Q_BYTES);
#pragma unroll
      for (int a = 0; a < C::ATOMS; ++a) {
        ptx::tma_load_3d(sQ + a * (2 * BQ * 128), &p.tq, q_full, h * HD + a * 64, q0, b);
        ptx::tma_load_3d(sQ + a * (2 * BQ * 128) + BQ * 128, &p.tq, q_full, h * HD + a * 64, q0 + BQ, b);
      }
    }
    __syncwarp();
    for (int j = 0; j < n_tot; ++j) {
      const int s = j & 1;
      ptx::mbar_wait(&k_empty[s], ((j >> 1) & 1) ^ 1);
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(&k_full[s], C::KV_BYTES);
#pragma unroll
        for (int a = 0; a < C::ATOMS; ++a)
          ptx::tma_load_3d(sK + s * C::KV_BYTES + a * (BKV * 128), &p.tk, &k_full[s], h * HD + a * 64, j * BKV, b);
      }
      __syncwarp();
      ptx::mbar_wait(&v_empty[s], ((j >> 1) & 1) ^ 1);
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(&v_full[s], C::KV_BYTES);
#pragma unroll
        for (int a = 0; a < C::ATOMS; ++a)
          ptx::tma_load_3d(sV + s * C::KV_BYTES + a * (BKV * 128), &p.tv, &v_full[s], h * HD + a * 64, j * BKV, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(BQ, BKV, 0, 0);
    constexpr uint32_t idesc_o = ptx::umma_idesc_bf16(BQ, HD, 0, 1);  // B = V is MN-major
    const uint32_t qbase = ptx::smem_u32(sQ);
    auto issue_s = [&](int g, int s) {   // S_g = Q_g K^T into columns [g*128, g*128+128)
      const uint32_t kb = ptx::smem_u32(sK + s * C::KV_BYTES);
#pragma unroll
      for (int k = 0; k < HD / 16; ++k) {
        const uint32_t qoff = (k >> 2) * (2 * BQ * 128) + g * (BQ * 128) + (k & 3) * 32;
        const uint32_t koff = (k >> 2) * (BKV * 128) + (k & 3) * 32;
        ptx::umma_f16_ss(tmem + g * 128, ptx::umma_smem_desc(qbase + qoff, 16, 1024),
                         ptx::umma_smem_desc(kb + koff, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
      }
      ptx::umma_commit(&s_full[g]);
    };
    auto issue_pv = [&](int g, int s, int j, int half) {   // O_g (+)= P_g[:, half] V[half, :]
      const uint32_t vb = ptx::smem_u32(sV + s * C::KV_BYTES);
#pragma unroll
      for (int k = half * (HB / 16); k < (half + 1) * (HB / 16); ++k)
        ptx::umma_f16_ts(tmem + 256 + g * 128, tmem + g * 128 + k * 8,
                         ptx::umma_smem_desc(vb + k * 2048, BKV * 128, 1024), idesc_o, (j | k) != 0 ? 1u : 0u);
    };
    ptx::mbar_wait(q_full, 0);
    ptx::mbar_wait(&k_full[0], 0);
    ptx::tc_fence_after();
    if (ptx::elect_one()) {
      if (n_kv[0] > 0) issue_s(0, 0);
      if (n_kv[1] > 0) issue_s(1, 0);
      ptx::umma_commit(&k_empty[0]);
    }
    __syncwarp();
    for (int j = 0; j < n_tot; ++j) {
      const int s = j & 1, sn = (j + 1) & 1;
      ptx::mbar_wait(&v_full[s], (j >> 1) & 1);
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if (j < n_kv[g]) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            ptx::mbar_wait(&p_full[g * 2 + half], j & 1);
            probe(2, j, g * 3 + half);
            ptx::tc_fence_after();
            if (ptx::elect_one()) issue_pv(g, s, j, half);
            __syncwarp();
          }
          if (j + 1 < n_kv[g]) {
            ptx::mbar_wait(&k_full[sn], ((j + 1) >> 1) & 1);
            ptx::tc_fence_after();
            if (ptx::elect_one()) issue_s(g, sn);
          } else if (ptx::elect_one()) {
            ptx::umma_commit(&o_done[g]);
          }
          __syncwarp();
          probe(2, j, g * 3 + 2);
        }
      }
      if (ptx::elect_one()) {
        ptx::umma_commit(&v_empty[s]);
        if (j + 1 < n_tot) ptx::umma_commit(&k_empty[sn]);
      }
      __syncwarp();
    }
  } else {
    // ---------------------------------------------------------------- softmax groups
    const int g = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int qg0 = q0 + g * BQ;
    const int q = qg0 + row;
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t colS = g * 128, colO = 256 + g * 128;
    const int n_mine = n_kv[g];
    float m_ref = -INFINITY, l_sum = 0.f;
    const uint32_t never = static_cast<uint32_t>(pp.src_len >> 31);   // 0 at run time (src_len >= 0), opaque to ptxas
    for (int j = 0; j < n_mine; ++j) {
      const int kv0 = j * BKV;
      ptx::mbar_wait(&s_full[g], j & 1);
      ptx::tc_fence_after();
      if (quarter == 2) probe(g, j, 0);
      const bool need_mask = (kv0 + BKV - 1 > qg0 + pp.src_len) || (kv0 + BKV > pp.T);
      // the whole score row of this tile in registers: ONE TMEM round trip (four loads in flight)
      uint32_t r[BKV];
#pragma unroll
      for (int c = 0; c < BKV / 32; ++c) ptx::tmem_ld_32x32p(tmem + lane_base + colS + c * 32, r + c * 32);
      ptx::tmem_ld_wait();
      if (quarter == 2) probe(g, j, 1);
      if (need_mask) {
#pragma unroll
        for (int e = 0; e < BKV; ++e) {
          const int kk = kv0 + e;
          if (kk > q + pp.src_len || kk >= pp.T) r[e] = 0xff800000u;  // -inf
        }
      }
      // max of the RAW scores (scale > 0 commutes with max): four independent chains
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int e = 0; e < BKV; e += 8) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          mx4[u] = fmaxf(mx4[u], fmaxf(__uint_as_float(r[e + 2 * u]), __uint_as_float(r[e + 2 * u + 1])));
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      const float m_cand = fmaxf(m_ref, mx * pp.scale_log2);
      const bool grow = (m_cand > m_ref + 8.0f) || (m_ref == -INFINITY && m_cand != -INFINITY);
      if (__any_sync(0xffffffffu, grow)) {
        // S_g(j) complete => PV_g(j-1) complete (issue order): O_g may be rescaled right here
        const float alpha = (m_ref == -INFINITY) ? 0.f : ptx::ex2(m_ref - m_cand);
        l_sum *= alpha;
        m_ref = m_cand;
        if (j > 0) {
          // rare path: 16 columns at a time so that it does not push the score row out of the registers
#pragma unroll 1
          for (int c = 0; c < HD / 16; ++c) {
            uint32_t o[16];
            ptx::tmem_ld_32x16p(tmem + lane_base + colO + c * 16, o);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
            ptx::tmem_st_32x16p(tmem + lane_base + colO + c * 16, o);
          }
        }
      }
      // p = 2^(s * scale_log2 - m): one FFMA + one MUFU per score; packed bf16 pairs overwrite S
      const float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;
      float l0 = 0.f, l1 = 0.f;
      if (quarter == 2) probe(g, j, 2);
      // dropout acts on the NORMALISED probabilities (base_blocks.py:193-194): the row sum stays undropped,
      // only the P that feeds P.V is masked and rescaled
      const unsigned long long drop_row = DROP ? ((static_cast<unsigned long long>(b) * pp.n_heads + h) * pp.T + q) *
                                                     static_cast<unsigned long long>((pp.T + 1) & ~1) + kv0 : 0ull;
      float nm = neg_m;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        // (nm carries a data dependence on the first half's arrival: ptxas would otherwise hoist all 128
        //  exponentials above the first hand-off)
        if (half == 1) {
          // registers: the CTA's 10 warps are allocated as 12, which caps a thread at 168 registers; the second
          // half of the score row is dropped after the max and re-read here (one more ~60-cycle TMEM load)
#pragma unroll
          for (int c = HB / 32; c < BKV / 32; ++c) ptx::tmem_ld_32x32p(tmem + lane_base + colS + c * 32, r + c * 32);
          ptx::tmem_ld_wait();
          if (need_mask) {
#pragma unroll
            for (int e = HB; e < BKV; ++e) {
              const int kk = kv0 + e;
              if (kk > q + pp.src_len || kk >= pp.T) r[e] = 0xff800000u;
            }
          }
        }
#pragma unroll
        for (int e = half * HB; e < (half + 1) * HB; e += 2) {
          float p0 = ptx::ex2(fmaf(__uint_as_float(r[e]), pp.scale_log2, nm));
          float p1 = ptx::ex2(fmaf(__uint_as_float(r[e + 1]), pp.scale_log2, nm));
          l0 += p0; l1 += p1;
          if (DROP) {
            const uint2 hsh = ptx::drop_hash(pp.drop_seed, p.drop_site, (drop_row + e) >> 1);
            p0 = hsh.x >= pp.drop_thresh ? p0 * pp.drop_scale : 0.f;
            p1 = hsh.y >= pp.drop_thresh ? p1 * pp.drop_scale : 0.f;
          }
          r[e >> 1] = ptx::pack_bf16(p0, p1);
        }
        // hand this half of P to the pipe: P.V on it overlaps the exponentials of the other half
#pragma unroll
        for (int c = 0; c < HB / 32; ++c)
          ptx::tmem_st_32x16p(tmem + lane_base + colS + half * (HB / 2) + c * 16, r + half * (HB / 2) + c * 16);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        if (half == 0) {
          const uint32_t tok = ptx::mbar_arrive_token(&p_full[g * 2]);
          nm = __uint_as_float(__float_as_uint(nm) ^ (tok & never));
        } else {
          ptx::mbar_arrive(&p_full[g * 2 + 1]);
        }
        if (quarter == 2) probe(g, j, 3 + half);
      }
      l_sum += l0 + l1;
    }
    if (n_mine > 0) {
      ptx::mbar_wait(&o_done[g], 0);
      ptx::tc_fence_after();
      const float inv = 1.f / l_sum;
      __nv_bfloat16* orow = p.o + (static_cast<long long>(b) * pp.T + q) * pp.ldo + h * HD;
#pragma unroll
      for (int c = 0; c < HD / 32; ++c) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem + lane_base + colO + c * 32, r);
        ptx::tmem_ld_wait();
        if (q < pp.T) {
#pragma unroll
          for (int e = 0; e < 32; e += 8) {
            uint4 o;
            o.x = ptx::pack_bf16(__uint_as_float(r[e]) * inv, __uint_as_float(r[e + 1]) * inv);
            o.y = ptx::pack_bf16(__uint_as_float(r[e + 2]) * inv, __uint_as_float(r[e + 3]) * inv);
            o.z = ptx::pack_bf16(__uint_as_float(r[e + 4]) * inv, __uint_as_float(r[e + 5]) * inv);
            o.w = ptx::pack_bf16(__uint_as_float(r[e + 6]) * inv, __uint_as_float(r[e + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + c * 32 + e) = o;
          }
        }
      }
      if (p.lse != nullptr && q < pp.T)
        p.lse[(static_cast<long long>(b) * pp.n_heads + h) * pp.T + q] =
            (m_ref + log2f(l_sum)) * 0.69314718055994530942f;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

long long* g_attn_trace = nullptr;

template <int HD, bool DROP>
int launch_tc2(int n, const sea_attn_args* a, cudaStream_t s) {
  using C = ACfg2<HD>;
  static bool attr_set[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 16 && !attr_set[dev]) {
    SEA_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc2_kernel<HD, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr_set[dev] = true;
  }
  AttnTcParams p;
  const uint64_t wq = static_cast<uint64_t>(a->n_heads) * HD;
  for (int i = 0; i < n; ++i) {
    const sea_attn_args& x = a[i];
    int rc = make_tmap_bf16_3d(&p.it[i].tq, x.q, wq, x.T, x.B, x.ldq, x.ldq * static_cast<uint64_t>(x.T), 64, BQ);
    if (rc) return rc;
    rc = make_tmap_bf16_3d(&p.it[i].tk, x.k, wq, x.T, x.B, x.ldk, x.ldk * static_cast<uint64_t>(x.T), 64, C::BKV);
    if (rc) return rc;
    rc = make_tmap_bf16_3d(&p.it[i].tv, x.v, wq, x.T, x.B, x.ldv, x.ldv * static_cast<uint64_t>(x.T), 64, C::BKV);
    if (rc) return rc;
    p.it[i].o = static_cast<__nv_bfloat16*>(x.o);
    p.it[i].lse = x.lse;
    p.it[i].drop_site = x.dropout_site;
  }
  p.drop_seed = a->dropout_seed;
  p.drop_thresh = a->dropout_p > 0.f ? static_cast<uint32_t>(static_cast<double>(a->dropout_p) * 4294967296.0) : 0u;
  p.drop_scale = 1.0f / (1.0f - a->dropout_p);
  p.ldo = a->ldo;
  p.B = a->B; p.T = a->T; p.n_heads = a->n_heads; p.src_len = a->src_len;
  p.scale_log2 = a->scale * 1.44269504088896340736f;
  dim3 grid(a->n_heads * a->B * n, 1, (a->T + 2 * BQ - 1) / (2 * BQ));
  p.trace = g_attn_trace;
  if constexpr (HD == 128 && !DROP) {
    if (g_attn_trace != nullptr) {
      SEA_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc2_kernel<HD, DROP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
      SEA_LAUNCH((attn_fwd_tc2_kernel<HD, DROP, true>), grid, kThreads2, C::SMEM, s, p);
      return static_cast<int>(cudaGetLastError());
    }
  }
  SEA_LAUNCH((attn_fwd_tc2_kernel<HD, DROP>), grid, kThreads2, C::SMEM, s, p);
  return static_cast<int>(cudaGetLastError());
}

template <int HD, int BKV, int STAGES_, bool DROP>
int launch_tc(int n, const sea_attn_args* a, cudaStream_t s) {
  using C = ACfg<HD, BKV, STAGES_>;
  static bool attr_set[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 16 && !attr_set[dev]) {
    SEA_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc_kernel<HD, BKV, STAGES_, DROP>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr_set[dev] = true;
  }
  AttnTcParams p;
  const uint64_t wq = static_cast<uint64_t>(a->n_heads) * HD;
  for (int i = 0; i < n; ++i) {
    const sea_attn_args& x = a[i];
    int rc = make_tmap_bf16_3d(&p.it[i].tq, x.q, wq, x.T, x.B, x.ldq, x.ldq * static_cast<uint64_t>(x.T), 64, BQ);
    if (rc) return rc;
    rc = make_tmap_bf16_3d(&p.it[i].tk, x.k, wq, x.T, x.B, x.ldk, x.ldk * static_cast<uint64_t>(x.T), 64, BKV);
    if (rc) return rc;
    rc = make_tmap_bf16_3d(&p.it[i].tv, x.v, wq, x.T, x.B, x.ldv, x.ldv * static_cast<uint64_t>(x.T), 64, BKV);
    if (rc) return rc;
    p.it[i].o = static_cast<__nv_bfloat16*>(x.o);
    p.it[i].lse = x.lse;
    p.it[i].drop_site = x.dropout_site;
  }
  p.drop_seed = a->dropout_seed;
  p.drop_thresh = a->dropout_p > 0.f ? static_cast<uint32_t>(static_cast<double>(a->dropout_p) * 4294967296.0) : 0u;
  p.drop_scale = 1.0f / (1.0f - a->dropout_p);
  p.ldo = a->ldo;
  p.B = a->B; p.T = a->T; p.n_heads = a->n_heads; p.src_len = a->src_len;
  p.scale_log2 = a->scale * 1.44269504088896340736f;
  dim3 grid(a->n_heads * a->B * n, 1, (a->T + BQ - 1) / BQ);
  SEA_LAUNCH((attn_fwd_tc_kernel<HD, BKV, STAGES_, DROP>), grid, kThreads, C::SMEM, s, p);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace

bool attention_tc_supported(const sea_attn_args* a) {
  if (a->prec != SEA_PREC_BF16) return false;
  if (a->head_dim != 64 && a->head_dim != 128 && a->head_dim != 256) return false;
  if (a->src_len < 0) return false;
  // TMA: 16-byte aligned bases and row pitches
  if ((a->ldq % 8) || (a->ldk % 8) || (a->ldv % 8) || (a->ldo % 8)) return false;
  if ((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) |
       reinterpret_cast<uintptr_t>(a->v) | reinterpret_cast<uintptr_t>(a->o)) & 15)
    return false;
  return true;
}

void attention_set_trace(void* dev_buf) { g_attn_trace = static_cast<long long*>(dev_buf); }

int g_attn_two_tiles = 1;  // tuning hook: 0 = one query tile per CTA also for long sequences

// n same-shape problems (B, T, n_heads, head_dim, src_len, scale, ldo equal; checked by the caller)
int attention_fwd_tc(int n, const sea_attn_args* a, cudaStream_t s) {
  int rc = ensure_init();
  if (rc) return rc;
  const int k_last = (a->T - 1 + a->src_len < a->T - 1) ? a->T - 1 + a->src_len : a->T - 1;
  const bool drop = a->dropout_p > 0.f;
#define SEA_ATTN_FWD(HD_, BKV_)                                                                     \
  do {                                                                                              \
    if (k_last < BKV_)                                                                              \
      return drop ? launch_tc<HD_, BKV_, 1, true>(n, a, s) : launch_tc<HD_, BKV_, 1, false>(n, a, s); \
    if (HD_ <= 128 && g_attn_two_tiles)                                                             \
      return drop ? launch_tc2<(HD_ <= 128 ? HD_ : 128), true>(n, a, s)                             \
                  : launch_tc2<(HD_ <= 128 ? HD_ : 128), false>(n, a, s);                           \
    return drop ? launch_tc<HD_, BKV_, 2, true>(n, a, s) : launch_tc<HD_, BKV_, 2, false>(n, a, s); \
  } while (0)
  switch (a->head_dim) {
    case 64: SEA_ATTN_FWD(64, 128);
    case 128: SEA_ATTN_FWD(128, 128);
    case 256: SEA_ATTN_FWD(256, 64);
    default: return SEA_ERR_UNSUPPORTED;
  }
#undef SEA_ATTN_FWD
}

}  // namespace sea

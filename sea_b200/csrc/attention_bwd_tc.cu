// K3 (tcgen05 variant) — backward of the fused causal attention on the 5th-gen tensor cores.
// Flash-style recomputation from (q, k, v, lse): nothing of size T x T is stored
// (the reference materialises softmax(QK^T) through autograd, models/base_blocks.py:191-197, :283-289).
//
// One templated kernel, two roles (no atomics, deterministic):
//   MODE 0 (dQ)     CTA = 128 queries of one (b, h); (Q_i, dO_i) stay in smem, (K_j, V_j) stream.
//                     S  = Q_i K_j^T, dP = dO_i V_j^T        (SS MMAs, fp32 in TMEM)
//                     dS = P o (dP - delta) * scale          (registers; bf16 back into TMEM)
//                     dQ_i += dS K_j                         (TS MMA, K_j MN-major)
//   MODE 1 (dK,dV)  CTA = 128 keys of one (b, h) [x one 128-wide half of the head dim when hd = 256];
//                   (K_j, V_j) stay, (Q_i, dO_i) stream.  Everything is transposed so keys sit on
//                   the TMEM lanes:  S^T = K_j Q_i^T, dP^T = V_j dO_i^T,
//                     dV_j += P^T dO_i ,  dK_j += dS^T Q_i   (TS MMAs, operands MN-major)
// Roles inside a CTA (320 threads): warp 0 TMA, warp 1 tcgen05.mma issue + TMEM, warps 2..9 = two compute
// warpgroups; a thread owns one TMEM lane (= one query / key row) and HALF of the streamed columns of a tile.
// Streamed tiles are 64 wide and S|dP are DOUBLE-BUFFERED in TMEM: the S / dP MMAs of tile it+1 are issued before
// the accumulate MMAs of tile it, so the tensor pipe recomputes the next scores while the compute warps turn the
// current ones into P / dS (exp2, mask, dropout, (dP - delta) * scale) — the pipe no longer idles during the
// softmax-backward arithmetic, nor the compute warps during the MMAs.  P / dS overwrite the S / dP columns in
// place (tcgen05.mma executes in issue order, so a later tile's S cannot overtake their consumer).
// The RoPE of the forward (fused in the projection GEMM epilogue) is undone on dQ / dK in the
// epilogue: rotating the gradient by -theta is the transpose of the forward rotation.
// TMEM columns: buffer b in {0,1}: S|P [128b, 128b+64)  dP|dS [128b+64, 128b+128);  acc0 [256,..)  acc1 [384,512).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/sea_b200.h"
#include "internal.h"
#include "ptx.cuh"

namespace sea {
int g_attn_bwd_probe = 0;
int g_attn_bwd_wide = 1;   // 1: 128-wide streamed tiles with two-half hand-over (HD <= 128); 0: 64-wide double-buffered plan
namespace {

constexpr int BR = 128;  // stationary rows per CTA (TMEM lanes)
constexpr int kComputeWarps = 8;
constexpr int kThreads = 64 + 32 * kComputeWarps;
constexpr uint32_t kColBuf = 128, kColAcc0 = 256, kColAcc1 = 384;   // S at buffer + 0, dP at buffer + BS
constexpr float kLog2e = 1.44269504088896340736f;

struct alignas(64) AttnBwdTcParams {
  CUtensorMap ta1, ta2, tb1, tb2;   // stationary pair (box 128 rows), streamed pair (box BS rows)
  __nv_bfloat16 *out0, *out1;       // MODE 0: dq, -   MODE 1: dv, dk
  long long ld0, ld1;
  const float *lse, *delta, *rope;
  int B, T, n_heads, src_len, rope_ld;
  float scale, scale_log2;
  unsigned long long drop_seed;   // probability dropout of the forward, regenerated here (DROP kernels)
  uint32_t drop_thresh, drop_site;
  float drop_scale;
  int probe;   // tuning probe (sea_attention_bwd_probe): 1 = compute warps skip TMEM traffic and arithmetic, 2 = skip arithmetic only
};

template <int HD, int BS, int MODE>
struct BCfg {
  static constexpr int ATOMS = HD / 64;
  static constexpr int DH = (MODE == 1 && HD > 128) ? 128 : HD;  // accumulator width per CTA
  static constexpr int HALVES = HD / DH;
  static constexpr int STAT_BYTES = BR * HD * 2;
  static constexpr int STR_BYTES = BS * HD * 2;
  static constexpr int FIT = (227 * 1024 - 4096 - 2 * STAT_BYTES) / (2 * STR_BYTES);
  static constexpr int STAGES = FIT >= 4 ? 4 : (FIT >= 3 ? 3 : (FIT >= 2 ? 2 : 1));
  static constexpr bool PIPE = STAGES >= 2;   // S/dP of tile it+1 ahead of the accumulate MMAs of tile it
  static constexpr int SMEM = 2 * STAT_BYTES + STAGES * 2 * STR_BYTES + 1024 + 512 + 4 * BS * 4;
  static_assert(BS == 64, "TMEM plan: two S|dP buffers of 2 x 64 columns");
};

template <int HD, int BS, int MODE, bool DROP>
__global__ void __launch_bounds__(kThreads, 1) attn_bwd_tc_kernel(const __grid_constant__ AttnBwdTcParams p) {
  using C = BCfg<HD, BS, MODE>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA1 = smem;
  uint8_t* sA2 = sA1 + C::STAT_BYTES;
  uint8_t* sB1 = sA2 + C::STAT_BYTES;                  // [STAGES][STR_BYTES]
  uint8_t* sB2 = sB1 + STAGES * C::STR_BYTES;          // [STAGES][STR_BYTES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB2 + STAGES * C::STR_BYTES);
  uint64_t* a_full = bars;         // 1
  uint64_t* b_full = bars + 1;     // [4]
  uint64_t* b_empty = bars + 5;    // [4]
  uint64_t* s_full = bars + 9;     // [2] MMA -> compute, one per TMEM buffer
  uint64_t* p_full = bars + 11;    // [2] compute (256 arrivals) -> MMA
  uint64_t* acc_done = bars + 13;  // MMA -> epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  float* lse_s = reinterpret_cast<float*>(bars + 64);  // [2][BS]  (MODE 1: per-column lse*log2e)
  float* dl_s = lse_s + 2 * BS;                        // [2][BS]  (MODE 1: per-column delta)

  // warp-uniform role index + elect.sync regions (see gemm.cu): no waterfall loops around UTMALDG / UTCHMMA
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  // blockIdx.x = (batch, head), blockIdx.z = tile: the heaviest tiles of ALL (batch, head) pairs launch first
  const int h = blockIdx.x % p.n_heads, b = blockIdx.x / p.n_heads;
  int tile, half = 0;
  if (MODE == 0) {
    tile = gridDim.z - 1 - blockIdx.z;  // late query tiles see the most keys: schedule them first
  } else {
    tile = blockIdx.z / C::HALVES;      // early key tiles are seen by the most queries
    half = blockIdx.z % C::HALVES;
  }
  const int r0 = tile * BR;
  // streamed tile range [it0, it0 + n_it)
  int it0, n_it;
  if (MODE == 0) {
    const int q_hi = min(p.T - 1, r0 + BR - 1);
    const int k_last = min(p.T - 1, q_hi + p.src_len);
    it0 = 0;
    n_it = k_last / BS + 1;
  } else {
    const int q_first = max(0, r0 - p.src_len);
    it0 = q_first / BS;
    n_it = (p.T - 1) / BS - it0 + 1;
  }

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&p.ta1);
    ptx::prefetch_tmap(&p.ta2);
    ptx::prefetch_tmap(&p.tb1);
    ptx::prefetch_tmap(&p.tb2);
    ptx::mbar_init(a_full, 1);
    for (int s = 0; s < 4; ++s) {
      ptx::mbar_init(&b_full[s], 1);
      ptx::mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&s_full[s], 1);
      ptx::mbar_init(&p_full[s], 32 * kComputeWarps);
    }
    ptx::mbar_init(acc_done, 1);
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
      ptx::mbar_expect_tx(a_full, 2 * C::STAT_BYTES);
#pragma unroll
      for (int a = 0; a < C::ATOMS; ++a) {
        ptx::tma_load_3d(sA1 + a * (BR * 128), &p.ta1, a_full, h * HD + a * 64, r0, b);
        ptx::tma_load_3d(sA2 + a * (BR * 128), &p.ta2, a_full, h * HD + a * 64, r0, b);
      }
    }
    __syncwarp();
    for (int it = 0; it < n_it; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      ptx::mbar_wait(&b_empty[s], ph ^ 1);
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(&b_full[s], 2 * C::STR_BYTES);
        const int row = (it0 + it) * BS;
#pragma unroll
        for (int a = 0; a < C::ATOMS; ++a) {
          ptx::tma_load_3d(sB1 + s * C::STR_BYTES + a * (BS * 128), &p.tb1, &b_full[s], h * HD + a * 64, row, b);
          ptx::tma_load_3d(sB2 + s * C::STR_BYTES + a * (BS * 128), &p.tb2, &b_full[s], h * HD + a * 64, row, b);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(BR, BS, 0, 0);
    constexpr uint32_t idesc_acc = ptx::umma_idesc_bf16(BR, C::DH, 0, 1);  // B operand MN-major
    ptx::mbar_wait(a_full, 0);
    const uint32_t a1 = ptx::smem_u32(sA1), a2 = ptx::smem_u32(sA2);
    // S = A1 . B1^T and dP = A2 . B2^T of streamed tile `it` into TMEM buffer it & 1
    auto issue_scores = [&](int it) {
      const int s = it % STAGES;
      ptx::mbar_wait(&b_full[s], (it / STAGES) & 1);
      ptx::tc_fence_after();
      const uint32_t b1 = ptx::smem_u32(sB1 + s * C::STR_BYTES), b2 = ptx::smem_u32(sB2 + s * C::STR_BYTES);
      const uint32_t buf = tmem + (it & 1) * kColBuf;
      if (ptx::elect_one()) {
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          const uint32_t aoff = (k >> 2) * (BR * 128) + (k & 3) * 32;
          const uint32_t boff = (k >> 2) * (BS * 128) + (k & 3) * 32;
          ptx::umma_f16_ss(buf, ptx::umma_smem_desc(a1 + aoff, 16, 1024),
                           ptx::umma_smem_desc(b1 + boff, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
        }
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          const uint32_t aoff = (k >> 2) * (BR * 128) + (k & 3) * 32;
          const uint32_t boff = (k >> 2) * (BS * 128) + (k & 3) * 32;
          ptx::umma_f16_ss(buf + BS, ptx::umma_smem_desc(a2 + aoff, 16, 1024),
                           ptx::umma_smem_desc(b2 + boff, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&s_full[it & 1]);
      }
      __syncwarp();
    };
    // dQ += dS K   /   dV += P^T dO, dK += dS^T Q   from the P | dS the compute warps left in buffer it & 1
    auto issue_accumulate = [&](int it) {
      const int s = it % STAGES;
      const uint32_t b1 = ptx::smem_u32(sB1 + s * C::STR_BYTES), b2 = ptx::smem_u32(sB2 + s * C::STR_BYTES);
      const uint32_t buf = tmem + (it & 1) * kColBuf;
      ptx::mbar_wait(&p_full[it & 1], (it >> 1) & 1);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t hoff = half * (C::DH / 64) * (BS * 128);
#pragma unroll
        for (int k = 0; k < BS / 16; ++k) {
          // 16 streamed rows per step = two 8-row groups (SBO 1024 B); 64-wide chunks LBO apart
          const uint32_t acc = (it | k) != 0 ? 1u : 0u;
          const uint32_t pk = (k >> 1) * 32 + (k & 1) * 8;   // 16 streamed columns = 8 packed TMEM columns, per 32-column half
          if (MODE == 0) {
            ptx::umma_f16_ts(tmem + kColAcc0, buf + BS + pk,
                             ptx::umma_smem_desc(b1 + k * 2048, BS * 128, 1024), idesc_acc, acc);
          } else {
            ptx::umma_f16_ts(tmem + kColAcc0, buf + pk,
                             ptx::umma_smem_desc(b2 + hoff + k * 2048, BS * 128, 1024), idesc_acc, acc);
            ptx::umma_f16_ts(tmem + kColAcc1, buf + BS + pk,
                             ptx::umma_smem_desc(b1 + hoff + k * 2048, BS * 128, 1024), idesc_acc, acc);
          }
        }
        ptx::umma_commit(&b_empty[s]);
        if (it == n_it - 1) ptx::umma_commit(acc_done);
      }
      __syncwarp();
    };
    if (C::PIPE) {
      issue_scores(0);
      for (int it = 0; it < n_it; ++it) {
        if (it + 1 < n_it) issue_scores(it + 1);
        issue_accumulate(it);
      }
    } else {
      for (int it = 0; it < n_it; ++it) {
        issue_scores(it);
        issue_accumulate(it);
      }
    }
  } else {
    // ---------------------------------------------------------------- compute warps
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int chalf = (warp - 2) >> 2;   // which 32 of a tile's 64 streamed columns this thread handles
    const int tid = threadIdx.x - 64;    // 0..255
    const int rpos = r0 + row;         // query (MODE 0) or key (MODE 1) position of this thread
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const long long stat_base = (static_cast<long long>(b) * p.n_heads + h) * p.T;
    float my_lse2 = 0.f, my_dl = 0.f;
    if (MODE == 0 && rpos < p.T) {
      my_lse2 = p.lse[stat_base + rpos] * kLog2e;
      my_dl = p.delta[stat_base + rpos];
    }
    for (int it = 0; it < n_it; ++it) {
      const int c0 = (it0 + it) * BS;  // first streamed position (keys in MODE 0, queries in MODE 1)
      if (MODE == 1) {
        if (tid < BS) {
          const int q = c0 + tid;
          lse_s[(it & 1) * BS + tid] = q < p.T ? p.lse[stat_base + q] * kLog2e : 0.f;
          dl_s[(it & 1) * BS + tid] = q < p.T ? p.delta[stat_base + q] : 0.f;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      bool need_mask;
      if (MODE == 0) need_mask = (c0 + BS - 1 > r0 + p.src_len) || (c0 + BS > p.T) || (r0 + BR > p.T);
      else need_mask = (r0 + BR - 1 > c0 + p.src_len) || (c0 + BS > p.T) || (r0 + BR > p.T);
      const float* lrow = lse_s + (it & 1) * BS;
      const float* drow = dl_s + (it & 1) * BS;
      ptx::mbar_wait(&s_full[it & 1], (it >> 1) & 1);
      ptx::tc_fence_after();
      // this thread's 32 columns of S and dP in one TMEM round trip, results written back in place
      if (p.probe != 1) {
        const uint32_t buf = tmem + lane_base + (it & 1) * kColBuf;
        const int cb = chalf * 32;
        uint32_t rs[32], rd[32];
        ptx::tmem_ld_32x32p(buf + cb, rs);
        ptx::tmem_ld_32x32p(buf + BS + cb, rd);
        ptx::tmem_ld_wait();
        if (p.probe != 2)
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          float pv[2], dv[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int cpos = c0 + cb + e + u;
            float l2, dl;
            if (MODE == 0) { l2 = my_lse2; dl = my_dl; }
            else { l2 = lrow[cb + e + u]; dl = drow[cb + e + u]; }
            float pe = ptx::ex2(fmaf(__uint_as_float(rs[e + u]), p.scale_log2, -l2));
            if (need_mask) {
              const int qq = MODE == 0 ? rpos : cpos;
              const int kk = MODE == 0 ? cpos : rpos;
              if (kk > qq + p.src_len || kk >= p.T || qq >= p.T) pe = 0.f;
            }
            float dpv = __uint_as_float(rd[e + u]);
            if (DROP) {
              // O = (mask/(1-p) o P) V: dV uses the dropped P, dP picks up the same factor; delta = rowsum(dO o O)
              // already equals sum_k P_k dP_k, so dS = P o (dP - delta) keeps the UNdropped P in front
              const int qq = MODE == 0 ? rpos : cpos;
              const int kk = MODE == 0 ? cpos : rpos;
              const float mult = ptx::drop_mult(p.drop_seed, p.drop_site,
                                                ((static_cast<unsigned long long>(b) * p.n_heads + h) * p.T + qq) *
                                                        static_cast<unsigned long long>((p.T + 1) & ~1) + kk,
                                                p.drop_thresh, p.drop_scale);
              dpv *= mult;
              pv[u] = pe * mult;
            } else {
              pv[u] = pe;
            }
            dv[u] = pe * (dpv - dl) * p.scale;
          }
          rs[e >> 1] = ptx::pack_bf16(pv[0], pv[1]);
          rd[e >> 1] = ptx::pack_bf16(dv[0], dv[1]);
        }
        // packed bf16 results go to the front of the columns THIS thread has just read (the other warpgroup may
        // still be loading its half): P at S + 32 * chalf, dS at dP + 32 * chalf, 16 columns each
        if (MODE == 1) ptx::tmem_st_32x16p(buf + cb, rs);
        ptx::tmem_st_32x16p(buf + BS + cb, rd);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&p_full[it & 1]);
    }
    // epilogue: accumulators -> (un-RoPE) -> bf16 rows
    ptx::mbar_wait(acc_done, 0);
    ptx::tc_fence_after();
    const long long orow = static_cast<long long>(b) * p.T + rpos;
    const int tpos = min(rpos, p.T - 1);
#pragma unroll
    for (int which = 0; which < (MODE == 0 ? 1 : 2); ++which) {
      const bool rot = p.rope != nullptr && (MODE == 0 || which == 1);
      __nv_bfloat16* out = (which == 0 ? p.out0 : p.out1) + orow * (which == 0 ? p.ld0 : p.ld1) + h * HD + half * C::DH;
      const uint32_t col = which == 0 ? kColAcc0 : kColAcc1;
#pragma unroll
      for (int cc = 0; cc < C::DH / 64; ++cc) {
        const int c = cc * 2 + chalf;     // the two compute warpgroups take alternate 32-column slabs
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem + lane_base + col + c * 32, r);
        ptx::tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r[e]);
        if (rot) {
          const float2* tab = reinterpret_cast<const float2*>(p.rope) +
                              static_cast<long long>((half * C::DH + c * 32) >> 1) * p.rope_ld + tpos;
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float2 cs = __ldg(tab + static_cast<long long>(e >> 1) * p.rope_ld);
            const float x0 = v[e], x1 = v[e + 1];
            v[e] = x0 * cs.x + x1 * cs.y;
            v[e + 1] = x1 * cs.x - x0 * cs.y;
          }
        }
        if (rpos < p.T) {
#pragma unroll
          for (int e = 0; e < 32; e += 8) {
            uint4 o;
            o.x = ptx::pack_bf16(v[e], v[e + 1]);
            o.y = ptx::pack_bf16(v[e + 2], v[e + 3]);
            o.z = ptx::pack_bf16(v[e + 4], v[e + 5]);
            o.w = ptx::pack_bf16(v[e + 6], v[e + 7]);
            *reinterpret_cast<uint4*>(out + c * 32 + e) = o;
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// 128-wide variant (HD <= 128).  tcgen05.mma with N = 64 costs ~75 cycles against 64 for N = 128 (the A operand is
// re-read per instruction), so the 64-wide plan above runs S and dP at half the pipe's rate.  Here S and dP are full
// 128-column MMAs again, and the overlap comes from elsewhere:
//   * the compute warps hand P | dS over in two 64-column halves: the accumulate MMAs of the first half run while the
//     second half is still being computed (K = 64 each; their N is the head dim, i.e. full width);
//   * MODE 0 (dQ) has 128 spare TMEM columns: S is double-buffered, so S(it+1) is computed while the compute warps work
//     on tile it; dP(it+1) follows the accumulate MMAs of tile it (it reuses the dS columns).
// TMEM: MODE 0: S0 [0,128) S1 [128,256) dP|dS [256,384) dQ [384,384+HD);  MODE 1: S|P [0,128) dP|dS [128,256)
// dV [256,256+HD) dK [384,384+HD).  Two operand stages of 128 streamed rows (192 KB at HD = 128).
template <int HD, int MODE>
struct B2Cfg {
  static constexpr int BS = 128;
  static constexpr int COMPUTE_WARPS = 16;            // four per TMEM lane quarter: a thread owns one row and 16 of a half's 64 columns
  static constexpr int THREADS = 64 + 32 * COMPUTE_WARPS;
  static constexpr int ATOMS = HD / 64;
  static constexpr int STAT_BYTES = BR * HD * 2;
  static constexpr int STR_BYTES = BS * HD * 2;
  static constexpr int STAGES = 2;
  static constexpr int SMEM = 2 * STAT_BYTES + STAGES * 2 * STR_BYTES + 1024 + 512 + 4 * BS * 4;
  static constexpr uint32_t COL_S = 0, COL_DP = MODE == 0 ? 256 : 128, COL_ACC0 = MODE == 0 ? 384 : 256, COL_ACC1 = 384;
  static_assert(HD <= 128, "accumulators of 128 columns");
};

template <int HD, int MODE, bool DROP>
__global__ void __launch_bounds__((B2Cfg<HD, MODE>::THREADS), 1) attn_bwd_tc2_kernel(const __grid_constant__ AttnBwdTcParams p) {
  using C = B2Cfg<HD, MODE>;
  constexpr int BS = C::BS, STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA1 = smem;
  uint8_t* sA2 = sA1 + C::STAT_BYTES;
  uint8_t* sB1 = sA2 + C::STAT_BYTES;                  // [STAGES][STR_BYTES]
  uint8_t* sB2 = sB1 + STAGES * C::STR_BYTES;          // [STAGES][STR_BYTES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB2 + STAGES * C::STR_BYTES);
  uint64_t* a_full = bars;         // 1
  uint64_t* b_full = bars + 1;     // [2]
  uint64_t* b_empty = bars + 3;    // [2]
  uint64_t* s_full = bars + 5;     // [2] MMA -> compute (MODE 0: one per S buffer; MODE 1: [0] covers S and dP)
  uint64_t* dp_full = bars + 7;    // MODE 0: dP of the tile is complete
  uint64_t* p_half = bars + 8;     // [2] compute (256 arrivals) -> MMA: half h of P | dS is in TMEM
  uint64_t* acc_done = bars + 10;  // MMA -> epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);
  float* lse_s = reinterpret_cast<float*>(bars + 64);  // [2][BS]  (MODE 1: per-column lse*log2e)
  float* dl_s = lse_s + 2 * BS;                        // [2][BS]  (MODE 1: per-column delta)

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int h = blockIdx.x % p.n_heads, b = blockIdx.x / p.n_heads;
  const int tile = MODE == 0 ? static_cast<int>(gridDim.z - 1 - blockIdx.z) : static_cast<int>(blockIdx.z);
  const int r0 = tile * BR;
  int it0, n_it;
  if (MODE == 0) {
    const int q_hi = min(p.T - 1, r0 + BR - 1);
    const int k_last = min(p.T - 1, q_hi + p.src_len);
    it0 = 0;
    n_it = k_last / BS + 1;
  } else {
    const int q_first = max(0, r0 - p.src_len);
    it0 = q_first / BS;
    n_it = (p.T - 1) / BS - it0 + 1;
  }

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&p.ta1);
    ptx::prefetch_tmap(&p.ta2);
    ptx::prefetch_tmap(&p.tb1);
    ptx::prefetch_tmap(&p.tb2);
    ptx::mbar_init(a_full, 1);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&b_full[s], 1);
      ptx::mbar_init(&b_empty[s], 1);
      ptx::mbar_init(&s_full[s], 1);
      ptx::mbar_init(&p_half[s], 32 * C::COMPUTE_WARPS);
    }
    ptx::mbar_init(dp_full, 1);
    ptx::mbar_init(acc_done, 1);
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
      ptx::mbar_expect_tx(a_full, 2 * C::STAT_BYTES);
#pragma unroll
      for (int a = 0; a < C::ATOMS; ++a) {
        ptx::tma_load_3d(sA1 + a * (BR * 128), &p.ta1, a_full, h * HD + a * 64, r0, b);
        ptx::tma_load_3d(sA2 + a * (BR * 128), &p.ta2, a_full, h * HD + a * 64, r0, b);
      }
    }
    __syncwarp();
    for (int it = 0; it < n_it; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      ptx::mbar_wait(&b_empty[s], ph ^ 1);
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(&b_full[s], 2 * C::STR_BYTES);
        const int row = (it0 + it) * BS;
#pragma unroll
        for (int a = 0; a < C::ATOMS; ++a) {
          ptx::tma_load_3d(sB1 + s * C::STR_BYTES + a * (BS * 128), &p.tb1, &b_full[s], h * HD + a * 64, row, b);
          ptx::tma_load_3d(sB2 + s * C::STR_BYTES + a * (BS * 128), &p.tb2, &b_full[s], h * HD + a * 64, row, b);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(BR, BS, 0, 0);
    constexpr uint32_t idesc_acc = ptx::umma_idesc_bf16(BR, HD, 0, 1);  // B operand MN-major
    ptx::mbar_wait(a_full, 0);
    const uint32_t a1 = ptx::smem_u32(sA1), a2 = ptx::smem_u32(sA2);
    // one 128 x 128 score block: D[dst] = A . B^T over the head dim
    auto issue_ss = [&](uint32_t dst, uint32_t a_base, uint32_t b_base) {
#pragma unroll
      for (int k = 0; k < HD / 16; ++k) {
        const uint32_t aoff = (k >> 2) * (BR * 128) + (k & 3) * 32;
        const uint32_t boff = (k >> 2) * (BS * 128) + (k & 3) * 32;
        ptx::umma_f16_ss(dst, ptx::umma_smem_desc(a_base + aoff, 16, 1024), ptx::umma_smem_desc(b_base + boff, 16, 1024),
                         idesc_s, k != 0 ? 1u : 0u);
      }
    };
    auto wait_stage = [&](int it) {
      ptx::mbar_wait(&b_full[it % STAGES], (it / STAGES) & 1);
      ptx::tc_fence_after();
    };
    // accumulate MMAs of one 64-column half of tile `it` (k-steps 4 * hf .. 4 * hf + 3); the caller has seen p_half[hf]
    auto do_acc_half = [&](int it, int hf) {
      const int s = it % STAGES;
      const uint32_t b1 = ptx::smem_u32(sB1 + s * C::STR_BYTES), b2 = ptx::smem_u32(sB2 + s * C::STR_BYTES);
      const uint32_t sbuf = tmem + C::COL_S + (MODE == 0 ? (it & 1) * 128 : 0);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int k = hf * 4 + kk;
          const uint32_t acc = (it | k) != 0 ? 1u : 0u;
          const uint32_t pk = k * 16;   // the 16 streamed columns of k-step k: 8 packed TMEM columns at the front of their own range
          if (MODE == 0) {
            ptx::umma_f16_ts(tmem + C::COL_ACC0, tmem + C::COL_DP + pk, ptx::umma_smem_desc(b1 + k * 2048, BS * 128, 1024),
                             idesc_acc, acc);
          } else {
            ptx::umma_f16_ts(tmem + C::COL_ACC0, sbuf + pk, ptx::umma_smem_desc(b2 + k * 2048, BS * 128, 1024), idesc_acc, acc);
            ptx::umma_f16_ts(tmem + C::COL_ACC1, tmem + C::COL_DP + pk, ptx::umma_smem_desc(b1 + k * 2048, BS * 128, 1024),
                             idesc_acc, acc);
          }
        }
        if (hf == 1) {
          ptx::umma_commit(&b_empty[s]);
          if (it == n_it - 1) ptx::umma_commit(acc_done);
        }
      }
      __syncwarp();
    };
    auto issue_acc_half = [&](int it, int hf) {
      ptx::mbar_wait(&p_half[hf], it & 1);
      do_acc_half(it, hf);
    };
    if (MODE == 0) {
      wait_stage(0);
      if (ptx::elect_one()) {
        issue_ss(tmem + C::COL_S, a1, ptx::smem_u32(sB1));
        ptx::umma_commit(&s_full[0]);
        issue_ss(tmem + C::COL_DP, a2, ptx::smem_u32(sB2));
        ptx::umma_commit(dp_full);
      }
      __syncwarp();
      for (int it = 0; it < n_it; ++it) {
        const bool more = it + 1 < n_it;
        const int sn = (it + 1) % STAGES;
        // S(it+1) as soon as its operands have landed, the accumulate halves as soon as the compute warps hand them
        // over — whichever comes first (the stage of tile it+1 is only released by the accumulate MMAs of tile it-1,
        // so its TMA may still be in flight when the first half of P | dS is ready)
        bool s_pending = more;
        int acc_stage = 0;
        while (s_pending || acc_stage < 2) {
          if (s_pending && ptx::mbar_try_wait(&b_full[sn], ((it + 1) / STAGES) & 1)) {
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
              issue_ss(tmem + C::COL_S + ((it + 1) & 1) * 128, a1, ptx::smem_u32(sB1 + sn * C::STR_BYTES));
              ptx::umma_commit(&s_full[(it + 1) & 1]);
            }
            __syncwarp();
            s_pending = false;
          }
          if (acc_stage < 2 && ptx::mbar_try_wait(&p_half[acc_stage], it & 1)) {
            do_acc_half(it, acc_stage);
            ++acc_stage;
          }
        }
        if (more) {   // dP(it+1) reuses the dS columns: behind the accumulate MMAs of tile it (the pipe executes in order)
          if (ptx::elect_one()) {
            issue_ss(tmem + C::COL_DP, a2, ptx::smem_u32(sB2 + sn * C::STR_BYTES));
            ptx::umma_commit(dp_full);
          }
          __syncwarp();
        }
      }
    } else {
      for (int it = 0; it < n_it; ++it) {
        const int s = it % STAGES;
        wait_stage(it);
        if (ptx::elect_one()) {
          issue_ss(tmem + C::COL_S, a1, ptx::smem_u32(sB1 + s * C::STR_BYTES));
          issue_ss(tmem + C::COL_DP, a2, ptx::smem_u32(sB2 + s * C::STR_BYTES));
          ptx::umma_commit(&s_full[0]);
        }
        __syncwarp();
        issue_acc_half(it, 0);
        issue_acc_half(it, 1);
      }
    }
  } else {
    // ---------------------------------------------------------------- compute warps
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int cq = (warp - 2) >> 2;      // which 16 of a half's 64 streamed columns this thread handles
    const int tid = threadIdx.x - 64;    // 0..511
    const int rpos = r0 + row;           // query (MODE 0) or key (MODE 1) position of this thread
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const long long stat_base = (static_cast<long long>(b) * p.n_heads + h) * p.T;
    float my_lse2 = 0.f, my_dl = 0.f;
    if (MODE == 0 && rpos < p.T) {
      my_lse2 = p.lse[stat_base + rpos] * kLog2e;
      my_dl = p.delta[stat_base + rpos];
    }
    // MODE 1: per-column statistics of the streamed queries, staged one tile ahead (global latency off the critical path)
    float nx_l = 0.f, nx_d = 0.f;
    if (MODE == 1) {
      if (tid < BS) {
        const int q = it0 * BS + tid;
        lse_s[tid] = q < p.T ? p.lse[stat_base + q] * kLog2e : 0.f;
        dl_s[tid] = q < p.T ? p.delta[stat_base + q] : 0.f;
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");
    }
    for (int it = 0; it < n_it; ++it) {
      const int c0 = (it0 + it) * BS;  // first streamed position (keys in MODE 0, queries in MODE 1)
      if (MODE == 1 && tid < BS && it + 1 < n_it) {
        const int q = c0 + BS + tid;
        nx_l = q < p.T ? p.lse[stat_base + q] * kLog2e : 0.f;
        nx_d = q < p.T ? p.delta[stat_base + q] : 0.f;
      }
      bool need_mask;
      if (MODE == 0) need_mask = (c0 + BS - 1 > r0 + p.src_len) || (c0 + BS > p.T) || (r0 + BR > p.T);
      else need_mask = (r0 + BR - 1 > c0 + p.src_len) || (c0 + BS > p.T) || (r0 + BR > p.T);
      const float* lrow = lse_s + (it & 1) * BS;
      const float* drow = dl_s + (it & 1) * BS;
      const uint32_t sbuf = tmem + lane_base + C::COL_S + (MODE == 0 ? (it & 1) * 128 : 0);
      const uint32_t dbuf = tmem + lane_base + C::COL_DP;
      if (MODE == 0) {
        ptx::mbar_wait(&s_full[it & 1], (it >> 1) & 1);
        ptx::mbar_wait(dp_full, it & 1);
      } else {
        ptx::mbar_wait(&s_full[0], it & 1);
      }
      ptx::tc_fence_after();
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int cb = hf * 64 + cq * 16;
        uint32_t rs[16], rd[16];
        ptx::tmem_ld_32x16p(sbuf + cb, rs);
        ptx::tmem_ld_32x16p(dbuf + cb, rd);
        float l2v[16], dlv[16];
        if (MODE == 1) {   // per-column statistics of the streamed queries: four 16-byte shared loads each
#pragma unroll
          for (int e = 0; e < 16; e += 4) {
            const float4 a4 = *reinterpret_cast<const float4*>(lrow + cb + e);
            const float4 d4 = *reinterpret_cast<const float4*>(drow + cb + e);
            l2v[e] = a4.x; l2v[e + 1] = a4.y; l2v[e + 2] = a4.z; l2v[e + 3] = a4.w;
            dlv[e] = d4.x; dlv[e + 1] = d4.y; dlv[e + 2] = d4.z; dlv[e + 3] = d4.w;
          }
        }
        ptx::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          float pv[2], dv[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int cpos = c0 + cb + e + u;
            float l2, dl;
            if (MODE == 0) { l2 = my_lse2; dl = my_dl; }
            else { l2 = l2v[e + u]; dl = dlv[e + u]; }
            float pe = ptx::ex2(fmaf(__uint_as_float(rs[e + u]), p.scale_log2, -l2));
            if (need_mask) {
              const int qq = MODE == 0 ? rpos : cpos;
              const int kk = MODE == 0 ? cpos : rpos;
              if (kk > qq + p.src_len || kk >= p.T || qq >= p.T) pe = 0.f;
            }
            float dpv = __uint_as_float(rd[e + u]);
            if (DROP) {
              const int qq = MODE == 0 ? rpos : cpos;
              const int kk = MODE == 0 ? cpos : rpos;
              const float mult = ptx::drop_mult(p.drop_seed, p.drop_site,
                                                ((static_cast<unsigned long long>(b) * p.n_heads + h) * p.T + qq) *
                                                        static_cast<unsigned long long>((p.T + 1) & ~1) + kk,
                                                p.drop_thresh, p.drop_scale);
              dpv *= mult;
              pv[u] = pe * mult;
            } else {
              pv[u] = pe;
            }
            dv[u] = pe * (dpv - dl);   // the softmax scale is applied once to dQ / dK in the epilogue
          }
          rs[e >> 1] = ptx::pack_bf16(pv[0], pv[1]);
          rd[e >> 1] = ptx::pack_bf16(dv[0], dv[1]);
        }
        // packed results over the front of the columns this thread has just read
        if (MODE == 1) ptx::tmem_st_32x8p(sbuf + cb, rs);
        ptx::tmem_st_32x8p(dbuf + cb, rd);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&p_half[hf]);
      }
      if (MODE == 1 && it + 1 < n_it) {
        if (tid < BS) {
          lse_s[((it + 1) & 1) * BS + tid] = nx_l;
          dl_s[((it + 1) & 1) * BS + tid] = nx_d;
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");
      }
    }
    // epilogue: accumulators -> (un-RoPE) -> bf16 rows
    ptx::mbar_wait(acc_done, 0);
    ptx::tc_fence_after();
    const long long orow = static_cast<long long>(b) * p.T + rpos;
    const int tpos = min(rpos, p.T - 1);
#pragma unroll
    for (int which = 0; which < (MODE == 0 ? 1 : 2); ++which) {
      const bool rot = p.rope != nullptr && (MODE == 0 || which == 1);
      __nv_bfloat16* out = (which == 0 ? p.out0 : p.out1) + orow * (which == 0 ? p.ld0 : p.ld1) + h * HD;
      const uint32_t col = which == 0 ? C::COL_ACC0 : C::COL_ACC1;
      for (int c = cq; c < HD / 32; c += 4) {   // the four warps of a lane quarter take one 32-column slab each
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem + lane_base + col + c * 32, r);
        ptx::tmem_ld_wait();
        float v[32];
        const float osc = (MODE == 0 || which == 1) ? p.scale : 1.0f;   // dS carried no scale: dQ, dK pick it up here
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r[e]) * osc;
        if (rot) {
          const float2* tab = reinterpret_cast<const float2*>(p.rope) + static_cast<long long>((c * 32) >> 1) * p.rope_ld + tpos;
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float2 cs = __ldg(tab + static_cast<long long>(e >> 1) * p.rope_ld);
            const float x0 = v[e], x1 = v[e + 1];
            v[e] = x0 * cs.x + x1 * cs.y;
            v[e + 1] = x1 * cs.x - x0 * cs.y;
          }
        }
        if (rpos < p.T) {
#pragma unroll
          for (int e = 0; e < 32; e += 8) {
            uint4 o;
            o.x = ptx::pack_bf16(v[e], v[e + 1]);
            o.y = ptx::pack_bf16(v[e + 2], v[e + 3]);
            o.z = ptx::pack_bf16(v[e + 4], v[e + 5]);
            o.w = ptx::pack_bf16(v[e + 6], v[e + 7]);
            *reinterpret_cast<uint4*>(out + c * 32 + e) = o;
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

// delta[b,h,t] = sum_d dO * O   (one warp per (row, head))
__global__ void __launch_bounds__(256) attn_bwd_delta_kernel(const __nv_bfloat16* __restrict__ o, long long ldo,
                                                             const __nv_bfloat16* __restrict__ d_o, long long lddo,
                                                             float* __restrict__ delta, int B, int T, int nh, int hd) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= B * T * nh) return;
  const int h = w % nh;
  const long long row = w / nh;
  const __nv_bfloat162* po = reinterpret_cast<const __nv_bfloat162*>(o + row * ldo + h * hd);
  const __nv_bfloat162* pg = reinterpret_cast<const __nv_bfloat162*>(d_o + row * lddo + h * hd);
  float s = 0.f;
  for (int d = lane; d < hd / 2; d += 32) {
    const float2 a = __bfloat1622float2(po[d]), g = __bfloat1622float2(pg[d]);
    s = fmaf(a.x, g.x, fmaf(a.y, g.y, s));
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) {
    const int bb = static_cast<int>(row / T), t = static_cast<int>(row % T);
    delta[(static_cast<long long>(bb) * nh + h) * T + t] = s;
  }
}

template <int HD, int BS, int MODE, bool DROP>
int launch_mode(const sea_attn_bwd_args* a, cudaStream_t s) {
  using C = BCfg<HD, BS, MODE>;
  static bool attr_set[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 16 && !attr_set[dev]) {
    SEA_CUDA_OK(cudaFuncSetAttribute(attn_bwd_tc_kernel<HD, BS, MODE, DROP>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr_set[dev] = true;
  }
  AttnBwdTcParams p;
  const uint64_t w = static_cast<uint64_t>(a->n_heads) * HD;
  const uint64_t T = a->T, B = a->B;
  auto mk = [&](CUtensorMap* m, const void* ptr, long long ld, int rows) {
    return make_tmap_bf16_3d(m, ptr, w, T, B, ld, static_cast<uint64_t>(ld) * T, 64, rows);
  };
  int rc;
  if (MODE == 0) {
    if ((rc = mk(&p.ta1, a->q, a->ldq, BR))) return rc;
    if ((rc = mk(&p.ta2, a->d_o, a->lddo, BR))) return rc;
    if ((rc = mk(&p.tb1, a->k, a->ldk, BS))) return rc;
    if ((rc = mk(&p.tb2, a->v, a->ldv, BS))) return rc;
    p.out0 = static_cast<__nv_bfloat16*>(a->dq); p.ld0 = a->lddq;
    p.out1 = nullptr; p.ld1 = 0;
  } else {
    if ((rc = mk(&p.ta1, a->k, a->ldk, BR))) return rc;
    if ((rc = mk(&p.ta2, a->v, a->ldv, BR))) return rc;
    if ((rc = mk(&p.tb1, a->q, a->ldq, BS))) return rc;
    if ((rc = mk(&p.tb2, a->d_o, a->lddo, BS))) return rc;
    p.out0 = static_cast<__nv_bfloat16*>(a->dv); p.ld0 = a->lddv;
    p.out1 = static_cast<__nv_bfloat16*>(a->dk); p.ld1 = a->lddk;
  }
  p.lse = a->lse; p.delta = a->delta; p.rope = a->rope_table; p.rope_ld = a->rope_ld;
  p.B = a->B; p.T = a->T; p.n_heads = a->n_heads; p.src_len = a->src_len;
  p.scale = a->scale; p.scale_log2 = a->scale * kLog2e;
  p.drop_seed = a->dropout_seed; p.drop_site = a->dropout_site;
  p.drop_thresh = a->dropout_p > 0.f ? static_cast<uint32_t>(static_cast<double>(a->dropout_p) * 4294967296.0) : 0u;
  p.drop_scale = 1.0f / (1.0f - a->dropout_p);
  p.probe = g_attn_bwd_probe;
  const int tiles = (a->T + BR - 1) / BR;
  dim3 grid(a->n_heads * a->B, 1, tiles * (MODE == 1 ? C::HALVES : 1));
  SEA_LAUNCH((attn_bwd_tc_kernel<HD, BS, MODE, DROP>), grid, kThreads, C::SMEM, s, p);
  return static_cast<int>(cudaGetLastError());
}

template <int HD, int MODE, bool DROP>
int launch_mode2(const sea_attn_bwd_args* a, cudaStream_t s) {
  using C = B2Cfg<HD, MODE>;
  static bool attr_set[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 16 && !attr_set[dev]) {
    SEA_CUDA_OK(cudaFuncSetAttribute(attn_bwd_tc2_kernel<HD, MODE, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr_set[dev] = true;
  }
  AttnBwdTcParams p;
  const uint64_t w = static_cast<uint64_t>(a->n_heads) * HD;
  const uint64_t T = a->T, B = a->B;
  auto mk = [&](CUtensorMap* m, const void* ptr, long long ld, int rows) {
    return make_tmap_bf16_3d(m, ptr, w, T, B, ld, static_cast<uint64_t>(ld) * T, 64, rows);
  };
  int rc;
  if (MODE == 0) {
    if ((rc = mk(&p.ta1, a->q, a->ldq, BR))) return rc;
    if ((rc = mk(&p.ta2, a->d_o, a->lddo, BR))) return rc;
    if ((rc = mk(&p.tb1, a->k, a->ldk, C::BS))) return rc;
    if ((rc = mk(&p.tb2, a->v, a->ldv, C::BS))) return rc;
    p.out0 = static_cast<__nv_bfloat16*>(a->dq); p.ld0 = a->lddq;
    p.out1 = nullptr; p.ld1 = 0;
  } else {
    if ((rc = mk(&p.ta1, a->k, a->ldk, BR))) return rc;
    if ((rc = mk(&p.ta2, a->v, a->ldv, BR))) return rc;
    if ((rc = mk(&p.tb1, a->q, a->ldq, C::BS))) return rc;
    if ((rc = mk(&p.tb2, a->d_o, a->lddo, C::BS))) return rc;
    p.out0 = static_cast<__nv_bfloat16*>(a->dv); p.ld0 = a->lddv;
    p.out1 = static_cast<__nv_bfloat16*>(a->dk); p.ld1 = a->lddk;
  }
  p.lse = a->lse; p.delta = a->delta; p.rope = a->rope_table; p.rope_ld = a->rope_ld;
  p.B = a->B; p.T = a->T; p.n_heads = a->n_heads; p.src_len = a->src_len;
  p.scale = a->scale; p.scale_log2 = a->scale * kLog2e;
  p.drop_seed = a->dropout_seed; p.drop_site = a->dropout_site;
  p.drop_thresh = a->dropout_p > 0.f ? static_cast<uint32_t>(static_cast<double>(a->dropout_p) * 4294967296.0) : 0u;
  p.drop_scale = 1.0f / (1.0f - a->dropout_p);
  p.probe = g_attn_bwd_probe;
  const int tiles = (a->T + BR - 1) / BR;
  dim3 grid(a->n_heads * a->B, 1, tiles);
  SEA_LAUNCH((attn_bwd_tc2_kernel<HD, MODE, DROP>), grid, C::THREADS, C::SMEM, s, p);
  return static_cast<int>(cudaGetLastError());
}

template <int HD>
int launch_both2(const sea_attn_bwd_args* a, cudaStream_t s) {
  if (a->dropout_p > 0.f) {
    int rc = launch_mode2<HD, 0, true>(a, s);
    if (rc) return rc;
    return launch_mode2<HD, 1, true>(a, s);
  }
  int rc = launch_mode2<HD, 0, false>(a, s);
  if (rc) return rc;
  return launch_mode2<HD, 1, false>(a, s);
}

template <int HD, int BS>
int launch_both(const sea_attn_bwd_args* a, cudaStream_t s) {
  if (a->dropout_p > 0.f) {
    int rc = launch_mode<HD, BS, 0, true>(a, s);
    if (rc) return rc;
    return launch_mode<HD, BS, 1, true>(a, s);
  }
  int rc = launch_mode<HD, BS, 0, false>(a, s);
  if (rc) return rc;
  return launch_mode<HD, BS, 1, false>(a, s);
}

}  // namespace

bool attention_bwd_tc_supported(const sea_attn_bwd_args* a) {
  if (a->prec != SEA_PREC_BF16) return false;
  if (a->head_dim != 64 && a->head_dim != 128 && a->head_dim != 256) return false;
  if (a->src_len < 0) return false;
  if ((a->ldq % 8) || (a->ldk % 8) || (a->ldv % 8) || (a->lddo % 8) || (a->ldo % 2)) return false;
  if ((a->lddq % 8) || (a->lddk % 8) || (a->lddv % 8)) return false;
  const uintptr_t all = reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) |
                        reinterpret_cast<uintptr_t>(a->v) | reinterpret_cast<uintptr_t>(a->d_o) |
                        reinterpret_cast<uintptr_t>(a->dq) | reinterpret_cast<uintptr_t>(a->dk) |
                        reinterpret_cast<uintptr_t>(a->dv);
  if (all & 15) return false;
  if (reinterpret_cast<uintptr_t>(a->o) & 3) return false;
  if (a->rope_table != nullptr && a->rope_ld < a->T) return false;
  if (a->dropout_p < 0.f || a->dropout_p >= 1.f) return false;
  return true;
}

int attention_bwd_tc(const sea_attn_bwd_args* a, cudaStream_t s) {
  int rc = ensure_init();
  if (rc) return rc;
  const int total_warps = a->B * a->T * a->n_heads;
  SEA_LAUNCH(attn_bwd_delta_kernel, (total_warps + 7) / 8, 256, 0, s,
             static_cast<const __nv_bfloat16*>(a->o), a->ldo, static_cast<const __nv_bfloat16*>(a->d_o), a->lddo,
             a->delta, a->B, a->T, a->n_heads, a->head_dim);
  switch (a->head_dim) {
    case 64: return g_attn_bwd_wide ? launch_both2<64>(a, s) : launch_both<64, 64>(a, s);
    case 128: return g_attn_bwd_wide ? launch_both2<128>(a, s) : launch_both<128, 64>(a, s);
    case 256: return launch_both<256, 64>(a, s);
    default: return SEA_ERR_UNSUPPORTED;
  }
}

}  // namespace sea

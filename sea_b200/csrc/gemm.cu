// K1 — persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M,N] = A[M,K] · B[N,K]^T   (both operands K-major bf16, fp32 accumulation in TMEM)
//
// Roles (320 threads, 1 CTA / SM, grid = min(#tiles, #SMs), static round-robin tile schedule):
//   warp 0      TMA producer: cp.async.bulk.tensor 128B-swizzled A (128x64) and B (BNx64) tiles
//               into a STAGES-deep shared-memory ring, completion on `full` mbarriers;
//   warp 1      allocates TMEM (2 accumulator stages x BN columns) and issues tcgen05.mma
//               (M=128, N=BN, K=16) from lane 0; tcgen05.commit releases ring slots (`empty`)
//               and publishes finished accumulators (`tfull`);
//   warps 2..9  epilogue (two warps per TMEM lane quarter, alternating 32-column chunks):
//               tcgen05.ld 32x32b (one accumulator row per thread), fused
//               bias / residual / RoPE / GELU' / GELU / dtype casts, vectorised global stores,
//               then hand the TMEM stage back (`tempty`) so the next tile's MMAs overlap.
//
// Up to 4 same-shape problems share one launch ("groups": the V independent field streams of
// models/temporal.py:135-146 run their Linear layers side by side).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#include "../../include/sea_b200.h"
#include "internal.h"
#include "ptx.cuh"

namespace sea {

namespace {

constexpr int BM = 128;
constexpr int KA = 64;  // one swizzle atom along K: 64 bf16 = 128 B
constexpr int kMaxGroups = 4;
constexpr int kThreads = 320;  // TMA warp + MMA warp + 8 epilogue warps

struct DevEpilogue {
  const float* bias;
  const float* residual;
  const __nv_bfloat16* gelu_grad_of;
  const float* rope_table;
  float* out_f32;
  __nv_bfloat16* out_pre_bf16;
  __nv_bfloat16* out_bf16;
  long long ld_residual, ld_gelu, ld_out_f32, ld_out_pre_bf16, ld_out_bf16;
  int act, rope_cols, head_dim, seq_len, rope_ld, rope_pos0;
  unsigned long long drop_seed; uint32_t drop_thresh, drop_site; float drop_scale;  // drop_thresh 0 = no dropout
  int res_rows; long long res_bs;  // res_rows > 0: residual row m at (m / res_rows) * res_bs + (m % res_rows) * ld_residual
  float rope_sign;
  int vec8;  // every row base / pitch is 32-byte aligned: use 256-bit global accesses
};

struct alignas(64) GemmParams {
  CUtensorMap tma_a[kMaxGroups];
  CUtensorMap tma_b[kMaxGroups];
  DevEpilogue epi[kMaxGroups];
  int M, N, K;
  int groups, tiles_m, tiles_n;
  int chunk_kb;  // k-blocks per accumulation chunk (== num_kb when not chunked)
  int cluster;              // 2: two-CTA clusters with multicast B halves (N-tile 256, data-parallel only)
  int sk_grid;              // CTAs of a launch with a stream-K tail
  int dp_tiles;             // tiles [0, dp_tiles) are processed whole, round-robin (full waves)
  int units_per_cta;        // > 0: the other tiles are stream-K'd: contiguous (tile, k-block) units per CTA
  float* sk_ws;             // stream-K partial tiles: [grid][2][BM x BN] fp32
  unsigned int* sk_counters;  // stream-K arrivals per tile (zero between launches)
  int b_is_static;  // B was written before the previous kernel in the stream started (weights)
  int debug;        // tuning probe: 1 = no TMA traffic (MMA pacing only), 2 = no MMA (TMA pacing only)
  unsigned long long* trace;  // tuning probe (sea_gemm_debug_trace): CTA 0 records %globaltimer at 8 hand-off points
};

// Compiled in only with -DSEA_GEMM_TRACE (scripts/gemm_latency.py documents the build): a `lane == 0` test inside the
// TMA / MMA warps' loops costs the uniform-datapath issue of UTMALDG / UTCHMMA (measured: +6 % on the whole rollout).
__device__ __forceinline__ void trace_mark(const GemmParams& p, int slot) {
#ifdef SEA_GEMM_TRACE
  if (p.trace != nullptr && blockIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[slot] = t;
  }
#else
  (void)p; (void)slot;
#endif
}

// BK = K extent of one pipeline stage.  The issuing thread pays a fixed ~390 cycles per stage
// (mbarrier wait, fence, tcgen05.commit, loop), independent of the tile; a stage must therefore
// carry at least that much tensor-pipe work: BK = 64 is enough for N = 256 / 192 (4 x 128 / 96
// cycles), the narrower tiles take two swizzle atoms per stage (BK = 128).
template <int BN>
struct Cfg {
  static constexpr int BK = (BN <= 128) ? 128 : 64;
  static constexpr int KATOMS = BK / KA;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 192 ? 5 : (BN == 128 ? 3 : 4));
  static constexpr int TMEM_COLS = (BN == 192) ? 512 : 2 * BN;  // allocation must be a power of two
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ +
                                    2 * BN * 4 /*bias slice of the tile in each accumulator stage*/;
};

// Residual values of one epilogue chunk (row m, columns [n0, n0 + 32)) into registers.  Issued ahead of the chunk's
// accumulator (the residual does not depend on this launch's MMAs), so its L2 round trip overlaps the mainloop / the
// previous chunk instead of sitting between tcgen05.ld and the stores.
__device__ __forceinline__ void load_residual_chunk(const DevEpilogue& e, int m, int n0, int M, int N, uint32_t (&res)[32]) {
  if (m >= M || n0 >= N) return;
  const int nvalid = min(32, N - n0);
  const float* rp = e.residual + (e.res_rows > 0 ? static_cast<long long>(m / e.res_rows) * e.res_bs + static_cast<long long>(m % e.res_rows) * e.ld_residual
                                                 : static_cast<long long>(m) * e.ld_residual) + n0;
  if (e.vec8) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      if (j < nvalid) {
        uint32_t t[8];
        ptx::ldg256(rp + j, t);
#pragma unroll
        for (int q = 0; q < 8; ++q) res[j + q] = t[q];
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (j < nvalid) {
        const float4 b = *reinterpret_cast<const float4*>(rp + j);
        res[j] = __float_as_uint(b.x); res[j + 1] = __float_as_uint(b.y);
        res[j + 2] = __float_as_uint(b.z); res[j + 3] = __float_as_uint(b.w);
      }
    }
  }
}

// bias_s: this chunk's 32 bias values in shared memory (staged per tile by the epilogue warps while the mainloop
// runs; zero past N); res: load_residual_chunk's registers (read only when first && e.residual).
// n0_next >= 0: as soon as this chunk's residual registers have been consumed they are refilled for the chunk at
// n0_next, so that load is in flight under this chunk's stores and the next accumulator read (a second register
// buffer would push the kernel from ~140 to 168 registers: measured, that leaves no room for an NCCL CTA beside a GEMM
// CTA on the SM and the overlapped data-parallel step at per-GPU batch 2 went from 2.29 to 2.63 ms on 2 GPUs).
__device__ __forceinline__ void epilogue_chunk(const DevEpilogue& e, const uint32_t (&r)[32], const float* bias_s,
                                               uint32_t (&res)[32], int n0_next,
                                               int m, int n0, int M, int N, bool first, bool last) {
  // One thread = one output row m, 32 consecutive columns starting at n0.
  if (m >= M || n0 >= N) return;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  const int nvalid = min(32, N - n0);  // multiple of 8 (N % 8 == 0 is enforced on the host)

  if (!first) {
    // chunked accumulation (fp32-parity mode): this thread wrote the running sum itself
    const float* ap = e.out_f32 + static_cast<long long>(m) * e.ld_out_f32 + n0;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (j < nvalid) {
        const float4 b = *reinterpret_cast<const float4*>(ap + j);
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    }
  }
  if (!last) {
    if (first) {
      if (e.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if (j < nvalid) {
            const float4 b = *reinterpret_cast<const float4*>(bias_s + j);
            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
          }
        }
      }
      if (e.residual != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) v[j] += __uint_as_float(res[j]);
        if (n0_next >= 0) load_residual_chunk(e, m, n0_next, M, N, res);
      }
    }
    float* op = e.out_f32 + static_cast<long long>(m) * e.ld_out_f32 + n0;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (j < nvalid) *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    return;
  }

  if (first && e.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (j < nvalid) {
        const float4 b = *reinterpret_cast<const float4*>(bias_s + j);
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    }
  }
  if (e.drop_thresh != 0u) {
    // nn.Dropout on the Linear's output, before the skip connection is added (base_blocks.py:47)
    const unsigned long long idx0 = static_cast<unsigned long long>(m) * N + n0;   // even: N % 8 == 0, n0 % 32 == 0
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const uint2 h = ptx::drop_hash(e.drop_seed, e.drop_site, (idx0 + j) >> 1);
      v[j] = h.x >= e.drop_thresh ? v[j] * e.drop_scale : 0.f;
      v[j + 1] = h.y >= e.drop_thresh ? v[j + 1] * e.drop_scale : 0.f;
    }
  }
  if (first && e.residual != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < nvalid) v[j] += __uint_as_float(res[j]);
    if (n0_next >= 0) load_residual_chunk(e, m, n0_next, M, N, res);
  }
  if (e.rope_table != nullptr && n0 < e.rope_cols) {
    // models/base_blocks.py:314-324 — interleaved pairs (x[2k], x[2k+1]) times (cos + i sin).
    // table is pair-major: tab[pair][t]; the 32 lanes of a warp hold consecutive rows, i.e.
    // consecutive t, so every load instruction reads one contiguous 256-byte run
    const int t = e.rope_pos0 + m % e.seq_len;
    const int d0 = n0 % e.head_dim;
    const float2* tab = reinterpret_cast<const float2*>(e.rope_table) +
                        static_cast<long long>(d0 >> 1) * e.rope_ld + t;
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const float2 cs = __ldg(tab + static_cast<long long>(j >> 1) * e.rope_ld);
      const float s = cs.y * e.rope_sign;
      const float x0 = v[j], x1 = v[j + 1];
      v[j] = x0 * cs.x - x1 * s;
      v[j + 1] = x0 * s + x1 * cs.x;
    }
  }
  if (e.gelu_grad_of != nullptr) {
    const __nv_bfloat16* gp = e.gelu_grad_of + static_cast<long long>(m) * e.ld_gelu + n0;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      if (j < nvalid) {
        const uint4 raw = *reinterpret_cast<const uint4*>(gp + j);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = __bfloat1622float2(h[q]);
          v[j + 2 * q] *= ptx::gelu_erf_grad(f.x);
          v[j + 2 * q + 1] *= ptx::gelu_erf_grad(f.y);
        }
      }
    }
  }
  if (e.out_f32 != nullptr) {
    float* op = e.out_f32 + static_cast<long long>(m) * e.ld_out_f32 + n0;
    if (e.vec8) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        if (j < nvalid) {
          uint32_t t[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) t[q] = __float_as_uint(v[j + q]);
          ptx::stg256(op + j, t);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        if (j < nvalid) *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
  }
  if (e.out_pre_bf16 != nullptr) {
    __nv_bfloat16* op = e.out_pre_bf16 + static_cast<long long>(m) * e.ld_out_pre_bf16 + n0;
    if (e.vec8) {
#pragma unroll
      for (int j = 0; j < 32; j += 16) {
        if (j < nvalid) {
          uint32_t t[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) t[q] = ptx::pack_bf16(v[j + 2 * q], v[j + 2 * q + 1]);
          ptx::stg256(op + j, t);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        if (j < nvalid) {
          uint4 o;
          o.x = ptx::pack_bf16(v[j], v[j + 1]);
          o.y = ptx::pack_bf16(v[j + 2], v[j + 3]);
          o.z = ptx::pack_bf16(v[j + 4], v[j + 5]);
          o.w = ptx::pack_bf16(v[j + 6], v[j + 7]);
          *reinterpret_cast<uint4*>(op + j) = o;
        }
      }
    }
  }
  if (e.out_bf16 != nullptr) {
    if (e.act == SEA_ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = ptx::gelu_fast(v[j]);
    }
    __nv_bfloat16* op = e.out_bf16 + static_cast<long long>(m) * e.ld_out_bf16 + n0;
    if (e.vec8) {
#pragma unroll
      for (int j = 0; j < 32; j += 16) {
        if (j < nvalid) {
          uint32_t t[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) t[q] = ptx::pack_bf16(v[j + 2 * q], v[j + 2 * q + 1]);
          ptx::stg256(op + j, t);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        if (j < nvalid) {
          uint4 o;
          o.x = ptx::pack_bf16(v[j], v[j + 1]);
          o.y = ptx::pack_bf16(v[j + 2], v[j + 3]);
          o.z = ptx::pack_bf16(v[j + 4], v[j + 5]);
          o.w = ptx::pack_bf16(v[j + 6], v[j + 7]);
          *reinterpret_cast<uint4*>(op + j) = o;
        }
      }
    }
  }
}

// AMN / BMN = false: the operand is stored [M,K] / [N,K], K contiguous (the nn.Linear forward case).
// AMN / BMN = true : the operand is stored [K,M] / [K,N] (M / N contiguous) and is consumed through
//   MN-major UMMA descriptors, so no transpose is ever materialised:
//     forward   y  = x W^T        A = x  [M,K]            B = W  [N,K]            (false, false)
//     dgrad     dx = dy W         A = dy [M,N] K-major    B = W  [N,K] = [Kc, N'] (false, true)
//     wgrad     dW = dy^T x       A = dy [M,N] = [Kc, M'] B = x  [M,K] = [Kc, N'] (true,  true)
//   Stage layout of an MN-major operand, per 64-wide M/N chunk: [BK k-rows][128 B]; descriptors step
//   16 k-rows (2048 B) per MMA, LBO = chunk stride, SBO = 1024 B (8 k-rows).
// CL = 2: thread-block cluster of two CTAs that work on vertically adjacent tiles (same N-tile, M-tiles
// 2k and 2k+1).  They need the same B tile, so each CTA requests only HALF of it and TMA multicasts that
// half into both CTAs' shared memory: B's L2 -> SM traffic halves (the 128x256 tiles are L2-bandwidth
// bound at full grid: 48 KB per 4.2 MFLOP).  A ring stage is refilled only when both CTAs have consumed
// it (the MMA warps' commits arrive on both CTAs' `empty` barriers).
// 144 registers, not the 168 that 320 threads (allocated as 12 warps) would allow: the data-parallel train step overlaps
// NCCL's all-reduce with the backward's GEMMs, and an NCCL CTA only finds room beside a GEMM CTA when the latter leaves
// registers free — at 168 the overlapped step at per-GPU batch 2 went from 2.29 to 2.63 ms on 2 GPUs (3.41 against
// 2.42 on 8); the rollout pays ~1 % for the cap.
template <int BN, bool AMN, bool BMN, int CL>
__global__ void __maxnreg__(144)
gemm_bf16_tn_kernel(const __grid_constant__ GemmParams p) {
  using C = Cfg<BN>;
  constexpr int STAGES = C::STAGES;
  constexpr int BK = C::BK;
  constexpr int A_BYTES = C::A_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * C::STAGE_BYTES);
  uint64_t* full = bars;                 // [STAGES]  TMA -> MMA
  uint64_t* empty = bars + STAGES;       // [STAGES]  MMA -> TMA
  uint64_t* tfull = bars + 2 * STAGES;   // [2]       MMA -> epilogue
  uint64_t* tempty = tfull + 2;          // [2]       epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  int* sk_flag = reinterpret_cast<int*>(tmem_slot + 1);
  float* bias_stage = reinterpret_cast<float*>(smem + STAGES * C::STAGE_BYTES + 256);   // [2][BN]

  // shfl-broadcast makes the warp index provably warp-uniform for the compiler: the role branches
  // below are then uniform, and the single-thread regions are entered through elect.sync, so the
  // uniform-datapath instructions (UTMALDG, UTCHMMA, UTCBAR) are issued directly.  With a plain
  // `lane == 0` test ptxas wraps EVERY such instruction in a divergence ("waterfall") loop that
  // costs ~130 cycles per tcgen05.mma and paces the whole mainloop.
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) trace_mark(p, 0);
  const int num_kb = (p.K + BK - 1) / BK;
  const int crank = CL > 1 ? static_cast<int>(ptx::cluster_ctarank()) : 0;
  const int vcta = blockIdx.x / CL, vgrid = gridDim.x / CL;      // tiles are dealt to clusters
  const int tiles_m_v = (p.tiles_m + CL - 1) / CL;                // M-tile pairs
  const int tiles_per_group = tiles_m_v * p.tiles_n;
  const int total_tiles = tiles_per_group * p.groups;

  if (threadIdx.x == 0) {
    for (int g = 0; g < p.groups; ++g) {
      ptx::prefetch_tmap(&p.tma_a[g]);
      ptx::prefetch_tmap(&p.tma_b[g]);
    }
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], CL);   // one commit per CTA of the cluster
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull[a], 1);
      ptx::mbar_init(&tempty[a], 8);  // one arrive per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, C::TMEM_COLS);
  ptx::tc_fence_before();
  if (CL > 1) ptx::cluster_sync_all();   // the peer's barriers must exist before anything lands on them
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) trace_mark(p, 1);
  // prologue done (barriers, TMEM, descriptor prefetch): let the next kernel start its own
  ptx::pdl_trigger();

  // B tile of one stage.  CL = 1: the whole tile.  CL = 2: this CTA's half (K-major: rows [crank*BN/2, +BN/2)
  // through a half-height box; MN-major: its half of the 64-column chunks), multicast to both CTAs.
  auto load_b_tile = [&](int stg, int g, int tn, int kb) {
    uint8_t* dst = smem_b + stg * C::B_BYTES;
    if (CL == 1) {
      if (BMN) {
#pragma unroll
        for (int a = 0; a < BN / 64; ++a)
          ptx::tma_load_2d(dst + a * (BK * 128), &p.tma_b[g], &full[stg], tn * BN + a * 64, kb * BK);
      } else {
#pragma unroll
        for (int a = 0; a < C::KATOMS; ++a)
          ptx::tma_load_2d(dst + a * (BN * 128), &p.tma_b[g], &full[stg], kb * BK + a * KA, tn * BN);
      }
    } else {
      constexpr uint16_t kMask = (1u << CL) - 1u;
      if (BMN) {
#pragma unroll
        for (int a0 = 0; a0 < BN / 64 / CL; ++a0) {
          const int a = crank * (BN / 64 / CL) + a0;
          ptx::tma_load_2d_mc(dst + a * (BK * 128), &p.tma_b[g], &full[stg], tn * BN + a * 64, kb * BK, kMask);
        }
      } else {
#pragma unroll
        for (int a = 0; a < C::KATOMS; ++a)
          ptx::tma_load_2d_mc(dst + a * (BN * 128) + crank * (BN / CL) * 128, &p.tma_b[g], &full[stg],
                              kb * BK + a * KA, tn * BN + crank * (BN / CL), kMask);
      }
    }
  };

  // Work decomposition.  A "segment" is a run of k-blocks [kb0, kb1) of one output tile, accumulated
  // into one TMEM stage.
  //   phase A (data-parallel): tiles [0, dp_tiles) round-robin over the CTAs, whole tiles (cut into
  //     chunk_kb pieces in the fp32-parity mode).  All CTAs walk K in lockstep, so the CTAs that share
  //     an A or B panel hit it in L2 at the same time.
  //   phase B (stream-K tail): the (tile, k-block) space of the remaining tiles — the ragged last wave,
  //     or everything when there are fewer tiles than SMs — is cut into equal contiguous ranges, one per
  //     CTA, so no SM idles.  A tile covered by several CTAs is finished by whichever of them arrives
  //     last (partials + arrival counter in the L2 workspace).
  struct Seg {
    int tile, kb0, kb1;
    bool sk;
  };
  struct SegIt { int cursor, sub, u, u1; };
  auto seg_begin = [&](SegIt& it) {
    it.cursor = vcta; it.sub = 0;
    it.u = vcta * p.units_per_cta;
    it.u1 = p.units_per_cta > 0 ? min(it.u + p.units_per_cta, (total_tiles - p.dp_tiles) * num_kb) : 0;
  };
  auto seg_next = [&](SegIt& it, Seg& sg) -> bool {
    if (it.cursor < p.dp_tiles && it.cursor < total_tiles) {
      sg.tile = it.cursor;
      sg.kb0 = it.sub;
      sg.kb1 = min(num_kb, it.sub + p.chunk_kb);
      sg.sk = false;
      it.sub = sg.kb1;
      if (it.sub >= num_kb) { it.sub = 0; it.cursor += vgrid; }
      return true;
    }
    if (p.units_per_cta > 0 && it.u < it.u1) {
      const int t = it.u / num_kb;
      sg.tile = p.dp_tiles + t;
      sg.kb0 = it.u - t * num_kb;
      sg.kb1 = min(num_kb, sg.kb0 + (it.u1 - it.u));
      sg.sk = true;
      it.u += sg.kb1 - sg.kb0;
      return true;
    }
    return false;
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    SegIt it;
    Seg sg;
    seg_begin(it);
    // The B operand is a weight matrix: it does not depend on the kernel running before this one,
    // so the first ring-full of weight tiles is requested BEFORE griddepcontrol.wait and streams
    // in from HBM while the upstream kernel drains.  Only the A tiles (activations) wait.
    int pre = 0;
    {
      SegIt it2 = it;
      Seg first_seg;
      if (p.b_is_static && seg_next(it2, first_seg)) {
        pre = min(first_seg.kb1 - first_seg.kb0, STAGES);
        const int g = first_seg.tile / tiles_per_group;
        const int tn = (first_seg.tile - g * tiles_per_group) / tiles_m_v;
        if (ptx::elect_one()) {
          for (int i = 0; i < pre; ++i) {
            const int kb = first_seg.kb0 + i;
            ptx::mbar_expect_tx(&full[i], C::STAGE_BYTES);
            load_b_tile(i, g, tn, kb);
          }
        }
        __syncwarp();
      }
    }
    ptx::pdl_wait();
    if (lane == 0) trace_mark(p, 2);
    int issued = 0;  // k-blocks issued by this CTA so far (the first `pre` already have their B tile)
    while (seg_next(it, sg)) {
      const int g = sg.tile / tiles_per_group;
      const int r = sg.tile - g * tiles_per_group;
      const int tm = (r % tiles_m_v) * CL + crank;
      const int tn = r / tiles_m_v;
      for (int kb = sg.kb0; kb < sg.kb1; ++kb, ++issued) {
        const bool prefetched = issued < pre;
        if (!prefetched) ptx::mbar_wait(&empty[stage], phase ^ 1);
        if (ptx::elect_one()) {
          auto load_a = [&]() {
            if (AMN) {
#pragma unroll
              for (int a = 0; a < BM / 64; ++a)
                ptx::tma_load_2d(smem_a + stage * A_BYTES + a * (BK * 128), &p.tma_a[g], &full[stage], tm * BM + a * 64, kb * BK);
            } else {
#pragma unroll
              for (int a = 0; a < C::KATOMS; ++a)
                ptx::tma_load_2d(smem_a + stage * A_BYTES + a * (BM * 128), &p.tma_a[g], &full[stage], kb * BK + a * KA, tm * BM);
            }
          };
          auto load_b = [&]() { load_b_tile(stage, g, tn, kb); };
          if (prefetched) {
            load_a();
          } else if (p.debug == 1) {
            ptx::mbar_arrive(&full[stage]);
          } else {
            ptx::mbar_expect_tx(&full[stage], C::STAGE_BYTES);
            load_a();
            load_b();
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(BM, BN, AMN ? 1 : 0, BMN ? 1 : 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    SegIt it;
    Seg sg;
    seg_begin(it);
    while (seg_next(it, sg)) {
      const int kb0 = sg.kb0, kb1 = sg.kb1;
      ptx::mbar_wait(&tempty[acc], acc_phase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(&full[stage], phase);
        ptx::tc_fence_after();
        if (kb == kb0 && lane == 0) trace_mark(p, 3);
        if (ptx::elect_one()) {
          const uint32_t a_base = ptx::smem_u32(smem_a + stage * A_BYTES);
          const uint32_t b_base = ptx::smem_u32(smem_b + stage * C::B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = AMN ? ptx::umma_smem_desc(a_base + k * 2048, BK * 128, 1024)
                                      : ptx::umma_smem_desc(a_base + (k >> 2) * (BM * 128) + (k & 3) * 32, 16, 1024);
            const uint64_t bdesc = BMN ? ptx::umma_smem_desc(b_base + k * 2048, BK * 128, 1024)
                                      : ptx::umma_smem_desc(b_base + (k >> 2) * (BN * 128) + (k & 3) * 32, 16, 1024);
            if (p.debug != 2) ptx::umma_f16_ss(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k != 0) ? 1u : 0u);
          }
          if (CL > 1) ptx::umma_commit_mc(&empty[stage], static_cast<uint16_t>((1u << CL) - 1u));
          else ptx::umma_commit(&empty[stage]);
          if (kb == kb1 - 1) ptx::umma_commit(&tfull[acc]);
        }
        __syncwarp();
        if (kb == kb1 - 1 && lane == 0) trace_mark(p, 4);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // --------------------------------------------------------------- epilogue
    ptx::pdl_wait();               // residual / gelu_grad_of / out_f32 (chunked) come from upstream kernels
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32) belong to this warp
    const int half = (warp - 2) >> 2;  // two warps share a lane quarter and split the column chunks
    const int row = quarter * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    SegIt it;
    Seg sg;
    seg_begin(it);
    int nseg = 0;   // stream-K segments seen so far
    while (seg_next(it, sg)) {
      const int g = sg.tile / tiles_per_group;
      const int r = sg.tile - g * tiles_per_group;
      const int tm = (r % tiles_m_v) * CL + crank;
      const int tn = r / tiles_m_v;
      const DevEpilogue& e = p.epi[g];
      const int m = tm * BM + row;
      const bool first = sg.kb0 == 0, last = sg.kb1 >= num_kb;
      const bool partial = sg.sk && !(first && last);
      // While the mainloop of this segment runs: bias slice of the tile -> shared memory, residual of the first chunk
      // -> registers.  Neither depends on the accumulator; fetched after the tfull wait they put one L2 round trip per
      // 32-column chunk on the launch's critical path (~0.4 us per chunk, 1-3 us of a single-wave GEMM).
      constexpr int NCH = BN / 64;          // 32-column chunks per thread
      float* bias_s = bias_stage + acc * BN;
      const bool use_res = first && e.residual != nullptr;
      uint32_t regs[32], resv[32];
      if (e.bias != nullptr) {
        const int te = static_cast<int>(threadIdx.x) - 64;
        if (te < BN) {
          const int col = tn * BN + te;
          bias_s[te] = col < p.N ? __ldg(e.bias + col) : 0.f;
        }
      }
      if (use_res && !partial) load_residual_chunk(e, m, tn * BN + half * 32, p.M, p.N, resv);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      ptx::mbar_wait(&tfull[acc], acc_phase);
      ptx::tc_fence_after();
      if (threadIdx.x == 128) trace_mark(p, 5);   // warp 4: TMEM lanes 0-31 = rows that exist at any M
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                             static_cast<uint32_t>(acc * BN);
      if (!partial) {
#pragma unroll 1
        for (int i = 0; i < NCH; ++i) {
          const int c = half + 2 * i;
          ptx::tmem_ld_32x32(t_row + c * 32, regs);
          ptx::tmem_ld_wait_on(regs);
          epilogue_chunk(e, regs, bias_s + c * 32, resv, (use_res && i + 1 < NCH) ? tn * BN + (c + 2) * 32 : -1,
                         m, tn * BN + c * 32, p.M, p.N, first, last);
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
      } else {
        // stream-K: this CTA holds only part of the tile's K range.  Park the fp32 partial in the
        // workspace (slot 0 = the CTA's first segment, slot 1 = any later one), then count arrivals.
        float* slot = p.sk_ws + (static_cast<size_t>(blockIdx.x) * 2 + (nseg == 0 ? 0 : 1)) * (BM * BN);
#pragma unroll 1
        for (int c = half; c < BN / 32; c += 2) {
          uint32_t part[32];
          ptx::tmem_ld_32x32(t_row + c * 32, part);
          ptx::tmem_ld_wait();
          // slot layout [chunk][j][row] in float4 units: lanes write consecutive 16-byte words (coalesced);
          // the reducer below uses the same (row, chunk, j) -> address map, nobody else reads a slot
          float4* dst = reinterpret_cast<float4*>(slot) + static_cast<size_t>(c) * 8 * BM + row;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[j * BM] = make_float4(__uint_as_float(part[4 * j]), __uint_as_float(part[4 * j + 1]),
                                      __uint_as_float(part[4 * j + 2]), __uint_as_float(part[4 * j + 3]));
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
        __threadfence();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int tl = sg.tile - p.dp_tiles;   // tile index inside the stream-K region
        const int c_first = (tl * num_kb) / p.units_per_cta;
        const int c_last = ((tl + 1) * num_kb - 1) / p.units_per_cta;
        if (threadIdx.x == 64) *sk_flag = static_cast<int>(atomicAdd(p.sk_counters + sg.tile, 1u));
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (*sk_flag == c_last - c_first) {   // last contributor: sum the partials in CTA order, finish the tile
          __threadfence();
#pragma unroll 1
          for (int c = half; c < BN / 32; c += 2) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
            for (int cc = c_first; cc <= c_last; ++cc) {
              // contributor cc parked this tile in slot 0 iff the tile is where its range starts
              const int which = ((cc * p.units_per_cta) / num_kb == tl) ? 0 : 1;
              const float4* src = reinterpret_cast<const float4*>(p.sk_ws + (static_cast<size_t>(cc) * 2 + which) * (BM * BN)) +
                                  static_cast<size_t>(c) * 8 * BM + row;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 t = __ldcg(src + j * BM);
                v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
              }
            }
            uint32_t sum[32], rres[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) sum[j] = __float_as_uint(v[j]);
            if (e.residual != nullptr) load_residual_chunk(e, m, tn * BN + c * 32, p.M, p.N, rres);
            epilogue_chunk(e, sum, bias_s + c * 32, rres, -1, m, tn * BN + c * 32, p.M, p.N, true, true);
          }
          if (threadIdx.x == 64) p.sk_counters[sg.tile] = 0u;   // self-cleaning for the next launch
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");   // sk_flag is reused by the next partial segment
      }
      if (sg.sk) ++nseg;
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
      if (threadIdx.x == 128) trace_mark(p, 6);
    }
  }

  ptx::tc_fence_before();
  if (CL > 1) ptx::cluster_sync_all();   // no CTA leaves while the peer may still multicast into it
  else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
  if (threadIdx.x == 0) trace_mark(p, 7);
}

unsigned long long* g_trace = nullptr;
int g_last_cfg[4] = {0, 0, 0, 0};   // tile width, grid, stream-K units per CTA, data-parallel tiles of the last launch
int g_force_bn = 0;
int g_debug = 0;
int g_cluster = 0;    // 1 = two-CTA multicast variant for 256-wide tiles.  Opt-in: measured +8 % on the multiphase
                      // MLP up-projection (3184x16384x2048), neutral to -1.5 % on the cylinder shapes (the big GEMMs
                      // already sit at the sustained cuBLAS rate, so halving B's L2 traffic buys little)
int g_stream_k = 1;   // 0 = never, 1 = when the cost model says so, 2 = whenever legal (tests)
constexpr int kSkMaxTiles = 16384;                      // arrival counters (64 KB)
thread_local unsigned int* g_sk_counters = nullptr;     // caller-owned stream-K workspace (this thread)
thread_local float* g_sk_ws = nullptr;
thread_local size_t g_sk_ws_bytes = 0;

template <int BN, bool AMN, bool BMN>
int launch(const GemmParams& p, int total_tiles, int num_kb, cudaStream_t stream) {
  using C = Cfg<BN>;
  static bool attr_set[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 16 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN, AMN, BMN, 1>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    if (BN == 256) {
      e = cudaFuncSetAttribute(gemm_bf16_tn_kernel<256, AMN, BMN, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               Cfg<256>::SMEM_BYTES);
      if (e != cudaSuccess) return static_cast<int>(e);
    }
    attr_set[dev] = true;
  }
  (void)num_kb;
  if (BN == 256 && p.cluster == 2) {
    // clusters of two CTAs over M-tile pairs, B halves multicast
    const int pairs = ((p.tiles_m + 1) / 2) * p.tiles_n * p.groups;
    int clusters = num_sms() / 2;
    if (pairs < clusters) clusters = pairs;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * clusters); cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = Cfg<256>::SMEM_BYTES; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl_enabled() && !pdl_take_fence()) ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_bf16_tn_kernel<256, AMN, BMN, 2>, p);
    return static_cast<int>(e != cudaSuccess ? e : cudaGetLastError());
  }
  int grid = total_tiles < num_sms() ? total_tiles : num_sms();
  if (p.units_per_cta > 0) grid = p.sk_grid;
  SEA_LAUNCH((gemm_bf16_tn_kernel<BN, AMN, BMN, 1>), grid, kThreads, C::SMEM_BYTES, stream, p);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace

int make_tmap_bf16_2d(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer,
                      uint64_t ld_elems, uint32_t box_inner, uint32_t box_outer) {
  auto encode = tensor_map_encoder();
  if (encode == nullptr) return SEA_ERR_NO_DEVICE;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims,
                      strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? SEA_OK : SEA_ERR_INVALID;
}

}  // namespace sea

extern "C" void sea_gemm_force_tile_n(int bn) { sea::g_force_bn = bn; }
extern "C" void sea_gemm_stream_k(int mode) { sea::g_stream_k = mode; }
extern "C" void sea_gemm_cluster(int on) { sea::g_cluster = on; }
extern "C" int sea_gemm_set_workspace(void* ws, size_t bytes) {
  using namespace sea;
  if (ws == nullptr || bytes <= 65536 + 2 * BM * 64 * sizeof(float)) {
    g_sk_counters = nullptr; g_sk_ws = nullptr; g_sk_ws_bytes = 0;
    return ws == nullptr ? SEA_OK : SEA_ERR_WORKSPACE;
  }
  if (reinterpret_cast<uintptr_t>(ws) & 255) return SEA_ERR_INVALID;
  g_sk_counters = static_cast<unsigned int*>(ws);
  g_sk_ws = reinterpret_cast<float*>(static_cast<char*>(ws) + 65536);
  g_sk_ws_bytes = bytes - 65536;
  return SEA_OK;
}
extern "C" void sea_gemm_debug_probe(int mode) { sea::g_debug = mode; }
extern "C" void sea_gemm_debug_trace(void* dev_buf) { sea::g_trace = static_cast<unsigned long long*>(dev_buf); }
extern "C" void sea_gemm_last_config(int* out) {
  out[0] = sea::g_last_cfg[0]; out[1] = sea::g_last_cfg[1]; out[2] = sea::g_last_cfg[2]; out[3] = sea::g_last_cfg[3];
}



extern "C" int sea_gemm_bf16_tn(int num_problems, const sea_gemm_problem* probs, int M, int N,
                                int K, sea_stream_t stream) {
  return sea_gemm_bf16_tn_chunked(num_problems, probs, M, N, K, 0, stream);
}

extern "C" int sea_gemm_bf16_tn_chunked(int num_problems, const sea_gemm_problem* probs, int M,
                                        int N, int K, int k_chunk, sea_stream_t stream) {
  using namespace sea;
  if (k_chunk < 0 || (k_chunk % 128) != 0) return SEA_ERR_INVALID;
  if (probs == nullptr || num_problems < 1 || num_problems > kMaxGroups) return SEA_ERR_INVALID;
  if (M <= 0 || N <= 0 || K <= 0) return SEA_ERR_INVALID;
  if ((N % 8) != 0) return SEA_ERR_UNSUPPORTED;  // K may be ragged: TMA zero-fills the tail
  const int mnm = probs[0].mn_major;
  for (int g = 0; g < num_problems; ++g)
    if (probs[g].mn_major != mnm) return SEA_ERR_INVALID;
  if (mnm != 0 && mnm != SEA_GEMM_B_MN && mnm != (SEA_GEMM_A_MN | SEA_GEMM_B_MN)) return SEA_ERR_UNSUPPORTED;
  const bool amn = (mnm & SEA_GEMM_A_MN) != 0, bmn = (mnm & SEA_GEMM_B_MN) != 0;
  if (amn && (M % 8) != 0) return SEA_ERR_UNSUPPORTED;
  int rc = ensure_init();
  if (rc != SEA_OK) return rc;

  // Tile width and decomposition.  Cost per candidate width, in units of one 64-wide k-block of an
  // N = 1 column: tile width / mainloop efficiency of that width (measured, scripts/shape_bench.py: the
  // issuing thread's fixed per-stage cost is amortised over N x BK, so narrow tiles stay below the MMA
  // rate: 0.6 / 0.8 / 0.87 / 0.9 of it at N = 64 / 128 / 192 / 256) times the k-blocks on the critical
  // CTA, plus an epilogue/latency term per tile pass.
  //   data-parallel: ceil(tiles / SMs) whole tiles on the critical CTA              (wave quantisation)
  //   hybrid       : floor(tiles / SMs) whole tiles + an equal share of the ragged remainder's k-blocks
  //                  (stream-K tail, a tile split at most 8 ways) + the partial-tile fix-up
  const int sms = num_sms();
  const bool sk_possible = g_sk_ws != nullptr && g_stream_k != 0 && !(k_chunk > 0 && k_chunk < K);
  int bn = g_force_bn;
  int units_per_cta = 0, dp_tiles = -1, sk_grid = 0;
  {
    const long long tm = (M + BM - 1) / BM;
    const int cand[4] = {64, 128, 192, 256};
    const double eff[4] = {0.6, 0.8, 0.87, 0.9};
    double best = 1e30;
    for (int i = 0; i < 4; ++i) {
      if (g_force_bn != 0 && cand[i] != g_force_bn) continue;
      const long long tiles = tm * ((N + cand[i] - 1) / cand[i]) * num_problems;
      const long long waves = (tiles + sms - 1) / sms;
      const double kb64 = (K + KA - 1) / KA;
      const double epi = 6.0 * cand[i] + 400.0;
      const double tile_cost = cand[i] / eff[i] * kb64 + epi;
      const double cost_dp = waves * tile_cost;
      if (cost_dp < best || (g_force_bn != 0 && dp_tiles < 0)) {
        best = cost_dp; bn = cand[i]; units_per_cta = 0; dp_tiles = static_cast<int>(tiles); sk_grid = 0;
      }
      const int bkc = cand[i] <= 128 ? 128 : 64;
      const long long nkb = (K + bkc - 1) / bkc;
      const long long full = tiles / sms, rem = tiles - full * sms;
      if (sk_possible && rem > 0 && nkb >= 8 && rem <= kSkMaxTiles) {
        const long long units = rem * nkb;
        long long min_upc = (nkb + 7) / 8;           // a tile is split at most ~8 ways
        if (min_upc < 4) min_upc = 4;
        long long grid_sk = full > 0 ? sms : (units / min_upc < sms ? units / min_upc : sms);
        if (grid_sk < 1) grid_sk = 1;
        long long upc = (units + grid_sk - 1) / grid_sk;
        if (upc < min_upc) upc = min_upc;
        if (full == 0) grid_sk = (units + upc - 1) / upc;
        const size_t need = static_cast<size_t>(grid_sk) * 2 * BM * cand[i] * sizeof(float);
        const double cost_sk = full * tile_cost + cand[i] / eff[i] * (upc * (bkc / 64.0)) + 3.0 * epi;
        // CTAs of a stream-K range sit at different K offsets, so panels shared between tiles are no
        // longer fetched in lockstep: without full waves in front, the operands must fit L2 comfortably
        // (measured: 796x2048x16384 drops from 1060 to 780 TFLOP/s otherwise).
        const double operand_mb = 2.0 * num_problems * (static_cast<double>(M) + N) * K / 1e6;
        // With full waves in front only the widest tile is worth it (fewest panel re-reads; narrower
        // tiles measured slower than the model predicts).
        const bool l2_ok = full > 0 ? cand[i] == 256 : operand_mb <= 48.0;
        const double margin = full > 0 ? 1.1 : 1.2;
        if (need <= g_sk_ws_bytes && upc < nkb && (g_stream_k == 2 || (l2_ok && cost_sk * margin < best))) {
          best = g_stream_k == 2 ? -1.0 : cost_sk;   // mode 2 (tests): first legal stream-K candidate wins
          bn = cand[i]; units_per_cta = static_cast<int>(upc);
          dp_tiles = static_cast<int>(full * sms); sk_grid = static_cast<int>(grid_sk);
        }
      }
    }
  }
  if (bn != 64 && bn != 128 && bn != 192 && bn != 256) return SEA_ERR_INVALID;

  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.groups = num_problems;
  p.tiles_m = (M + BM - 1) / BM;
  p.tiles_n = (N + bn - 1) / bn;
  const int bk = bn <= 128 ? 128 : 64;   // Cfg<BN>::BK
  p.chunk_kb = (k_chunk > 0 && k_chunk < K) ? k_chunk / bk : (K + bk - 1) / bk;
  const bool chunked = p.chunk_kb < (K + bk - 1) / bk;
  p.debug = g_debug;
  p.trace = g_trace;
  p.units_per_cta = units_per_cta;
  p.dp_tiles = units_per_cta > 0 ? dp_tiles : 0x7fffffff;
  p.sk_grid = sk_grid;
  p.cluster = (bn == 256 && units_per_cta == 0 && g_cluster != 0 && !(k_chunk > 0 && k_chunk < K) &&
               (M + BM - 1) / BM >= 2) ? 2 : 1;
  p.sk_ws = g_sk_ws;
  p.sk_counters = g_sk_counters;
  p.b_is_static = 1;
  for (int g = 0; g < num_problems; ++g) p.b_is_static &= probs[g].b_is_static != 0;
  for (int g = 0; g < num_problems; ++g) {
    const sea_gemm_problem& q = probs[g];
    const sea_gemm_epilogue& e = q.epi;
    if (q.a == nullptr || q.b == nullptr) return SEA_ERR_INVALID;
    if ((q.lda % 8) != 0 || (q.ldb % 8) != 0) return SEA_ERR_INVALID;
    if (q.lda < (amn ? M : K) || q.ldb < (bmn ? N : K)) return SEA_ERR_INVALID;
    if ((reinterpret_cast<uintptr_t>(q.a) & 15) || (reinterpret_cast<uintptr_t>(q.b) & 15))
      return SEA_ERR_INVALID;
    if (e.out_f32 == nullptr && e.out_bf16 == nullptr && e.out_pre_bf16 == nullptr)
      return SEA_ERR_INVALID;
    // chunked: out_f32 holds the running sum.  residual == out_f32 (in-place accumulate, e.g. dW += ...) is fine: the
    // thread that owns an element reads the residual in the first chunk's epilogue and is the only one to write it
    if (chunked && e.out_f32 == nullptr) return SEA_ERR_INVALID;
    if (e.out_f32 && ((e.ld_out_f32 % 4) || (reinterpret_cast<uintptr_t>(e.out_f32) & 15)))
      return SEA_ERR_INVALID;
    if (e.out_bf16 && ((e.ld_out_bf16 % 8) || (reinterpret_cast<uintptr_t>(e.out_bf16) & 15)))
      return SEA_ERR_INVALID;
    if (e.out_pre_bf16 &&
        ((e.ld_out_pre_bf16 % 8) || (reinterpret_cast<uintptr_t>(e.out_pre_bf16) & 15)))
      return SEA_ERR_INVALID;
    if (e.residual && ((e.ld_residual % 4) || (reinterpret_cast<uintptr_t>(e.residual) & 15)))
      return SEA_ERR_INVALID;
    if (e.gelu_grad_of && ((e.ld_gelu % 8) || (reinterpret_cast<uintptr_t>(e.gelu_grad_of) & 15)))
      return SEA_ERR_INVALID;
    if (e.bias && (reinterpret_cast<uintptr_t>(e.bias) & 15)) return SEA_ERR_INVALID;
    if (e.rope_table != nullptr && e.rope_cols > 0) {
      if (e.head_dim <= 0 || (e.head_dim % 32) || (e.rope_cols % e.head_dim) || e.seq_len <= 0 ||
          e.rope_pos0 < 0 || e.rope_ld < e.rope_pos0 + e.seq_len)
        return SEA_ERR_UNSUPPORTED;
    }
    // K-major operand: boxes of 64 k-columns x BM / BN rows; MN-major: 64 M/N-columns x BK k-rows
    rc = amn ? make_tmap_bf16_2d(&p.tma_a[g], q.a, M, K, q.lda, 64, bk)
             : make_tmap_bf16_2d(&p.tma_a[g], q.a, K, M, q.lda, KA, BM);
    if (rc != SEA_OK) return rc;
    rc = bmn ? make_tmap_bf16_2d(&p.tma_b[g], q.b, N, K, q.ldb, 64, bk)
             : make_tmap_bf16_2d(&p.tma_b[g], q.b, K, N, q.ldb, KA, p.cluster == 2 ? bn / 2 : bn);
    if (rc != SEA_OK) return rc;
    DevEpilogue& d = p.epi[g];
    d.bias = e.bias;
    d.residual = e.residual;
    d.gelu_grad_of = static_cast<const __nv_bfloat16*>(e.gelu_grad_of);
    d.rope_table = (e.rope_cols > 0) ? e.rope_table : nullptr;
    d.out_f32 = e.out_f32;
    d.out_pre_bf16 = static_cast<__nv_bfloat16*>(e.out_pre_bf16);
    d.out_bf16 = static_cast<__nv_bfloat16*>(e.out_bf16);
    d.ld_residual = e.ld_residual;
    d.ld_gelu = e.ld_gelu;
    d.ld_out_f32 = e.ld_out_f32;
    d.ld_out_pre_bf16 = e.ld_out_pre_bf16;
    d.ld_out_bf16 = e.ld_out_bf16;
    d.act = e.act;
    d.rope_cols = e.rope_cols;
    d.head_dim = e.head_dim;
    d.seq_len = e.seq_len;
    d.rope_ld = e.rope_ld;
    d.rope_pos0 = e.rope_pos0;
    d.res_rows = e.res_rows_per_batch; d.res_bs = e.res_batch_stride;
    if (e.dropout_p < 0.f || e.dropout_p >= 1.f || (e.dropout_p > 0.f && chunked)) return SEA_ERR_INVALID;
    d.drop_seed = e.dropout_seed; d.drop_site = e.dropout_site;
    d.drop_thresh = e.dropout_p > 0.f ? static_cast<uint32_t>(static_cast<double>(e.dropout_p) * 4294967296.0) : 0u;
    d.drop_scale = 1.0f / (1.0f - e.dropout_p);
    d.rope_sign = (e.rope_sign == 0.0f) ? 1.0f : e.rope_sign;
    auto al32 = [](const void* ptr, long long ld_elems, int esz) {
      return ptr == nullptr || (((reinterpret_cast<uintptr_t>(ptr) & 31) == 0) && ((ld_elems * esz) % 32 == 0));
    };
    if (e.res_rows_per_batch < 0 || (e.res_rows_per_batch > 0 && (e.res_batch_stride % 4))) return SEA_ERR_INVALID;
    d.vec8 = (N % 16 == 0) && al32(e.residual, e.ld_residual, 4) && (e.res_rows_per_batch == 0 || (e.res_batch_stride * 4) % 32 == 0) && al32(e.out_f32, e.ld_out_f32, 4) &&
             al32(e.out_pre_bf16, e.ld_out_pre_bf16, 2) && al32(e.out_bf16, e.ld_out_bf16, 2);
  }
  const int total = p.tiles_m * p.tiles_n * p.groups;
  const int nkb = (K + bk - 1) / bk;
  g_last_cfg[0] = bn; g_last_cfg[1] = units_per_cta > 0 ? sk_grid : (total < sms ? total : sms);
  g_last_cfg[2] = units_per_cta; g_last_cfg[3] = units_per_cta > 0 ? dp_tiles : total;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
#define SEA_GEMM_DISPATCH(AMN_, BMN_)                                   \
  do {                                                                  \
    if (bn == 256) return launch<256, AMN_, BMN_>(p, total, nkb, s);    \
    if (bn == 192) return launch<192, AMN_, BMN_>(p, total, nkb, s);    \
    if (bn == 128) return launch<128, AMN_, BMN_>(p, total, nkb, s);    \
    return launch<64, AMN_, BMN_>(p, total, nkb, s);                    \
  } while (0)
  if (amn && bmn) SEA_GEMM_DISPATCH(true, true);
  if (bmn) SEA_GEMM_DISPATCH(false, true);
  SEA_GEMM_DISPATCH(false, false);
#undef SEA_GEMM_DISPATCH
}

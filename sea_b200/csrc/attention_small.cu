// K2 for short sequences (T <= 128): the whole (batch, head) problem in one small CTA, warp-level mma.sync.
//
// Why a second forward kernel: the rollout benchmark calls the causal attention 300 times per trajectory-step at
// prefix lengths 1..100.  The tcgen05 kernel (attention_tc.cu) pads every such problem to one 128-query x 128-key tile
// and pays its fixed costs per CTA — TMEM allocation, five mbarriers, three 128-row TMA boxes that are mostly
// zero-fill, a softmax warpgroup waiting on a 128 x 128 MMA — with 100-200 KB of shared memory, i.e. one or two CTAs
// per SM and 2-4 waves for the 256-512 (batch, head) pairs: 9-17 us per launch for a few MFLOP of work
// (profiles/r2_launches_rollout_cylinder.csv.gz).  Here a CTA has one warp per 16 query rows, K / V / Q rows only
// as far as they exist, scores and the output tile in registers (FlashAttention-2 register layout: the S accumulators
// ARE the A operand of P.V), no online softmax (all keys fit).  Same semantics and auxiliary output as the tcgen05
// kernel: key k is visible to query q iff k <= q + src_len (models/base_blocks.py:191-197, 283-289), P is rounded to
// bf16 before P.V, the row sum is taken in fp32 before the rounding, lse = ln sum exp(scaled scores).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/sea_b200.h"
#include "internal.h"
#include "ptx.cuh"

namespace sea {
namespace {

struct SmallItem {
  const __nv_bfloat16* q;
  const __nv_bfloat16* k;
  const __nv_bfloat16* v;
  __nv_bfloat16* o;
  float* lse;
  long long ldq, ldk, ldv;
};
struct SmallParams {
  SmallItem it[SEA_MAX_STREAMS];
  int B, T, H, src_len;
  long long ldo;
  float scale_log2;   // softmax scale * log2(e)
};

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

constexpr int kMaxT = 128;

// One CTA per (batch, head) of problem blockIdx.y; blockDim.x = 32 * ceil(T / 16).
template <int HD>
__global__ void __launch_bounds__(256) attn_fwd_small_kernel(const __grid_constant__ SmallParams p) {
  constexpr int PB = HD * 2 + 16;        // row pitch in bytes: an odd number of 16-byte chunks (ldmatrix conflict-free)
  constexpr int CPR = HD / 8;            // 16-byte chunks per row
  extern __shared__ __align__(128) uint8_t sm_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nwarps = blockDim.x >> 5;
  const int Tp = nwarps * 16;
  uint8_t* Ks = sm_raw;
  uint8_t* Vs = Ks + Tp * PB;
  uint8_t* Qs = Vs + Tp * PB;
  const SmallItem& it = p.it[blockIdx.y];
  const int b = blockIdx.x / p.H, h = blockIdx.x - b * p.H;
  const int T = p.T;
  ptx::pdl_trigger();
  ptx::pdl_wait();

  // ---- stage K, V, Q rows [0, T) of this (batch, head); rows past T are zero (P = 0 times garbage must stay 0)
  {
    const long long row0 = static_cast<long long>(b) * T;
    const __nv_bfloat16* qg = it.q + row0 * it.ldq + h * HD;
    const __nv_bfloat16* kg = it.k + row0 * it.ldk + h * HD;
    const __nv_bfloat16* vg = it.v + row0 * it.ldv + h * HD;
    for (int i = tid; i < Tp * CPR; i += blockDim.x) {
      const int r = i / CPR, c = i - r * CPR;
      const uint32_t off = r * PB + c * 16;
      if (r < T) {
        cp_async16(ptx::smem_u32(Ks + off), kg + r * it.ldk + c * 8);
        cp_async16(ptx::smem_u32(Vs + off), vg + r * it.ldv + c * 8);
        cp_async16(ptx::smem_u32(Qs + off), qg + r * it.ldq + c * 8);
      } else {
        const uint4 z = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(Ks + off) = z;
        *reinterpret_cast<uint4*>(Vs + off) = z;
        *reinterpret_cast<uint4*>(Qs + off) = z;
      }
    }
    asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();

  const int q0 = warp * 16;
  const int g = lane >> 2, qd = lane & 3;
  // keys this warp's rows can see: k <= q0 + 15 + src_len and k < T; in 16-key steps
  int kmax = q0 + 16 + p.src_len;
  if (kmax > T) kmax = T;
  const int nsteps = (kmax + 15) >> 4;   // 1..8, warp-uniform

  // ---- S = Q K^T
  float s[kMaxT / 8][4];
#pragma unroll
  for (int nt = 0; nt < kMaxT / 8; ++nt) { s[nt][0] = 0.f; s[nt][1] = 0.f; s[nt][2] = 0.f; s[nt][3] = 0.f; }
  const uint32_t q_base = ptx::smem_u32(Qs) + (q0 + (lane & 15)) * PB + (lane >> 4) * 16;
  const uint32_t k_base = ptx::smem_u32(Ks) + (((lane >> 4) << 3) + (lane & 7)) * PB + ((lane >> 3) & 1) * 16;
#pragma unroll
  for (int kk = 0; kk < HD / 16; ++kk) {
    uint32_t a0, a1, a2, a3;
    ldsm_x4(q_base + kk * 32, a0, a1, a2, a3);
#pragma unroll
    for (int j = 0; j < kMaxT / 16; ++j) {
      if (j < nsteps) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(k_base + j * 16 * PB + kk * 32, b0, b1, b2, b3);
        mma_bf16(s[2 * j], a0, a1, a2, a3, b0, b1);
        mma_bf16(s[2 * j + 1], a0, a1, a2, a3, b2, b3);
      }
    }
  }

  // ---- mask, softmax (rows g and g + 8 of the warp's block; a quad of lanes shares a row)
  const int qa = q0 + g, qb = q0 + g + 8;
  const int la = min(qa + p.src_len, T - 1), lb = min(qb + p.src_len, T - 1);   // last visible key of each row
  float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < kMaxT / 8; ++nt) {
    if (nt < 2 * nsteps) {
      const int k0 = nt * 8 + qd * 2;
      s[nt][0] = (k0 <= la) ? s[nt][0] * p.scale_log2 : -INFINITY;
      s[nt][1] = (k0 + 1 <= la) ? s[nt][1] * p.scale_log2 : -INFINITY;
      s[nt][2] = (k0 <= lb) ? s[nt][2] * p.scale_log2 : -INFINITY;
      s[nt][3] = (k0 + 1 <= lb) ? s[nt][3] * p.scale_log2 : -INFINITY;
      ma = fmaxf(ma, fmaxf(s[nt][0], s[nt][1]));
      mb = fmaxf(mb, fmaxf(s[nt][2], s[nt][3]));
    }
  }
  ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 1));
  ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 2));
  mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 1));
  mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 2));
  float suma = 0.f, sumb = 0.f;
  uint32_t pk[kMaxT / 8][2];   // P as bf16 pairs: [nt][0] = row g, [nt][1] = row g + 8
#pragma unroll
  for (int nt = 0; nt < kMaxT / 8; ++nt) {
    if (nt < 2 * nsteps) {
      const float p0 = exp2f(s[nt][0] - ma), p1 = exp2f(s[nt][1] - ma);
      const float p2 = exp2f(s[nt][2] - mb), p3 = exp2f(s[nt][3] - mb);
      suma += p0 + p1;
      sumb += p2 + p3;
      pk[nt][0] = ptx::pack_bf16(p0, p1);
      pk[nt][1] = ptx::pack_bf16(p2, p3);
    }
  }
  suma += __shfl_xor_sync(0xffffffffu, suma, 1);
  suma += __shfl_xor_sync(0xffffffffu, suma, 2);
  sumb += __shfl_xor_sync(0xffffffffu, sumb, 1);
  sumb += __shfl_xor_sync(0xffffffffu, sumb, 2);

  // ---- O = P V
  float o[HD / 8][4];
#pragma unroll
  for (int f = 0; f < HD / 8; ++f) { o[f][0] = 0.f; o[f][1] = 0.f; o[f][2] = 0.f; o[f][3] = 0.f; }
  const uint32_t v_base = ptx::smem_u32(Vs) + ((((lane >> 3) & 1) << 3) + (lane & 7)) * PB + (lane >> 4) * 16;
#pragma unroll
  for (int j = 0; j < kMaxT / 16; ++j) {
    if (j < nsteps) {
      const uint32_t a0 = pk[2 * j][0], a1 = pk[2 * j][1], a2 = pk[2 * j + 1][0], a3 = pk[2 * j + 1][1];
#pragma unroll
      for (int f = 0; f < HD / 16; ++f) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(v_base + j * 16 * PB + f * 32, b0, b1, b2, b3);
        mma_bf16(o[2 * f], a0, a1, a2, a3, b0, b1);
        mma_bf16(o[2 * f + 1], a0, a1, a2, a3, b2, b3);
      }
    }
  }

  // ---- O / l -> bf16 through the warp's own (now free) Q rows, then 16-byte stores; log-sum-exp
  const float ia = 1.f / suma, ib = 1.f / sumb;
  __syncwarp();
  uint8_t* stage = Qs + q0 * PB;
#pragma unroll
  for (int f = 0; f < HD / 8; ++f) {
    *reinterpret_cast<uint32_t*>(stage + g * PB + (f * 8 + qd * 2) * 2) = ptx::pack_bf16(o[f][0] * ia, o[f][1] * ia);
    *reinterpret_cast<uint32_t*>(stage + (g + 8) * PB + (f * 8 + qd * 2) * 2) = ptx::pack_bf16(o[f][2] * ib, o[f][3] * ib);
  }
  __syncwarp();
  __nv_bfloat16* og = it.o + (static_cast<long long>(b) * T + q0) * p.ldo + h * HD;
#pragma unroll
  for (int i = 0; i < 16 * CPR / 32; ++i) {
    const int ch = i * 32 + lane;
    const int r = ch / CPR, c = ch - r * CPR;
    if (q0 + r < T) *reinterpret_cast<uint4*>(og + r * p.ldo + c * 8) = *reinterpret_cast<const uint4*>(stage + r * PB + c * 16);
  }
  if (it.lse != nullptr && qd == 0) {
    float* lp = it.lse + (static_cast<long long>(b) * p.H + h) * T;
    if (qa < T) lp[qa] = (ma + log2f(suma)) * 0.69314718055994530942f;
    if (qb < T) lp[qb] = (mb + log2f(sumb)) * 0.69314718055994530942f;
  }
}

template <int HD>
int launch_small(int n, const sea_attn_args* a, cudaStream_t s) {
  constexpr int PB = HD * 2 + 16;
  static bool attr_set[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 16 && !attr_set[dev]) {
    SEA_CUDA_OK(cudaFuncSetAttribute(attn_fwd_small_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * kMaxT * PB));
    attr_set[dev] = true;
  }
  SmallParams p;
  for (int i = 0; i < n; ++i) {
    p.it[i].q = static_cast<const __nv_bfloat16*>(a[i].q);
    p.it[i].k = static_cast<const __nv_bfloat16*>(a[i].k);
    p.it[i].v = static_cast<const __nv_bfloat16*>(a[i].v);
    p.it[i].o = static_cast<__nv_bfloat16*>(a[i].o);
    p.it[i].lse = a[i].lse;
    p.it[i].ldq = a[i].ldq; p.it[i].ldk = a[i].ldk; p.it[i].ldv = a[i].ldv;
  }
  p.B = a->B; p.T = a->T; p.H = a->n_heads; p.src_len = a->src_len; p.ldo = a->ldo;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  const int nwarps = (a->T + 15) / 16;
  const dim3 grid(a->B * a->n_heads, n);
  SEA_LAUNCH((attn_fwd_small_kernel<HD>), grid, 32 * nwarps, static_cast<size_t>(3) * nwarps * 16 * PB, s, p);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace

int g_attn_small = 1;   // 0: never, 1: where it is faster, 2: whenever legal (T <= 128; tests)

// bf16, head_dim 64 / 128, no probability dropout; alignment as for the tcgen05 path (checked by the caller).
// Measured inside a CUDA graph, B = 32, 8 heads (scripts/attn_small_bench.py, profiles/r2_attention_small.md): 4.5 / 5.1 us
// against 7.3 us at T = 4 / 16 (head_dim 128), 3.2 / 3.6 / 5.0 against 6.2 at T = 4 / 16 / 32 (head_dim 64); from T ~ 32 the
// legacy mma.sync rate (a warp-level MMA is ~1/8 of the tcgen05 pipe) costs more than the tcgen05 kernel's fixed
// overheads, so longer prefixes stay there.
bool attention_small_supported(const sea_attn_args* a) {
  if (g_attn_small == 0 || a->prec != SEA_PREC_BF16 || (a->head_dim != 64 && a->head_dim != 128) || a->T > kMaxT ||
      a->dropout_p != 0.f || a->src_len < 0)
    return false;
  return g_attn_small == 2 || a->T <= (a->head_dim == 64 ? 40 : 24);
}

int attention_fwd_small(int n, const sea_attn_args* a, cudaStream_t s) {
  return a->head_dim == 64 ? launch_small<64>(n, a, s) : launch_small<128>(n, a, s);
}

}  // namespace sea

extern "C" void sea_attention_small(int on) { sea::g_attn_small = on; }

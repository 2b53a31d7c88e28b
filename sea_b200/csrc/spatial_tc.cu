// K7' / K8' — the ViT-mesh patch encoder / decoder (models/encoder_decoder.py:75-146) on the tensor cores.
//
// Same decomposition as spatial.cu (one CTA per snapshot, the 64 x Es latent state resident in shared memory from the
// patch gather to the final LayerNorm, weights streamed from L2), but every contraction — the per-group patch MLPs
// (C*g -> Hs -> D and back), q|k|v / projection / MLP of the 12 encoder blocks, and the 8-head attention itself
// (QK^T with the head dim zero-padded to k = 16, P.V with P taken straight from the score accumulators) — is a warp-level
// bf16 mma.sync.m16n8k16 with fp32 accumulation.  At these widths (Es = 32 / 64, head dim 4 / 8, 64 tokens) a
// 128-row tcgen05 tile would be three quarters padding; the warp-level shape fits the problem exactly.
//
// Operand layout trick: the contraction index may be permuted freely as long as A and B use the same permutation, so
// thread t of a quad owns the 8 CONSECUTIVE k elements [32*i + 8t, 32*i + 8t + 8) of two k-steps at once: one 16-byte
// load per operand row feeds two MMAs (no ldmatrix, no transposed copies; B = the nn.Linear weight [N, K] as it lies).
// Activations are bf16 in shared memory with a row pitch == 32 (mod 64) elements, which makes the quad-wise 16-byte
// loads bank-conflict free; the residual stream, LayerNorm statistics, softmax and GELU stay fp32.
//
// Weights are pre-rounded to bf16 once (sea_spatial_pack: q|k|v fused along N) into a caller-owned cache.  n_inp (cells of
// the fullest patch: data dependent, utils/data_processors.py:61-88, train/train_temporal.py:147) is arbitrary: the
// pack step pads every field's cell axis of the patch-MLP weights to a multiple of 16 with zeros, the kernels pad the
// snapshot the same way in shared memory and mask the decoder's stores.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/sea_b200.h"
#include "internal.h"
#include "ptx.cuh"

namespace sea {
namespace {

using bf16 = __nv_bfloat16;
constexpr int P = 64;
constexpr int kThreads = 256;             // 8 warps = 4 row tiles x 2 column groups; two or more CTAs share an SM, so
constexpr int kWarps = kThreads / 32;     // one snapshot's barrier / latency stalls are filled by another's work
constexpr int kColGroups = kWarps / 4;

struct LayerTC {
  const bf16 *qkv_w, *proj_w, *mlp0_w, *mlp3_w;
  const float *qkv_b, *ln1_w, *ln2_w, *mlp0_b, *mlp_ln_w, *mlp_ln_b, *mlp3_b;
};
struct SpatialTC {
  int n_groups, n_fields, C, Cp, Hs, D, n_heads, num_layers;   // Cp: cells per patch padded to a multiple of 16
  int g_first[4], g_count[4];
  const bf16 *enc_w1[4], *enc_w2[4], *dec_w1[4], *dec_w2[4];
  const float *enc_b2[4], *dec_b2[4];
  const float *ln_w, *ln_b, *pe;
  LayerTC layers[16];
};

__host__ __device__ inline int pitch_of(int K) {   // smallest pitch >= K with pitch % 64 == 32 (bf16 elements)
  int p = (K / 64) * 64 + 32;
  return p >= K ? p : p + 64;
}

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Y[64, N] = X[64, K] . W[N, K]^T, X bf16 in shared memory (pitch ldx), W bf16 in global memory (row pitch ldw).
// kWarps = 4 row tiles x kColGroups column groups; a warp walks its n-tiles (8 columns each) NT at a time.
// epi(row, col, v0, v1) receives the two adjacent columns (col even) of one row.  K % 16 == 0, N % 8 == 0.
template <int NT, class Epi>
__device__ __forceinline__ void gemm64(const bf16* __restrict__ X, int ldx, int K, const bf16* __restrict__ W,
                                       long long ldw, int N, Epi&& epi) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = lane & 3, r = lane >> 2;
  const int rt = warp & 3, cg = warp >> 2;
  const int ntiles = N >> 3;
  const bf16* xa = X + (rt * 16 + r) * ldx;
  const bf16* xb = xa + 8 * ldx;
  for (int j0 = cg; j0 < ntiles; j0 += kColGroups * NT) {
    float acc[NT][4];
    const bf16* wp[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
      const int j = min(j0 + kColGroups * i, ntiles - 1);
      wp[i] = W + static_cast<long long>(j * 8 + r) * ldw;
    }
    int k0 = 0;
#pragma unroll 2
    for (; k0 + 32 <= K; k0 += 32) {
      const uint4 alo = *reinterpret_cast<const uint4*>(xa + k0 + t * 8);
      const uint4 ahi = *reinterpret_cast<const uint4*>(xb + k0 + t * 8);
      uint4 b[NT];
#pragma unroll
      for (int i = 0; i < NT; ++i) b[i] = __ldg(reinterpret_cast<const uint4*>(wp[i] + k0 + t * 8));
#pragma unroll
      for (int i = 0; i < NT; ++i) {
        mma16816(acc[i], alo.x, ahi.x, alo.y, ahi.y, b[i].x, b[i].y);
        mma16816(acc[i], alo.z, ahi.z, alo.w, ahi.w, b[i].z, b[i].w);
      }
    }
    if (k0 < K) {   // 16-wide tail: thread t owns k0 + 4t .. 4t + 3
      const uint2 alo = *reinterpret_cast<const uint2*>(xa + k0 + t * 4);
      const uint2 ahi = *reinterpret_cast<const uint2*>(xb + k0 + t * 4);
#pragma unroll
      for (int i = 0; i < NT; ++i) {
        const uint2 b = __ldg(reinterpret_cast<const uint2*>(wp[i] + k0 + t * 4));
        mma16816(acc[i], alo.x, ahi.x, alo.y, ahi.y, b.x, b.y);
      }
    }
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      const int j = j0 + kColGroups * i;
      if (j < ntiles) {
        epi(rt * 16 + r, j * 8 + 2 * t, acc[i][0], acc[i][1]);
        epi(rt * 16 + r + 8, j * 8 + 2 * t, acc[i][2], acc[i][3]);
      }
    }
  }
}

template <class Epi>
__device__ __forceinline__ void gemm64_any(const bf16* X, int ldx, int K, const bf16* W, long long ldw, int N, Epi&& epi) {
  if (N >= 128) gemm64<4>(X, ldx, K, W, ldw, N, epi);
  else gemm64<1>(X, ldx, K, W, ldw, N, epi);
}

// Row LayerNorm of the fp32 state (warp per row, the row in registers: NPL = d / 32 values per lane, two-pass
// statistics): Y (bf16, pitch ldy) = (x - mean) * rstd * w (+ b), optional GELU (tanh-fitted form: the result is rounded to
// bf16 next).  `Y` may alias `X` when the bf16 row fits in front of its own fp32 row (a warp reads its whole row first).
template <int NPL>
__device__ __forceinline__ void layernorm_rows_t(const float* X, int ldx, const float* __restrict__ w,
                                                 const float* __restrict__ b, bf16* Y, int ldy, bool gelu) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr float inv_d = 1.0f / (32 * NPL);
  float wv[NPL], bv[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) { wv[i] = __ldg(w + lane + 32 * i); bv[i] = b ? __ldg(b + lane + 32 * i) : 0.f; }
  for (int row = warp; row < P; row += kWarps) {
    const float* xr = X + row * ldx;
    float v[NPL], s = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) { v[i] = xr[lane + 32 * i]; s += v[i]; }
    const float mean = warp_sum(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) { v[i] -= mean; q = fmaf(v[i], v[i], q); }
    const float rstd = rsqrtf(fmaf(warp_sum(q), inv_d, 1e-5f));
    __syncwarp();   // in-place use: every lane has read its fp32 values before any bf16 value lands on them
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const float y = fmaf(v[i] * rstd, wv[i], bv[i]);
      Y[row * ldy + lane + 32 * i] = __float2bfloat16_rn(gelu ? ptx::gelu_bf16(y) : y);
    }
  }
}
__device__ __forceinline__ void layernorm_rows(const float* X, int ldx, int d, const float* __restrict__ w,
                                               const float* __restrict__ b, bf16* Y, int ldy, bool gelu) {
  switch (d) {
    case 32: layernorm_rows_t<1>(X, ldx, w, b, Y, ldy, gelu); return;
    case 64: layernorm_rows_t<2>(X, ldx, w, b, Y, ldy, gelu); return;
    case 96: layernorm_rows_t<3>(X, ldx, w, b, Y, ldy, gelu); return;
    case 128: layernorm_rows_t<4>(X, ldx, w, b, Y, ldy, gelu); return;
    case 192: layernorm_rows_t<6>(X, ldx, w, b, Y, ldy, gelu); return;
    case 256: layernorm_rows_t<8>(X, ldx, w, b, Y, ldy, gelu); return;
    default: break;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float inv_d = 1.0f / d;
  for (int row = warp; row < P; row += kWarps) {
    const float* xr = X + row * ldx;
    float s = 0.f;
    for (int c = lane; c < d; c += 32) s += xr[c];
    const float mean = warp_sum(s) * inv_d;
    float q = 0.f;
    for (int c = lane; c < d; c += 32) { const float e = xr[c] - mean; q = fmaf(e, e, q); }
    const float rstd = rsqrtf(fmaf(warp_sum(q), inv_d, 1e-5f));
    for (int c = lane; c < d; c += 32) {
      float y = (xr[c] - mean) * rstd * __ldg(w + c);
      if (b) y += __ldg(b + c);
      Y[row * ldy + c] = __float2bfloat16_rn(gelu ? ptx::gelu_bf16(y) : y);
    }
  }
}

// Non-causal attention over the 64 patches for all heads (models/base_blocks.py:105-121) on the tensor cores.
// Q, K: bf16 [64][ld] (column h*HD + d); Vt: bf16 [Es (+8)][ldv], row h*HD + d, column = key; O: bf16 [64][ld].
// One (head, 16-query tile) per warp pass: S = Q K^T (head dim zero-padded to k = 16), softmax in registers
// (a query row lives in one quad), P re-used as the A operand of P.V from the accumulator registers.
template <int HD>
__device__ __forceinline__ void attention_tc(const bf16* __restrict__ Q, const bf16* __restrict__ Kk, int ld,
                                             const bf16* __restrict__ Vt, int ldv, int n_heads, bf16* __restrict__ O) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = lane & 3, r = lane >> 2;
  const float sl2 = rsqrtf(static_cast<float>(HD)) * 1.4426950408889634f;
  constexpr int NTV = HD > 8 ? HD / 8 : 1;
  for (int item = warp; item < n_heads * 4; item += kWarps) {
    const int h = item >> 2, rt = item & 3;
    // the 16-dim group of the k = 16 MMA that contains this head's HD dims: Q and K fragments are loaded unconditionally
    // from it, and the lanes of the A fragment that belong to other heads are zeroed (S = A_h . K_group^T is exact)
    const int gbase = (h * HD) & ~15;
    const bf16* q0 = Q + (rt * 16 + r) * ld + gbase;
    const int dlo = gbase + 2 * t - h * HD, dhi = dlo + 8;
    const bool lo_ok = dlo >= 0 && dlo < HD, hi_ok = dhi >= 0 && dhi < HD;
    const uint32_t a0 = lo_ok ? *reinterpret_cast<const uint32_t*>(q0 + 2 * t) : 0u;
    const uint32_t a1 = lo_ok ? *reinterpret_cast<const uint32_t*>(q0 + 8 * ld + 2 * t) : 0u;
    const uint32_t a2 = hi_ok ? *reinterpret_cast<const uint32_t*>(q0 + 2 * t + 8) : 0u;
    const uint32_t a3 = hi_ok ? *reinterpret_cast<const uint32_t*>(q0 + 8 * ld + 2 * t + 8) : 0u;
    float s[8][4];
    const bf16* kr = Kk + r * ld + gbase + 2 * t;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kr + j * 8 * ld);
      const uint32_t b1 = *reinterpret_cast<const uint32_t*>(kr + j * 8 * ld + 8);
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      mma16816(s[j], a0, a1, a2, a3, b0, b1);
    }
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      m0 = fmaxf(m0, fmaxf(s[j][0], s[j][1]));
      m1 = fmaxf(m1, fmaxf(s[j][2], s[j][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float l0 = 0.f, l1 = 0.f;
    const float nm0 = -m0 * sl2, nm1 = -m1 * sl2;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = ptx::ex2(fmaf(s[j][0], sl2, nm0)); s[j][1] = ptx::ex2(fmaf(s[j][1], sl2, nm0));
      s[j][2] = ptx::ex2(fmaf(s[j][2], sl2, nm1)); s[j][3] = ptx::ex2(fmaf(s[j][3], sl2, nm1));
      l0 += s[j][0] + s[j][1];
      l1 += s[j][2] + s[j][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    float o[NTV][4];
#pragma unroll
    for (int nt = 0; nt < NTV; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const uint32_t p0 = ptx::pack_bf16(s[2 * ks][0], s[2 * ks][1]);
      const uint32_t p1 = ptx::pack_bf16(s[2 * ks][2], s[2 * ks][3]);
      const uint32_t p2 = ptx::pack_bf16(s[2 * ks + 1][0], s[2 * ks + 1][1]);
      const uint32_t p3 = ptx::pack_bf16(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
      for (int nt = 0; nt < NTV; ++nt) {
        const bf16* vr = Vt + (h * HD + nt * 8 + r) * ldv + ks * 16 + 2 * t;   // rows past the head: discarded columns
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(vr);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(vr + 8);
        mma16816(o[nt], p0, p1, p2, p3, b0, b1);
      }
    }
    float i0, i1;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(i0) : "f"(l0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(i1) : "f"(l1));
#pragma unroll
    for (int nt = 0; nt < NTV; ++nt) {
      const int d = nt * 8 + 2 * t;
      if (d < HD) {
        bf16* orow = O + (rt * 16 + r) * ld + h * HD + d;
        *reinterpret_cast<uint32_t*>(orow) = ptx::pack_bf16(o[nt][0] * i0, o[nt][1] * i0);
        *reinterpret_cast<uint32_t*>(orow + 8 * ld) = ptx::pack_bf16(o[nt][2] * i1, o[nt][3] * i1);
      }
    }
  }
}

__device__ __forceinline__ void attention_dispatch_tc(int hd, const bf16* Q, const bf16* Kk, int ld, const bf16* Vt, int ldv,
                                                      int n_heads, bf16* O) {
  switch (hd) {
    case 2: attention_tc<2>(Q, Kk, ld, Vt, ldv, n_heads, O); break;
    case 4: attention_tc<4>(Q, Kk, ld, Vt, ldv, n_heads, O); break;
    case 8: attention_tc<8>(Q, Kk, ld, Vt, ldv, n_heads, O); break;
    default: attention_tc<16>(Q, Kk, ld, Vt, ldv, n_heads, O); break;
  }
}

__device__ __forceinline__ long long latent_index(int b, int p, int g, int d, int G, int D, int layout) {
  return layout == 0 ? ((static_cast<long long>(b) * P + p) * G + g) * D + d
                     : ((static_cast<long long>(b) * G + g) * P + p) * D + d;
}

// Shared-memory plan of the encoder (bytes, 16-byte aligned regions).  The snapshot is staged one field GROUP at a
// time (bf16), the patch-MLP hidden layer is processed in three column chunks (the second Linear accumulates over
// them), the MLP's fp32 pre-LN hidden reuses the q | k | v^T area plus the patch-MLP area (both dead by then) and its
// bf16 GELU output overwrites the front of its own fp32 row: a cylinder_flow snapshot needs 67 KB (three CTAs per SM),
// a multiphase_flow one 112 KB (two).
struct EncPlan { int ld_in, ld_hid, ld_e, ld_v, ld_hh, hchunk; size_t off_z, off_n, off_q, off_k, off_v, off_a, a_hid, a_hh, total; };
__host__ __device__ inline EncPlan enc_plan(int group_width, int Hs, int Es) {
  EncPlan p{};
  p.hchunk = ((Hs + 2) / 3 + 15) / 16 * 16;
  p.ld_in = pitch_of(group_width); p.ld_hid = pitch_of(p.hchunk); p.ld_e = pitch_of(Es);
  p.ld_v = P + 8;          // 36 words: the 8 value rows of a B fragment fall on distinct banks
  p.ld_hh = 4 * Es + 16;   // fp32 pitch; read as bf16 the same rows have pitch 2 * ld_hh == 32 (mod 64)
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t a = o; o += (bytes + 15) & ~static_cast<size_t>(15); return a; };
  p.off_z = take(sizeof(float) * P * Es);
  p.off_n = take(sizeof(bf16) * P * p.ld_e);
  p.off_q = take(sizeof(bf16) * P * p.ld_e);
  p.off_k = take(sizeof(bf16) * P * p.ld_e);
  p.off_v = take(sizeof(bf16) * (Es + 8) * p.ld_v);
  p.off_a = take(sizeof(bf16) * P * p.ld_in);
  p.a_hid = take(sizeof(bf16) * P * p.ld_hid);
  p.a_hh = p.off_q;
  const size_t end_hh = p.a_hh + sizeof(float) * P * p.ld_hh;
  p.total = o > end_hh ? o : end_hh;
  return p;
}

// ------------------------------------------------------------------------------------ encoder
__global__ void __launch_bounds__(kThreads, 3) spatial_encode_tc_kernel(const SpatialTC a, float* __restrict__ x,
                                                                        float* __restrict__ z, int layout, float pad_idx,
                                                                        int fix_pad) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  extern __shared__ __align__(16) unsigned char smraw[];
  const int FC = a.n_fields * a.C, Es = a.n_groups * a.D, hd = Es / a.n_heads;
  int gmax = 0;
  for (int g = 0; g < a.n_groups; ++g) gmax = max(gmax, a.g_count[g]);
  const EncPlan pl = enc_plan(gmax * a.Cp, a.Hs, Es);
  float* Z = reinterpret_cast<float*>(smraw + pl.off_z);     // [64][Es] fp32 residual state
  bf16* Nn = reinterpret_cast<bf16*>(smraw + pl.off_n);      // [64][ld_e] normed input / attention output
  bf16* Qs = reinterpret_cast<bf16*>(smraw + pl.off_q);
  bf16* Ks = reinterpret_cast<bf16*>(smraw + pl.off_k);
  bf16* Vt = reinterpret_cast<bf16*>(smraw + pl.off_v);      // [Es + 8][ld_v] values, transposed
  bf16* Xin = reinterpret_cast<bf16*>(smraw + pl.off_a);     // [64][ld_in] one field group of the snapshot
  bf16* Hid = reinterpret_cast<bf16*>(smraw + pl.a_hid);     // [64][ld_hid]
  float* Hh = reinterpret_cast<float*>(smraw + pl.a_hh);     // [64][ld_hh] fp32 (pre-LN MLP hidden), over q | k | v^T | Xin..
  bf16* Hb = reinterpret_cast<bf16*>(Hh);                    // its GELU(LN(.)) in bf16, in place: pitch 2 * ld_hh
  const int b = blockIdx.x;
  float* xb = x + static_cast<long long>(b) * P * FC;
  // (1) generate_padding_mask (models/encoder_decoder.py:173-176): x[x == pad_idx] = 0, in place, whole snapshot
  if (fix_pad) {
    for (int i = threadIdx.x * 4; i < P * FC; i += kThreads * 4) {   // P * FC is a multiple of 64: aligned float4
      float4 v = *reinterpret_cast<const float4*>(xb + i);
      if (v.x == pad_idx || v.y == pad_idx || v.z == pad_idx || v.w == pad_idx) {
        v.x = v.x == pad_idx ? 0.f : v.x; v.y = v.y == pad_idx ? 0.f : v.y;
        v.z = v.z == pad_idx ? 0.f : v.z; v.w = v.w == pad_idx ? 0.f : v.w;
        *reinterpret_cast<float4*>(xb + i) = v;
      }
    }
    __syncthreads();
  }
  // (2) per-group patch MLP: Linear(C*g, Hs, no bias) -> GELU -> Linear(Hs, D) + b, + positional encoding (:108-114)
  for (int g = 0; g < a.n_groups; ++g) {
    const int cnt = a.g_count[g], Kin = cnt * a.Cp, gw = cnt * a.C;
    const float* xg = xb + a.g_first[g] * a.C;        // row p of the group: xg + p * FC, gw contiguous floats
    if ((a.C & 3) == 0) {
      const int q4 = gw >> 2;
      for (int i = threadIdx.x; i < P * q4; i += kThreads) {
        const int p = i / q4, c4 = (i - p * q4) * 4;
        const float4 v = *reinterpret_cast<const float4*>(xg + p * FC + c4);
        const int f = c4 / a.C, c = c4 - f * a.C;      // C % 4 == 0: the four cells share a field
        *reinterpret_cast<uint2*>(Xin + p * pl.ld_in + f * a.Cp + c) = make_uint2(ptx::pack_bf16(v.x, v.y), ptx::pack_bf16(v.z, v.w));
      }
    } else {
      for (int i = threadIdx.x; i < P * gw; i += kThreads) {
        const int p = i / gw, cc = i - p * gw;
        const int f = cc / a.C, c = cc - f * a.C;
        Xin[p * pl.ld_in + f * a.Cp + c] = __float2bfloat16_rn(xg[p * FC + cc]);
      }
    }
    if (a.Cp != a.C) {   // zero the padded cells (their weight columns are zero too, but smem garbage may be NaN)
      const int padw = a.Cp - a.C;
      for (int i = threadIdx.x; i < P * cnt * padw; i += kThreads) {
        const int pf = i / padw, c = a.C + i - pf * padw;
        const int p = pf / cnt, f = pf - p * cnt;
        Xin[p * pl.ld_in + f * a.Cp + c] = __float2bfloat16_rn(0.f);
      }
    }
    __syncthreads();
    const float* b2 = a.enc_b2[g];
    for (int c0 = 0; c0 < a.Hs; c0 += pl.hchunk) {
      const int ch = min(pl.hchunk, a.Hs - c0);
      gemm64_any(Xin, pl.ld_in, Kin, a.enc_w1[g] + static_cast<long long>(c0) * Kin, Kin, ch,
                 [&](int row, int col, float v0, float v1) {
                   *reinterpret_cast<uint32_t*>(Hid + row * pl.ld_hid + col) = ptx::pack_bf16(ptx::gelu_bf16(v0), ptx::gelu_bf16(v1));
                 });
      __syncthreads();
      gemm64_any(Hid, pl.ld_hid, ch, a.enc_w2[g] + c0, a.Hs, a.D, [&](int row, int col, float v0, float v1) {
        const int c = g * a.D + col;
        float2* zp = reinterpret_cast<float2*>(Z + row * Es + c);
        if (c0 == 0) {
          const float2 pe = __ldg(reinterpret_cast<const float2*>(a.pe + row * Es + c));
          *zp = make_float2(v0 + __ldg(b2 + col) + pe.x, v1 + __ldg(b2 + col + 1) + pe.y);
        } else {
          const float2 zv = *zp;
          *zp = make_float2(zv.x + v0, zv.y + v1);
        }
      });
      __syncthreads();
    }
  }
  // (3) encoder blocks (base_blocks.py:123-138)
  for (int l = 0; l < a.num_layers; ++l) {
    const LayerTC& L = a.layers[l];
    layernorm_rows(Z, Es, Es, L.ln1_w, nullptr, Nn, pl.ld_e, false);
    // rows past the last head of v^T feed discarded accumulator columns only, but must not hold NaN bit patterns
    // (the area is shared with the previous layer's fp32 MLP hidden)
    for (int i = threadIdx.x; i < 8 * pl.ld_v; i += kThreads) Vt[Es * pl.ld_v + i] = __float2bfloat16_rn(0.f);
    __syncthreads();
    gemm64_any(Nn, pl.ld_e, Es, L.qkv_w, Es, 3 * Es, [&](int row, int col, float v0, float v1) {
      v0 += __ldg(L.qkv_b + col); v1 += __ldg(L.qkv_b + col + 1);
      if (col < Es) *reinterpret_cast<uint32_t*>(Qs + row * pl.ld_e + col) = ptx::pack_bf16(v0, v1);
      else if (col < 2 * Es) *reinterpret_cast<uint32_t*>(Ks + row * pl.ld_e + col - Es) = ptx::pack_bf16(v0, v1);
      else {
        Vt[(col - 2 * Es) * pl.ld_v + row] = __float2bfloat16_rn(v0);
        Vt[(col - 2 * Es + 1) * pl.ld_v + row] = __float2bfloat16_rn(v1);
      }
    });
    __syncthreads();
    attention_dispatch_tc(hd, Qs, Ks, pl.ld_e, Vt, pl.ld_v, a.n_heads, Nn);
    __syncthreads();
    gemm64_any(Nn, pl.ld_e, Es, L.proj_w, Es, Es, [&](int row, int col, float v0, float v1) {
      float2* zp = reinterpret_cast<float2*>(Z + row * Es + col);
      const float2 zv = *zp;
      *zp = make_float2(zv.x + v0, zv.y + v1);
    });
    __syncthreads();
    layernorm_rows(Z, Es, Es, L.ln2_w, nullptr, Nn, pl.ld_e, false);
    __syncthreads();
    gemm64_any(Nn, pl.ld_e, Es, L.mlp0_w, Es, 4 * Es, [&](int row, int col, float v0, float v1) {
      *reinterpret_cast<float2*>(Hh + row * pl.ld_hh + col) = make_float2(v0 + __ldg(L.mlp0_b + col), v1 + __ldg(L.mlp0_b + col + 1));
    });
    __syncthreads();
    layernorm_rows(Hh, pl.ld_hh, 4 * Es, L.mlp_ln_w, L.mlp_ln_b, Hb, 2 * pl.ld_hh, true);
    __syncthreads();
    gemm64_any(Hb, 2 * pl.ld_hh, 4 * Es, L.mlp3_w, 4 * Es, Es, [&](int row, int col, float v0, float v1) {
      float2* zp = reinterpret_cast<float2*>(Z + row * Es + col);
      const float2 zv = *zp;
      *zp = make_float2(zv.x + v0 + __ldg(L.mlp3_b + col), zv.y + v1 + __ldg(L.mlp3_b + col + 1));
    });
    __syncthreads();
  }
  // (4) final nn.LayerNorm(Es) (fp32) and the latent store (optionally already in the temporal layout)
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int row = warp; row < P; row += kWarps) {
      const float* xr = Z + row * Es;
      float s = 0.f;
      for (int c = lane; c < Es; c += 32) s += xr[c];
      const float mean = warp_sum(s) / Es;
      float q = 0.f;
      for (int c = lane; c < Es; c += 32) { const float e = xr[c] - mean; q = fmaf(e, e, q); }
      const float rstd = rsqrtf(warp_sum(q) / Es + 1e-5f);
      for (int c = lane; c < Es; c += 32) {
        const int g = c / a.D, d = c - g * a.D;
        z[latent_index(b, row, g, d, a.n_groups, a.D, layout)] = (xr[c] - mean) * rstd * __ldg(a.ln_w + c) + __ldg(a.ln_b + c);
      }
    }
  }
}

// ------------------------------------------------------------------------------------ decoder
__global__ void __launch_bounds__(kThreads, 3) spatial_decode_tc_kernel(const SpatialTC a, const float* __restrict__ z,
                                                                        float* __restrict__ out, int layout) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  extern __shared__ __align__(16) unsigned char smraw[];
  const int FC = a.n_fields * a.C, Es = a.n_groups * a.D;
  const int ld_e = pitch_of(Es), ld_hid = pitch_of(a.Hs);
  bf16* Zb = reinterpret_cast<bf16*>(smraw);                                  // [64][ld_e]
  bf16* Hid = Zb + ((P * ld_e + 7) & ~7);                                     // [64][ld_hid]
  const int b = blockIdx.x;
  float* ob = out + static_cast<long long>(b) * P * FC;
  for (int i = threadIdx.x; i < P * Es; i += kThreads) {
    const int p = i / Es, c = i - p * Es, g = c / a.D, d = c - g * a.D;
    Zb[p * ld_e + c] = __float2bfloat16_rn(z[latent_index(b, p, g, d, a.n_groups, a.D, layout)]);
  }
  __syncthreads();
  // per group: Linear(D, Hs, no bias) -> GELU -> Linear(Hs, C*g) + b   (encoder_decoder.py:140-143)
  for (int g = 0; g < a.n_groups; ++g) {
    const int Nout = a.g_count[g] * a.Cp;   // padded: columns with cell index >= C are masked below
    gemm64_any(Zb + g * a.D, ld_e, a.D, a.dec_w1[g], a.D, a.Hs, [&](int row, int col, float v0, float v1) {
      *reinterpret_cast<uint32_t*>(Hid + row * ld_hid + col) = ptx::pack_bf16(ptx::gelu_bf16(v0), ptx::gelu_bf16(v1));
    });
    __syncthreads();
    const float* b2 = a.dec_b2[g];
    float* og = ob + a.g_first[g] * a.C;
    const bool pair_ok = (a.C & 1) == 0;
    gemm64_any(Hid, ld_hid, a.Hs, a.dec_w2[g], a.Hs, Nout, [&](int row, int col, float v0, float v1) {
      const int fi = col / a.Cp, c = col - fi * a.Cp;   // Cp even: both columns belong to the same field
      if (c >= a.C) return;
      const int n = fi * a.C + c;
      float* dst = og + row * FC + n;
      if (pair_ok) {   // C even: c + 1 < C as well, and the address is 8-byte aligned
        *reinterpret_cast<float2*>(dst) = make_float2(v0 + __ldg(b2 + n), v1 + __ldg(b2 + n + 1));
      } else {
        dst[0] = v0 + __ldg(b2 + n);
        if (c + 1 < a.C) dst[1] = v1 + __ldg(b2 + n + 1);
      }
    });
    __syncthreads();
  }
}

// fp32 [R, G*C] -> bf16 [R, G*Cp] (pad_rows = 0: every field's C columns padded to Cp with zeros) or
// fp32 [G*C, Kc] -> bf16 [G*Cp, Kc] (pad_rows = 1: every field's C rows padded to Cp zero rows).
__global__ void __launch_bounds__(256) pad_cast_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int R, int G,
                                                       int C, int Cp, int Kc, int pad_rows) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const long long total = pad_rows ? static_cast<long long>(G) * Cp * Kc : static_cast<long long>(R) * G * Cp;
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
    float v = 0.f;
    if (pad_rows) {
      const long long rp = i / Kc; const int k = static_cast<int>(i - rp * Kc);
      const int f = static_cast<int>(rp / Cp), c = static_cast<int>(rp - static_cast<long long>(f) * Cp);
      if (c < C) v = src[(static_cast<long long>(f) * C + c) * Kc + k];
    } else {
      const long long r = i / (static_cast<long long>(G) * Cp); const int cp = static_cast<int>(i - r * G * Cp);
      const int f = cp / Cp, c = cp - f * Cp;
      if (c < C) v = src[r * G * C + f * C + c];
    }
    dst[i] = __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------------------------------------ host side
struct CacheMap {
  size_t enc_w1[4], enc_w2[4], dec_w1[4], dec_w2[4];
  size_t qkv_w[16], proj_w[16], mlp0_w[16], mlp3_w[16], qkv_b[16];
  size_t total;
};

int check_desc(const sea_spatial_desc* d) {
  if (!d || d->n_groups < 1 || d->n_groups > 4 || d->num_layers < 0 || d->num_layers > 16) return SEA_ERR_UNSUPPORTED;
  if (d->n_patches != P || d->n_heads < 1) return SEA_ERR_UNSUPPORTED;
  const int Es = d->n_groups * d->embed_dim;
  if (Es % d->n_heads) return SEA_ERR_UNSUPPORTED;
  const int hd = Es / d->n_heads;
  if (hd != 2 && hd != 4 && hd != 8 && hd != 16) return SEA_ERR_UNSUPPORTED;
  // contraction widths must be multiples of 16 (one k-step), output widths multiples of 8 (one n-tile)
  // (n_inp is arbitrary: its axis is padded to a multiple of 16 at pack time)
  // (the MLP's in-place LayerNorm needs its 4*Es-wide row in registers: Es in {16, 32, 48, 64})
  if ((d->embed_dim % 16) || (d->mlp_hidden % 16) || (Es % 16) || Es > 64 || d->n_inp < 1) return SEA_ERR_UNSUPPORTED;
  for (int g = 0; g < d->n_groups; ++g) {
    const int cnt = d->group_num_fields[g], first = d->group_first_field[g];
    if (cnt < 1 || first < 0 || first + cnt > d->n_fields) return SEA_ERR_INVALID;
  }
  return SEA_OK;
}

void map_cache(const sea_spatial_desc* d, CacheMap& m) {
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t a = o; o += (bytes + 255) & ~static_cast<size_t>(255); return a; };
  const size_t Hs = d->mlp_hidden, D = d->embed_dim, Es = static_cast<size_t>(d->n_groups) * D;
  for (int g = 0; g < d->n_groups; ++g) {
    const size_t Kg = static_cast<size_t>(d->group_num_fields[g]) * ((d->n_inp + 15) / 16 * 16);
    m.enc_w1[g] = take(2 * Hs * Kg); m.enc_w2[g] = take(2 * D * Hs);
    m.dec_w1[g] = take(2 * Hs * D); m.dec_w2[g] = take(2 * Kg * Hs);
  }
  for (int l = 0; l < d->num_layers; ++l) {
    m.qkv_w[l] = take(2 * 3 * Es * Es); m.proj_w[l] = take(2 * Es * Es);
    m.mlp0_w[l] = take(2 * 4 * Es * Es); m.mlp3_w[l] = take(2 * 4 * Es * Es);
    m.qkv_b[l] = take(4 * 3 * Es);
  }
  m.total = o + 256;
}

int fill_tc(const sea_spatial_desc* d, const void* cache, SpatialTC& a) {
  int rc = check_desc(d);
  if (rc) return rc;
  if (!cache || (reinterpret_cast<uintptr_t>(cache) & 255)) return SEA_ERR_INVALID;
  CacheMap m;
  map_cache(d, m);
  const char* base = static_cast<const char*>(cache);
  a.n_groups = d->n_groups; a.n_fields = d->n_fields; a.C = d->n_inp; a.Cp = (d->n_inp + 15) / 16 * 16; a.Hs = d->mlp_hidden;
  a.D = d->embed_dim; a.n_heads = d->n_heads; a.num_layers = d->num_layers;
  for (int g = 0; g < d->n_groups; ++g) {
    a.g_first[g] = d->group_first_field[g]; a.g_count[g] = d->group_num_fields[g];
    a.enc_w1[g] = reinterpret_cast<const bf16*>(base + m.enc_w1[g]); a.enc_w2[g] = reinterpret_cast<const bf16*>(base + m.enc_w2[g]);
    a.dec_w1[g] = reinterpret_cast<const bf16*>(base + m.dec_w1[g]); a.dec_w2[g] = reinterpret_cast<const bf16*>(base + m.dec_w2[g]);
    a.enc_b2[g] = d->enc_b2[g]; a.dec_b2[g] = d->dec_b2[g];
  }
  a.ln_w = d->ln_w; a.ln_b = d->ln_b; a.pe = d->pe;
  for (int l = 0; l < d->num_layers; ++l) {
    const sea_spatial_layer& s = d->layers[l];
    a.layers[l] = LayerTC{reinterpret_cast<const bf16*>(base + m.qkv_w[l]), reinterpret_cast<const bf16*>(base + m.proj_w[l]),
                          reinterpret_cast<const bf16*>(base + m.mlp0_w[l]), reinterpret_cast<const bf16*>(base + m.mlp3_w[l]),
                          reinterpret_cast<const float*>(base + m.qkv_b[l]), s.ln1_w, s.ln2_w, s.mlp0_b, s.mlp_ln_w,
                          s.mlp_ln_b, s.mlp3_b};
  }
  return SEA_OK;
}

}  // namespace
}  // namespace sea

using namespace sea;

extern "C" size_t sea_spatial_cache_bytes(const sea_spatial_desc* d) {
  if (check_desc(d) != SEA_OK) return 0;
  CacheMap m;
  map_cache(d, m);
  return m.total;
}

extern "C" int sea_spatial_pack(const sea_spatial_desc* d, void* cache, size_t cache_bytes, sea_stream_t stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  if (!cache || (reinterpret_cast<uintptr_t>(cache) & 255)) return SEA_ERR_INVALID;
  CacheMap m;
  map_cache(d, m);
  if (cache_bytes < m.total) return SEA_ERR_WORKSPACE;
  char* base = static_cast<char*>(cache);
  const int64_t Hs = d->mlp_hidden, D = d->embed_dim, Es = static_cast<int64_t>(d->n_groups) * D;
  auto cast = [&](const float* src, size_t off, int64_t n) -> int {
    if (!src) return SEA_ERR_INVALID;
    return sea_cast_f32_bf16(src, base + off, n, stream);
  };
#define SEA_TRY_(e) do { int _rc = (e); if (_rc != SEA_OK) return _rc; } while (0)
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int Cn = d->n_inp, Cp = (Cn + 15) / 16 * 16;
  for (int g = 0; g < d->n_groups; ++g) {
    const int cnt = d->group_num_fields[g];
    if (d->enc_w1[g]) {
      if (!d->enc_w2[g]) return SEA_ERR_INVALID;
      SEA_LAUNCH(pad_cast_kernel, 296, 256, 0, s, d->enc_w1[g], reinterpret_cast<bf16*>(base + m.enc_w1[g]), static_cast<int>(Hs),
                 cnt, Cn, Cp, 0, 0);
      SEA_TRY_(cast(d->enc_w2[g], m.enc_w2[g], D * Hs));
    }
    if (d->dec_w1[g]) {
      if (!d->dec_w2[g]) return SEA_ERR_INVALID;
      SEA_TRY_(cast(d->dec_w1[g], m.dec_w1[g], Hs * D));
      SEA_LAUNCH(pad_cast_kernel, 296, 256, 0, s, d->dec_w2[g], reinterpret_cast<bf16*>(base + m.dec_w2[g]), 0, cnt, Cn, Cp,
                 static_cast<int>(Hs), 1);
    }
  }
  SEA_CUDA_OK(cudaGetLastError());
  if (d->num_layers > 0 && !d->layers) return SEA_ERR_INVALID;
  for (int l = 0; l < d->num_layers; ++l) {
    const sea_spatial_layer& L = d->layers[l];
    SEA_TRY_(cast(L.q_w, m.qkv_w[l], Es * Es));
    SEA_TRY_(cast(L.k_w, m.qkv_w[l] + 2 * Es * Es, Es * Es));
    SEA_TRY_(cast(L.v_w, m.qkv_w[l] + 4 * Es * Es, Es * Es));
    SEA_TRY_(cast(L.proj_w, m.proj_w[l], Es * Es));
    SEA_TRY_(cast(L.mlp0_w, m.mlp0_w[l], 4 * Es * Es));
    SEA_TRY_(cast(L.mlp3_w, m.mlp3_w[l], 4 * Es * Es));
    if (!L.q_b || !L.k_b || !L.v_b) return SEA_ERR_INVALID;
    SEA_CUDA_OK(cudaMemcpyAsync(base + m.qkv_b[l], L.q_b, 4 * Es, cudaMemcpyDeviceToDevice, s));
    SEA_CUDA_OK(cudaMemcpyAsync(base + m.qkv_b[l] + 4 * Es, L.k_b, 4 * Es, cudaMemcpyDeviceToDevice, s));
    SEA_CUDA_OK(cudaMemcpyAsync(base + m.qkv_b[l] + 8 * Es, L.v_b, 4 * Es, cudaMemcpyDeviceToDevice, s));
  }
#undef SEA_TRY_
  return SEA_OK;
}

extern "C" int sea_spatial_encode_tc(const sea_spatial_desc* d, const void* cache, float* x, float* z, int B,
                                     int latent_layout, float pad_idx, int fix_pad, sea_stream_t stream) {
  if (!x || !z || B <= 0) return SEA_ERR_INVALID;
  SpatialTC a{};
  int rc = fill_tc(d, cache, a);
  if (rc) return rc;
  for (int g = 0; g < a.n_groups; ++g)
    if (!d->enc_w1[g] || !d->enc_w2[g] || !a.enc_b2[g]) return SEA_ERR_INVALID;
  if (!a.ln_w || !a.ln_b || !a.pe) return SEA_ERR_INVALID;
  int gmax = 0;
  for (int g = 0; g < a.n_groups; ++g) gmax = a.g_count[g] > gmax ? a.g_count[g] : gmax;
  const EncPlan pl = enc_plan(gmax * a.Cp, a.Hs, a.n_groups * a.D);
  if (pl.total > 227 * 1024) return SEA_ERR_UNSUPPORTED;
  SEA_CUDA_OK(cudaFuncSetAttribute(spatial_encode_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(pl.total)));
  SEA_CUDA_OK(cudaFuncSetAttribute(spatial_encode_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  SEA_LAUNCH(spatial_encode_tc_kernel, B, kThreads, pl.total, reinterpret_cast<cudaStream_t>(stream), a, x, z, latent_layout,
             pad_idx, fix_pad);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int sea_spatial_decode_tc(const sea_spatial_desc* d, const void* cache, const float* z, float* out, int B,
                                     int latent_layout, sea_stream_t stream) {
  if (!z || !out || B <= 0) return SEA_ERR_INVALID;
  SpatialTC a{};
  int rc = fill_tc(d, cache, a);
  if (rc) return rc;
  for (int g = 0; g < a.n_groups; ++g)
    if (!d->dec_w1[g] || !d->dec_w2[g] || !a.dec_b2[g]) return SEA_ERR_INVALID;
  const int Es = a.n_groups * a.D;
  const size_t smem = sizeof(bf16) * (((P * pitch_of(Es) + 7) & ~7) + static_cast<size_t>(P) * pitch_of(a.Hs)) + 16;
  if (smem > 227 * 1024) return SEA_ERR_UNSUPPORTED;
  SEA_CUDA_OK(cudaFuncSetAttribute(spatial_decode_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  SEA_CUDA_OK(cudaFuncSetAttribute(spatial_decode_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  SEA_LAUNCH(spatial_decode_tc_kernel, B, kThreads, smem, reinterpret_cast<cudaStream_t>(stream), a, z, out, latent_layout);
  return static_cast<int>(cudaGetLastError());
}

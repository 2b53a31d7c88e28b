// K3 (CUDA-core variant) — backward of the fused causal attention.
// Flash-style recomputation: P = exp(S*scale - lse) is rebuilt from Q, K and the saved
// log-sum-exp; nothing of size T x T is stored.  Three kernels:
//   prep   delta[b,h,t] = sum_d dO * O
//   dq     one CTA per 16 queries, keys stream through smem (mirrors the forward kernel)
//   dkdv   one CTA per 16 keys, queries stream through smem
// The RoPE of the forward (fused into the projection GEMM epilogue) is undone on dQ / dK here:
// rotating the gradient back by -theta is the transpose of the forward rotation.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/sea_b200.h"
#include "internal.h"
#include "ptx.cuh"

namespace sea {
namespace {

constexpr int kWarps = 4;
constexpr int kPer = 4;                 // queries (dq) or keys (dkdv) per warp
constexpr int kPerCta = kWarps * kPer;  // 16
constexpr int kTile = 32;

template <typename T>
__device__ __forceinline__ float ldf(const T* p);
template <>
__device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

struct BwdDev {
  const void *q, *k, *v, *o, *d_o;
  long long ldq, ldk, ldv, ldo, lddo;
  const float* lse;
  float* delta;
  void *dq, *dk, *dv;
  long long lddq, lddk, lddv;
  int B, T, n_heads, hd, src_len;
  float scale;
  const float* rope;
  int rope_ld;
};

template <typename T>
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const BwdDev a) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int total = a.B * a.T * a.n_heads;
  if (warp >= total) return;
  const int h = warp % a.n_heads;
  const long long row = warp / a.n_heads;  // b*T + t
  const T* o = static_cast<const T*>(a.o) + row * a.ldo + h * a.hd;
  const T* d_o = static_cast<const T*>(a.d_o) + row * a.lddo + h * a.hd;
  float s = 0.f;
  for (int d = lane; d < a.hd; d += 32) s += ldf(o + d) * ldf(d_o + d);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) {
    const int b = static_cast<int>(row / a.T), t = static_cast<int>(row % a.T);
    a.delta[(static_cast<long long>(b) * a.n_heads + h) * a.T + t] = s;
  }
}

// value of column d = lane + 32c, rotated back:  d even: y*c + partner*s ; d odd: y*c - partner*s
__device__ __forceinline__ float unrope(float y, int d, int t, int rope_ld, const float* rope, int lane) {
  const float partner = __shfl_xor_sync(0xffffffffu, y, 1);
  if (rope == nullptr) return y;
  const float2 cs = *reinterpret_cast<const float2*>(rope + (static_cast<long long>(d >> 1) * rope_ld + t) * 2);
  return (lane & 1) ? (y * cs.x - partner * cs.y) : (y * cs.x + partner * cs.y);
}

template <typename T>
__global__ void __launch_bounds__(kWarps * 32) attn_bwd_dq_kernel(const BwdDev a) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  extern __shared__ float smem[];
  const int hd = a.hd, dpl = hd >> 5;
  float* Ks = smem;                          // [32][hd+1]
  float* Vs = Ks + kTile * (hd + 1);         // [32][hd+1]
  float* Qt = Vs + kTile * (hd + 1);         // [warps][hd][4]
  float* Gt = Qt + kWarps * hd * 4;          // [warps][hd][4]   dO^T
  float* Ss = Gt + kWarps * hd * 4;          // [warps][32][4]   dS
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y;
  const int q0 = blockIdx.x * kPerCta, qw = q0 + warp * kPer;
  const long long row0 = static_cast<long long>(b) * a.T;
  const T* Q = static_cast<const T*>(a.q);
  const T* K = static_cast<const T*>(a.k);
  const T* V = static_cast<const T*>(a.v);
  const T* G = static_cast<const T*>(a.d_o);
  float* myQ = Qt + warp * hd * 4;
  float* myG = Gt + warp * hd * 4;
  float* myS = Ss + warp * kTile * 4;
  for (int idx = lane; idx < hd * kPer; idx += 32) {
    const int i = idx / hd, d = idx - i * hd, q = qw + i;
    myQ[d * 4 + i] = (q < a.T) ? ldf(Q + (row0 + q) * a.ldq + h * hd + d) : 0.f;
    myG[d * 4 + i] = (q < a.T) ? ldf(G + (row0 + q) * a.lddo + h * hd + d) : 0.f;
  }
  float lse[kPer], dl[kPer], acc[kPer][8];
#pragma unroll
  for (int i = 0; i < kPer; ++i) {
    const int q = min(qw + i, a.T - 1);
    const long long si = (static_cast<long long>(b) * a.n_heads + h) * a.T + q;
    lse[i] = a.lse[si];
    dl[i] = a.delta[si];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[i][c] = 0.f;
  }
  const int k_last = min(a.T - 1, min(a.T - 1, q0 + kPerCta - 1) + a.src_len);
  const int n_tiles = k_last / kTile + 1;
  for (int kt = 0; kt < n_tiles; ++kt) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < kTile * hd; idx += blockDim.x) {
      const int r = idx / hd, d = idx - r * hd, key = kt * kTile + r;
      float kv = 0.f, vv = 0.f;
      if (key < a.T) {
        kv = ldf(K + (row0 + key) * a.ldk + h * hd + d);
        vv = ldf(V + (row0 + key) * a.ldv + h * hd + d);
      }
      Ks[r * (hd + 1) + d] = kv;
      Vs[r * (hd + 1) + d] = vv;
    }
    __syncthreads();
    if (kt * kTile > min(a.T - 1, qw + kPer - 1 + a.src_len)) continue;
    float s[kPer] = {0.f, 0.f, 0.f, 0.f}, dp[kPer] = {0.f, 0.f, 0.f, 0.f};
    const float* krow = Ks + lane * (hd + 1);
    const float* vrow = Vs + lane * (hd + 1);
    for (int d = 0; d < hd; ++d) {
      const float kv = krow[d], vv = vrow[d];
      const float4 q4 = *reinterpret_cast<const float4*>(myQ + d * 4);
      const float4 g4 = *reinterpret_cast<const float4*>(myG + d * 4);
      s[0] = fmaf(q4.x, kv, s[0]); s[1] = fmaf(q4.y, kv, s[1]);
      s[2] = fmaf(q4.z, kv, s[2]); s[3] = fmaf(q4.w, kv, s[3]);
      dp[0] = fmaf(g4.x, vv, dp[0]); dp[1] = fmaf(g4.y, vv, dp[1]);
      dp[2] = fmaf(g4.z, vv, dp[2]); dp[3] = fmaf(g4.w, vv, dp[3]);
    }
    const int key = kt * kTile + lane;
    float ds[kPer];
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int q = qw + i;
      const bool ok = (key < a.T) && (key <= q + a.src_len) && (q < a.T);
      const float p = ok ? expf(s[i] * a.scale - lse[i]) : 0.f;
      ds[i] = p * (dp[i] - dl[i]) * a.scale;
    }
    *reinterpret_cast<float4*>(myS + lane * 4) = make_float4(ds[0], ds[1], ds[2], ds[3]);
    __syncwarp();
    for (int j = 0; j < kTile; ++j) {
      const float4 s4 = *reinterpret_cast<const float4*>(myS + j * 4);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c < dpl) {
          const float kv = Ks[j * (hd + 1) + lane + 32 * c];
          acc[0][c] = fmaf(s4.x, kv, acc[0][c]); acc[1][c] = fmaf(s4.y, kv, acc[1][c]);
          acc[2][c] = fmaf(s4.z, kv, acc[2][c]); acc[3][c] = fmaf(s4.w, kv, acc[3][c]);
        }
      }
    }
    __syncwarp();
  }
  T* DQ = static_cast<T*>(a.dq);
#pragma unroll
  for (int i = 0; i < kPer; ++i) {
    const int q = qw + i;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (c < dpl) {
        const int d = lane + 32 * c;
        const float v = unrope(acc[i][c], d, min(q, a.T - 1), a.rope_ld, a.rope, lane);
        if (q < a.T) stf(DQ + (row0 + q) * a.lddq + h * hd + d, v);
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kWarps * 32) attn_bwd_dkdv_kernel(const BwdDev a) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  extern __shared__ float smem[];
  const int hd = a.hd, dpl = hd >> 5;
  float* Qs = smem;                          // [32][hd+1]
  float* Gs = Qs + kTile * (hd + 1);         // [32][hd+1]   dO
  float* Kt = Gs + kTile * (hd + 1);         // [warps][hd][4]
  float* Vt = Kt + kWarps * hd * 4;          // [warps][hd][4]
  float* Ps = Vt + kWarps * hd * 4;          // [warps][32][4]
  float* Ss = Ps + kWarps * kTile * 4;       // [warps][32][4]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y;
  const int k0 = blockIdx.x * kPerCta, kw = k0 + warp * kPer;
  const long long row0 = static_cast<long long>(b) * a.T;
  const T* Q = static_cast<const T*>(a.q);
  const T* K = static_cast<const T*>(a.k);
  const T* V = static_cast<const T*>(a.v);
  const T* G = static_cast<const T*>(a.d_o);
  float* myK = Kt + warp * hd * 4;
  float* myV = Vt + warp * hd * 4;
  float* myP = Ps + warp * kTile * 4;
  float* myS = Ss + warp * kTile * 4;
  for (int idx = lane; idx < hd * kPer; idx += 32) {
    const int i = idx / hd, d = idx - i * hd, key = kw + i;
    myK[d * 4 + i] = (key < a.T) ? ldf(K + (row0 + key) * a.ldk + h * hd + d) : 0.f;
    myV[d * 4 + i] = (key < a.T) ? ldf(V + (row0 + key) * a.ldv + h * hd + d) : 0.f;
  }
  float dk[kPer][8], dv[kPer][8];
#pragma unroll
  for (int i = 0; i < kPer; ++i)
#pragma unroll
    for (int c = 0; c < 8; ++c) { dk[i][c] = 0.f; dv[i][c] = 0.f; }
  const int q_first = max(0, k0 - a.src_len);
  for (int qt = q_first / kTile; qt * kTile < a.T; ++qt) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < kTile * hd; idx += blockDim.x) {
      const int r = idx / hd, d = idx - r * hd, q = qt * kTile + r;
      float qv = 0.f, gv = 0.f;
      if (q < a.T) {
        qv = ldf(Q + (row0 + q) * a.ldq + h * hd + d);
        gv = ldf(G + (row0 + q) * a.lddo + h * hd + d);
      }
      Qs[r * (hd + 1) + d] = qv;
      Gs[r * (hd + 1) + d] = gv;
    }
    __syncthreads();
    const int q = qt * kTile + lane;
    if (qt * kTile + kTile - 1 + a.src_len < kw) continue;  // no query of this tile sees this warp's keys
    float s[kPer] = {0.f, 0.f, 0.f, 0.f}, dp[kPer] = {0.f, 0.f, 0.f, 0.f};
    const float* qrow = Qs + lane * (hd + 1);
    const float* grow = Gs + lane * (hd + 1);
    for (int d = 0; d < hd; ++d) {
      const float qv = qrow[d], gv = grow[d];
      const float4 k4 = *reinterpret_cast<const float4*>(myK + d * 4);
      const float4 v4 = *reinterpret_cast<const float4*>(myV + d * 4);
      s[0] = fmaf(qv, k4.x, s[0]); s[1] = fmaf(qv, k4.y, s[1]);
      s[2] = fmaf(qv, k4.z, s[2]); s[3] = fmaf(qv, k4.w, s[3]);
      dp[0] = fmaf(gv, v4.x, dp[0]); dp[1] = fmaf(gv, v4.y, dp[1]);
      dp[2] = fmaf(gv, v4.z, dp[2]); dp[3] = fmaf(gv, v4.w, dp[3]);
    }
    float lse = 0.f, dl = 0.f;
    if (q < a.T) {
      const long long si = (static_cast<long long>(b) * a.n_heads + h) * a.T + q;
      lse = a.lse[si];
      dl = a.delta[si];
    }
    float p[kPer], ds[kPer];
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int key = kw + i;
      const bool ok = (q < a.T) && (key < a.T) && (key <= q + a.src_len);
      p[i] = ok ? expf(s[i] * a.scale - lse) : 0.f;
      ds[i] = p[i] * (dp[i] - dl) * a.scale;
    }
    *reinterpret_cast<float4*>(myP + lane * 4) = make_float4(p[0], p[1], p[2], p[3]);
    *reinterpret_cast<float4*>(myS + lane * 4) = make_float4(ds[0], ds[1], ds[2], ds[3]);
    __syncwarp();
    for (int j = 0; j < kTile; ++j) {
      const float4 p4 = *reinterpret_cast<const float4*>(myP + j * 4);
      const float4 s4 = *reinterpret_cast<const float4*>(myS + j * 4);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c < dpl) {
          const float gv = Gs[j * (hd + 1) + lane + 32 * c];
          const float qv = Qs[j * (hd + 1) + lane + 32 * c];
          dv[0][c] = fmaf(p4.x, gv, dv[0][c]); dv[1][c] = fmaf(p4.y, gv, dv[1][c]);
          dv[2][c] = fmaf(p4.z, gv, dv[2][c]); dv[3][c] = fmaf(p4.w, gv, dv[3][c]);
          dk[0][c] = fmaf(s4.x, qv, dk[0][c]); dk[1][c] = fmaf(s4.y, qv, dk[1][c]);
          dk[2][c] = fmaf(s4.z, qv, dk[2][c]); dk[3][c] = fmaf(s4.w, qv, dk[3][c]);
        }
      }
    }
    __syncwarp();
  }
  T* DK = static_cast<T*>(a.dk);
  T* DV = static_cast<T*>(a.dv);
#pragma unroll
  for (int i = 0; i < kPer; ++i) {
    const int key = kw + i;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (c < dpl) {
        const int d = lane + 32 * c;
        const float kv = unrope(dk[i][c], d, min(key, a.T - 1), a.rope_ld, a.rope, lane);
        if (key < a.T) {
          stf(DK + (row0 + key) * a.lddk + h * hd + d, kv);
          stf(DV + (row0 + key) * a.lddv + h * hd + d, dv[i][c]);
        }
      }
    }
  }
}

template <typename T>
int launch_bwd(const BwdDev& d, cudaStream_t s) {
  const int hd = d.hd;
  const size_t smem_dq = sizeof(float) * (2 * kTile * (hd + 1) + 2 * kWarps * hd * 4 + kWarps * kTile * 4);
  const size_t smem_kv = sizeof(float) * (2 * kTile * (hd + 1) + 2 * kWarps * hd * 4 + 2 * kWarps * kTile * 4);
  static bool attr_set[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 16 && !attr_set[dev]) {
    SEA_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dq_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    SEA_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dkdv_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    attr_set[dev] = true;
  }
  const int total_warps = d.B * d.T * d.n_heads;
  SEA_LAUNCH((attn_bwd_prep_kernel<T>), (total_warps + 7) / 8, 256, 0, s, d);
  dim3 grid((d.T + kPerCta - 1) / kPerCta, d.n_heads, d.B);
  SEA_LAUNCH((attn_bwd_dq_kernel<T>), grid, kWarps * 32, smem_dq, s, d);
  SEA_LAUNCH((attn_bwd_dkdv_kernel<T>), grid, kWarps * 32, smem_kv, s, d);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace
extern int g_force_simt;                                                  // attention.cu
bool attention_bwd_tc_supported(const sea_attn_bwd_args* a);              // attention_bwd_tc.cu
int attention_bwd_tc(const sea_attn_bwd_args* a, cudaStream_t s);
}  // namespace sea

extern "C" int sea_attention_bwd(const sea_attn_bwd_args* a, sea_stream_t stream) {
  using namespace sea;
  if (!a || !a->q || !a->k || !a->v || !a->o || !a->d_o || !a->lse || !a->delta || !a->dq || !a->dk || !a->dv)
    return SEA_ERR_INVALID;
  if (a->B <= 0 || a->T <= 0 || a->n_heads <= 0) return SEA_ERR_INVALID;
  if (a->head_dim % 32 || a->head_dim > 256 || a->head_dim <= 0) return SEA_ERR_UNSUPPORTED;
  BwdDev d;
  d.q = a->q; d.k = a->k; d.v = a->v; d.o = a->o; d.d_o = a->d_o;
  d.ldq = a->ldq; d.ldk = a->ldk; d.ldv = a->ldv; d.ldo = a->ldo; d.lddo = a->lddo;
  d.lse = a->lse; d.delta = a->delta;
  d.dq = a->dq; d.dk = a->dk; d.dv = a->dv; d.lddq = a->lddq; d.lddk = a->lddk; d.lddv = a->lddv;
  d.B = a->B; d.T = a->T; d.n_heads = a->n_heads; d.hd = a->head_dim; d.src_len = a->src_len;
  d.scale = a->scale; d.rope = a->rope_table; d.rope_ld = a->rope_ld;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (!g_force_simt && attention_bwd_tc_supported(a)) return attention_bwd_tc(a, s);
  if (a->dropout_p != 0.f) return SEA_ERR_UNSUPPORTED;   // probability dropout lives in the tensor-core kernels only
  if (a->prec == SEA_PREC_FP32) return launch_bwd<float>(d, s);
  if (a->prec == SEA_PREC_BF16) return launch_bwd<__nv_bfloat16>(d, s);
  return SEA_ERR_INVALID;
}

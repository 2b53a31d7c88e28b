// K2 (CUDA-core variant) — fused causal attention forward with online softmax in registers.
// Serves the fp32-parity mode (all math fp32) and head dims the tcgen05 kernel does not take.
//
// CTA = 4 warps = 16 queries of one (batch, head); keys/values stream through shared memory in
// tiles of 32; each lane owns one key of the tile for Q·K^T (4 queries register-blocked) and
// head_dim/32 output columns for P·V.  The [T,T] score matrix never exists in memory
// (the reference materialises it: models/base_blocks.py:191-194).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/sea_b200.h"
#include "internal.h"
#include "ptx.cuh"

namespace sea {
namespace {

constexpr int kQPerWarp = 4;
constexpr int kWarps = 4;
constexpr int kQPerCta = kQPerWarp * kWarps;
constexpr int kKeyTile = 32;

template <typename T>
__device__ __forceinline__ float ldf(const T* p);
template <>
__device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename T>
__global__ void __launch_bounds__(kWarps * 32)
attn_fwd_simt_kernel(const T* __restrict__ Q, const T* __restrict__ K, const T* __restrict__ V,
                     long long ldq, long long ldk, long long ldv, T* __restrict__ O, long long ldo,
                     float* __restrict__ lse, int Tlen, int n_heads, int hd, int src_len, float scale) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  extern __shared__ float smem[];
  const int dpl = hd >> 5;
  float* Ks = smem;                         // [32][hd + 1]
  float* Vs = Ks + kKeyTile * (hd + 1);     // [32][hd]
  float* Qt = Vs + kKeyTile * hd;           // [warps][hd][4]
  float* Ps = Qt + kWarps * hd * 4;         // [warps][32][4]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y;
  const int q0 = blockIdx.x * kQPerCta;
  const int qw = q0 + warp * kQPerWarp;
  const long long row0 = static_cast<long long>(b) * Tlen;
  float* myQ = Qt + warp * hd * 4;
  float* myP = Ps + warp * kKeyTile * 4;

  for (int idx = lane; idx < hd * kQPerWarp; idx += 32) {
    const int i = idx / hd, d = idx - i * hd;
    const int q = qw + i;
    myQ[d * 4 + i] = (q < Tlen) ? ldf(Q + (row0 + q) * ldq + h * hd + d) : 0.f;
  }
  float m_run[kQPerWarp], l_run[kQPerWarp], o[kQPerWarp][8];
#pragma unroll
  for (int i = 0; i < kQPerWarp; ++i) {
    m_run[i] = -INFINITY;
    l_run[i] = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) o[i][c] = 0.f;
  }
  const int k_last = min(Tlen - 1, min(Tlen - 1, q0 + kQPerCta - 1) + src_len);
  const int n_tiles = k_last / kKeyTile + 1;
  for (int kt = 0; kt < n_tiles; ++kt) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < kKeyTile * hd; idx += blockDim.x) {
      const int r = idx / hd, d = idx - r * hd;
      const int key = kt * kKeyTile + r;
      float kv = 0.f, vv = 0.f;
      if (key < Tlen) {
        kv = ldf(K + (row0 + key) * ldk + h * hd + d);
        vv = ldf(V + (row0 + key) * ldv + h * hd + d);
      }
      Ks[r * (hd + 1) + d] = kv;
      Vs[r * hd + d] = vv;
    }
    __syncthreads();
    if (kt * kKeyTile > min(Tlen - 1, qw + kQPerWarp - 1 + src_len)) continue;  // warp-uniform
    float s[kQPerWarp] = {0.f, 0.f, 0.f, 0.f};
    const float* krow = Ks + lane * (hd + 1);
    for (int d = 0; d < hd; ++d) {
      const float kv = krow[d];
      const float4 q4 = *reinterpret_cast<const float4*>(myQ + d * 4);
      s[0] = fmaf(q4.x, kv, s[0]);
      s[1] = fmaf(q4.y, kv, s[1]);
      s[2] = fmaf(q4.z, kv, s[2]);
      s[3] = fmaf(q4.w, kv, s[3]);
    }
    const int key = kt * kKeyTile + lane;
    float p[kQPerWarp];
#pragma unroll
    for (int i = 0; i < kQPerWarp; ++i) {
      const int q = qw + i;
      const bool ok = (key < Tlen) && (key <= q + src_len) && (q < Tlen);
      const float sv = ok ? s[i] * scale : -INFINITY;
      float mt = sv;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, off));
      const float m_new = fmaxf(m_run[i], mt);
      float alpha = 1.f, pv = 0.f;
      if (m_new != -INFINITY) {
        alpha = expf(m_run[i] - m_new);
        pv = ok ? expf(sv - m_new) : 0.f;
      }
      float ps = pv;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
      l_run[i] = l_run[i] * alpha + ps;
      m_run[i] = m_new;
#pragma unroll
      for (int c = 0; c < 8; ++c) o[i][c] *= alpha;
      p[i] = pv;
    }
    *reinterpret_cast<float4*>(myP + lane * 4) = make_float4(p[0], p[1], p[2], p[3]);
    __syncwarp();
    for (int j = 0; j < kKeyTile; ++j) {
      const float4 p4 = *reinterpret_cast<const float4*>(myP + j * 4);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c < dpl) {
          const float vv = Vs[j * hd + lane + 32 * c];
          o[0][c] = fmaf(p4.x, vv, o[0][c]);
          o[1][c] = fmaf(p4.y, vv, o[1][c]);
          o[2][c] = fmaf(p4.z, vv, o[2][c]);
          o[3][c] = fmaf(p4.w, vv, o[3][c]);
        }
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int i = 0; i < kQPerWarp; ++i) {
    const int q = qw + i;
    if (q < Tlen) {
      const float inv = 1.f / l_run[i];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c < dpl) stf(O + (row0 + q) * ldo + h * hd + lane + 32 * c, o[i][c] * inv);
      }
      if (lse && lane == 0)
        lse[(static_cast<long long>(b) * n_heads + h) * Tlen + q] = m_run[i] + logf(l_run[i]);
    }
  }
}

template <typename T>
int launch_simt(const sea_attn_args* a, cudaStream_t s) {
  const int hd = a->head_dim;
  const size_t smem = sizeof(float) * (kKeyTile * (hd + 1) + kKeyTile * hd + kWarps * hd * 4 + kWarps * kKeyTile * 4);
  static bool attr_set[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 16 && !attr_set[dev]) {
    SEA_CUDA_OK(cudaFuncSetAttribute(attn_fwd_simt_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr_set[dev] = true;
  }
  dim3 grid((a->T + kQPerCta - 1) / kQPerCta, a->n_heads, a->B);
  SEA_LAUNCH((attn_fwd_simt_kernel<T>), grid, kWarps * 32, smem, s, static_cast<const T*>(a->q), static_cast<const T*>(a->k), static_cast<const T*>(a->v), a->ldq, a->ldk, a->ldv, static_cast<T*>(a->o), a->ldo, a->lse, a->T, a->n_heads, hd, a->src_len, a->scale);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace

int attention_fwd_simt(const sea_attn_args* a, cudaStream_t s) {
  if (a->head_dim % 32 || a->head_dim > 256) return SEA_ERR_UNSUPPORTED;
  if (a->prec == SEA_PREC_FP32) return launch_simt<float>(a, s);
  return launch_simt<__nv_bfloat16>(a, s);
}

}  // namespace sea

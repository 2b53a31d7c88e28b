// K2'' — decode attention for the KV-cached incremental step.
// One NEW query per (trajectory, head) at absolute position `pos` against the cached keys / values
// 0 .. pos (causal, models/base_blocks.py:191-197 / :283-289 restricted to the last row).  The work
// per (b, h) is (pos+1) x head_dim multiply-adds twice — far below a tensor-core tile — so this is a
// CUDA-core kernel: one CTA of 4 warps per (b, h),
//   phase 1  threads stride over the keys, each computes whole q.k dot products (q broadcast from
//            shared memory, 16-byte loads of its key row) into a score row in shared memory;
//   phase 2  block max / sum of exp2 over the score row;
//   phase 3  warps stride over the value rows, lanes own head_dim/32 consecutive output columns
//            (coalesced row reads, several rows in flight), partial outputs combined in shared memory.
// Up to SEA_MAX_STREAMS same-shape problems per launch (blockIdx.z).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/sea_b200.h"
#include "internal.h"
#include "ptx.cuh"

namespace sea {
namespace {

constexpr int kWarpsPerCta = 4;

struct DecItem { const void *q, *k, *v; void* o; };
struct DecParams {
  DecItem it[SEA_MAX_STREAMS];
  long long ldq, ldo;          // row pitch between trajectories of q / o (one row per trajectory)
  long long ldk, ldv;          // row pitch between cached positions
  long long bsk, bsv;          // pitch between trajectories of the caches
  int B, n_keys, n_heads, hd;
  float scale_log2;
};

template <typename T> struct Vec8;
template <> struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 raw = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
  }
};
template <> struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&f)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
};
__device__ __forceinline__ float tof(float v) { return v; }
__device__ __forceinline__ float tof(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ void sto(float* p, float v) { *p = v; }
__device__ __forceinline__ void sto(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename T>
__global__ void __launch_bounds__(kWarpsPerCta * 32) attn_decode_kernel(const DecParams p) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  extern __shared__ float dsm[];
  __shared__ float wred[2 * kWarpsPerCta];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / p.n_heads, h = blockIdx.x - b * p.n_heads;
  const DecItem& it = p.it[blockIdx.z];
  const int hd = p.hd, n = p.n_keys, cpl = hd >> 5;  // columns per lane in phase 3
  float* qs = dsm;                        // [hd]
  float* sc = qs + hd;                    // [n rounded to 32]
  float* part = sc + ((n + 31) & ~31);    // [warps][hd] partial outputs
  const T* q = static_cast<const T*>(it.q) + static_cast<long long>(b) * p.ldq + h * hd;
  const T* K = static_cast<const T*>(it.k) + static_cast<long long>(b) * p.bsk + h * hd;
  const T* V = static_cast<const T*>(it.v) + static_cast<long long>(b) * p.bsv + h * hd;
  for (int d = tid; d < hd; d += blockDim.x) qs[d] = tof(q[d]);
  __syncthreads();
  // phase 1: one key per thread (stride 128): whole q.k dot products, log2 domain
  float mx = -INFINITY;
  for (int k = tid; k < n; k += blockDim.x) {
    const T* kr = K + static_cast<long long>(k) * p.ldk;
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 4
    for (int d = 0; d < hd; d += 16) {   // 16-byte loads, several in flight
      float f[8], g[8];
      Vec8<T>::load(kr + d, f);
      Vec8<T>::load(kr + d + 8, g);
#pragma unroll
      for (int e = 0; e < 8; ++e) { acc0 = fmaf(f[e], qs[d + e], acc0); acc1 = fmaf(g[e], qs[d + 8 + e], acc1); }
    }
    const float acc = (acc0 + acc1) * p.scale_log2;
    sc[k] = acc;
    mx = fmaxf(mx, acc);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  if (lane == 0) wred[warp] = mx;
  __syncthreads();
  mx = wred[0];
#pragma unroll
  for (int w = 1; w < kWarpsPerCta; ++w) mx = fmaxf(mx, wred[w]);
  // phase 2: probabilities
  float sum = 0.f;
  for (int k = tid; k < n; k += blockDim.x) {
    const float e = exp2f(sc[k] - mx);
    sc[k] = e;
    sum += e;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  if (lane == 0) wred[kWarpsPerCta + warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < kWarpsPerCta; ++w) sum += wred[kWarpsPerCta + w];
  // phase 3: warp w sweeps value rows w, w+4, ...; lane owns `cpl` consecutive columns
  float o[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) o[c] = 0.f;
#pragma unroll 4
  for (int k = warp; k < n; k += kWarpsPerCta) {
    const float pk = sc[k];
    const T* vr = V + static_cast<long long>(k) * p.ldv + lane * cpl;
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c < cpl) o[c] = fmaf(pk, tof(vr[c]), o[c]);
  }
#pragma unroll
  for (int c = 0; c < 8; ++c)
    if (c < cpl) part[warp * hd + lane * cpl + c] = o[c];
  __syncthreads();
  const float inv = 1.f / sum;
  T* orow = static_cast<T*>(it.o) + static_cast<long long>(b) * p.ldo + h * hd;
  for (int d = tid; d < hd; d += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kWarpsPerCta; ++w) v += part[w * hd + d];
    sto(orow + d, v * inv);
  }
}

}  // namespace
}  // namespace sea

extern "C" int sea_attention_decode_group(int n, const sea_attn_decode_args* a, sea_stream_t stream) {
  using namespace sea;
  if (!a || n < 1 || n > SEA_MAX_STREAMS) return SEA_ERR_INVALID;
  DecParams p;
  for (int i = 0; i < n; ++i) {
    const sea_attn_decode_args& x = a[i];
    if (!x.q || !x.k || !x.v || !x.o) return SEA_ERR_INVALID;
    if (x.B != a->B || x.n_keys != a->n_keys || x.n_heads != a->n_heads || x.head_dim != a->head_dim ||
        x.prec != a->prec || x.scale != a->scale || x.ldq != a->ldq || x.ldo != a->ldo || x.ldk != a->ldk ||
        x.ldv != a->ldv || x.k_batch_stride != a->k_batch_stride || x.v_batch_stride != a->v_batch_stride)
      return SEA_ERR_INVALID;
    p.it[i] = DecItem{x.q, x.k, x.v, x.o};
  }
  if (a->B <= 0 || a->n_keys <= 0 || a->n_heads <= 0) return SEA_ERR_INVALID;
  if (a->head_dim <= 0 || (a->head_dim % 32) || a->head_dim > 256) return SEA_ERR_UNSUPPORTED;
  const int esz = a->prec == SEA_PREC_FP32 ? 4 : 2;
  if (a->prec != SEA_PREC_FP32 && a->prec != SEA_PREC_BF16) return SEA_ERR_INVALID;
  // 16-byte loads of key rows
  if (((a->ldk * esz) % 16) || ((a->k_batch_stride * esz) % 16)) return SEA_ERR_INVALID;
  for (int i = 0; i < n; ++i)
    if (reinterpret_cast<uintptr_t>(a[i].k) & 15) return SEA_ERR_INVALID;
  p.ldq = a->ldq; p.ldo = a->ldo; p.ldk = a->ldk; p.ldv = a->ldv;
  p.bsk = a->k_batch_stride; p.bsv = a->v_batch_stride;
  p.B = a->B; p.n_keys = a->n_keys; p.n_heads = a->n_heads; p.hd = a->head_dim;
  p.scale_log2 = a->scale * 1.44269504088896340736f;
  const size_t smem = sizeof(float) * (a->head_dim + ((a->n_keys + 31) & ~31) + kWarpsPerCta * a->head_dim);
  if (smem > 200 * 1024) return SEA_ERR_UNSUPPORTED;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  static bool attr_set[16][2] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  const int ti = a->prec == SEA_PREC_FP32 ? 1 : 0;
  if (dev < 16 && !attr_set[dev][ti]) {
    if (ti) SEA_CUDA_OK(cudaFuncSetAttribute(attn_decode_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    else SEA_CUDA_OK(cudaFuncSetAttribute(attn_decode_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set[dev][ti] = true;
  }
  const dim3 grid(a->B * a->n_heads, 1, n);
  if (ti) SEA_LAUNCH(attn_decode_kernel<float>, grid, kWarpsPerCta * 32, smem, s, p);
  else SEA_LAUNCH(attn_decode_kernel<__nv_bfloat16>, grid, kWarpsPerCta * 32, smem, s, p);
  return static_cast<int>(cudaGetLastError());
}

// K7 / K8 — fused ViT-mesh patch encoder and decoder (models/encoder_decoder.py:75-146).
//
// One CTA per snapshot.  The whole per-snapshot state (64 patches x Es latent, Es = G*D = 32 / 64)
// lives in shared memory from the patch gather to the final LayerNorm: the 12 pre-LN encoder
// blocks (non-causal MHA with head_dim 4 / 8, MLP x4 with inner LayerNorm + GELU) never touch HBM
// except to stream their weights (1.5 / 3.7 MB per model, L2-resident, shared by all CTAs).
// HBM traffic per snapshot is therefore the algorithmic minimum: the field patches in, the latent
// out (4*P*F*C + 4*P*G*D bytes).  Widths this small are not a tensor-core problem
// (SURVEY hard-part 11): all contractions are register-blocked fp32 FFMA (8 rows x 1 column per
// thread, float4 along K, activations broadcast from smem), results match the fp32 reference to ~1e-6.
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/sea_b200.h"
#include "internal.h"
#include "ptx.cuh"

namespace sea {
namespace {

constexpr int P = 64;         // patches per snapshot (m = n = 9 -> 64), fixed by the reference configs
constexpr int kThreads = 256;
constexpr int kChunk = 96;    // hidden columns of the patch MLPs processed per pass

struct LayerDev {
  const float *ln1_w, *q_w, *q_b, *k_w, *k_b, *v_w, *v_b, *proj_w, *ln2_w;
  const float *mlp0_w, *mlp0_b, *mlp_ln_w, *mlp_ln_b, *mlp3_w, *mlp3_b;
};
struct SpatialDev {
  int n_groups, n_fields, C, Hs, D, n_heads, num_layers;
  int g_first[4], g_count[4];
  const float *enc_w1[4], *enc_w2[4], *enc_b2[4];
  const float *dec_w1[4], *dec_w2[4], *dec_b2[4];
  const float *ln_w, *ln_b, *pe;
  LayerDev layers[16];
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Y[64][ldy] (=|+=) act(X[64][ldx](:, 0:K) * W[N][ldw]^T + bias).  X, Y in shared memory, W global.
// Work item = (column n, group of 8 rows); consecutive threads take consecutive n so the X reads of
// a warp are broadcasts and the Y writes are conflict-free.  K % 4 == 0, 16-byte aligned rows.
template <bool ACC>
__device__ __forceinline__ void linear64(const float* __restrict__ X, int ldx, int K,
                                         const float* __restrict__ W, int ldw,
                                         const float* __restrict__ bias, int N, float* __restrict__ Y,
                                         int ldy, bool gelu) {
  for (int item = threadIdx.x; item < N * 8; item += kThreads) {
    const int rg = item / N, n = item - rg * N;
    float acc[8];
    const float b = bias ? __ldg(bias + n) : 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = ACC ? Y[(rg * 8 + r) * ldy + n] + b : b;
    const float* wrow = W + static_cast<long long>(n) * ldw;
    const float* xrow = X + rg * 8 * ldx;
    for (int k = 0; k < K; k += 4) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(wrow + k));
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float4 x = *reinterpret_cast<const float4*>(xrow + r * ldx + k);
        acc[r] = fmaf(x.x, w.x, acc[r]);
        acc[r] = fmaf(x.y, w.y, acc[r]);
        acc[r] = fmaf(x.z, w.z, acc[r]);
        acc[r] = fmaf(x.w, w.w, acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) Y[(rg * 8 + r) * ldy + n] = gelu ? ptx::gelu_erf(acc[r]) : acc[r];
  }
}

// Same contraction with scalar loads: rows of X / W that are not 16-byte aligned (n_inp not a multiple of 4 — the
// reference sets n_inp to the cell count of the fullest patch, utils/data_processors.py:61-88, an arbitrary integer).
__device__ __forceinline__ void linear64_scalar(const float* __restrict__ X, int ldx, int K, const float* __restrict__ W,
                                                int ldw, int N, float* __restrict__ Y, int ldy, bool gelu) {
  for (int item = threadIdx.x; item < N * 8; item += kThreads) {
    const int rg = item / N, n = item - rg * N;
    float acc[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = 0.f;
    const float* wrow = W + static_cast<long long>(n) * ldw;
    const float* xrow = X + rg * 8 * ldx;
    for (int k = 0; k < K; ++k) {
      const float w = __ldg(wrow + k);
#pragma unroll
      for (int r = 0; r < 8; ++r) acc[r] = fmaf(xrow[r * ldx + k], w, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) Y[(rg * 8 + r) * ldy + n] = gelu ? ptx::gelu_erf(acc[r]) : acc[r];
  }
}

// Row LayerNorm over d columns for the 64 rows (warp per row): Y = (X-mean)*rstd*w (+b), opt. GELU.
__device__ __forceinline__ void layernorm64(const float* __restrict__ X, int ldx, int d,
                                            const float* __restrict__ w, const float* __restrict__ b,
                                            float* __restrict__ Y, int ldy, bool gelu) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int row = warp; row < P; row += kThreads / 32) {
    const float* xr = X + row * ldx;
    float s = 0.f;
    for (int c = lane; c < d; c += 32) s += xr[c];
    const float mean = warp_sum(s) / d;
    float q = 0.f;
    for (int c = lane; c < d; c += 32) q += (xr[c] - mean) * (xr[c] - mean);
    const float rstd = rsqrtf(warp_sum(q) / d + 1e-5f);
    for (int c = lane; c < d; c += 32) {
      float y = (xr[c] - mean) * rstd * __ldg(w + c);
      if (b) y += __ldg(b + c);
      Y[row * ldy + c] = gelu ? ptx::gelu_erf(y) : y;
    }
  }
}

// Non-causal multi-head attention over the 64 patches (models/base_blocks.py:105-121).
// QKV: smem [64][3*Es] (q | k | v); O: smem [64][Es].  One (head, query) pair per thread pass;
// a warp holds 32 queries of ONE head, so every K / V read is a broadcast.
template <int HD>
__device__ __forceinline__ void attention64(const float* __restrict__ QKV, int Es, int n_heads,
                                            float* __restrict__ O) {
  const int ld = 3 * Es;
  const float scale = rsqrtf(static_cast<float>(HD));
  for (int pair = threadIdx.x; pair < n_heads * P; pair += kThreads) {
    const int h = pair / P, i = pair - h * P;
    float q[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) q[d] = QKV[i * ld + h * HD + d] * scale;
    float s[P];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const float* kr = QKV + j * ld + Es + h * HD;
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) a = fmaf(q[d], kr[d], a);
      s[j] = a;
      m = fmaxf(m, a);
    }
    float l = 0.f, o[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = 0.f;
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const float p = expf(s[j] - m);
      l += p;
      const float* vr = QKV + j * ld + 2 * Es + h * HD;
#pragma unroll
      for (int d = 0; d < HD; ++d) o[d] = fmaf(p, vr[d], o[d]);
    }
    const float inv = 1.f / l;
#pragma unroll
    for (int d = 0; d < HD; ++d) O[i * Es + h * HD + d] = o[d] * inv;
  }
}

__device__ __forceinline__ void attention_dispatch(const float* QKV, int Es, int n_heads, float* O) {
  switch (Es / n_heads) {
    case 2: attention64<2>(QKV, Es, n_heads, O); break;
    case 4: attention64<4>(QKV, Es, n_heads, O); break;
    case 8: attention64<8>(QKV, Es, n_heads, O); break;
    default: attention64<16>(QKV, Es, n_heads, O); break;
  }
}

__device__ __forceinline__ long long latent_index(int b, int p, int g, int d, int G, int D, int layout) {
  // layout 0: [B, P, G, D] (module output, models/encoder_decoder.py:121)
  // layout 1: [B, G, P*D]  (transform_processed_data, utils/train_utils.py:315-337)
  return layout == 0 ? ((static_cast<long long>(b) * P + p) * G + g) * D + d
                     : ((static_cast<long long>(b) * G + g) * P + p) * D + d;
}

// ------------------------------------------------------------------------------------ encoder
__global__ void __launch_bounds__(kThreads) spatial_encode_kernel(const SpatialDev a, float* __restrict__ x,
                                                                  float* __restrict__ z, int layout,
                                                                  float pad_idx, int fix_pad, int staged) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  extern __shared__ __align__(16) float sm[];
  const int FC = a.n_fields * a.C, Es = a.n_groups * a.D;
  const int big_w = max(4 * Es, kChunk);
  // staged == 0 (snapshot too wide for shared memory): the patch MLP reads the snapshot in place
  float* Xin = sm;                  // [64][FC]      (only until the patch MLPs are done)
  float* Z = Xin + (staged ? P * FC : 0);   // [64][Es]  residual state
  float* Nn = Z + P * Es;           // [64][Es]      normed input / attention output
  float* BIG = Nn + P * Es;         // [64][big_w]   q|k|v, MLP hidden, patch-MLP hidden chunk
  const int b = blockIdx.x;
  float* xb = x + static_cast<long long>(b) * P * FC;
  // (1) gather the snapshot; generate_padding_mask (models/encoder_decoder.py:173-176) in place
  for (int i = threadIdx.x * 4; i < P * FC; i += kThreads * 4) {
    float4 v = *reinterpret_cast<const float4*>(xb + i);
    if (fix_pad) {
      const bool hit = v.x == pad_idx || v.y == pad_idx || v.z == pad_idx || v.w == pad_idx;
      if (hit) {
        v.x = v.x == pad_idx ? 0.f : v.x; v.y = v.y == pad_idx ? 0.f : v.y;
        v.z = v.z == pad_idx ? 0.f : v.z; v.w = v.w == pad_idx ? 0.f : v.w;
        *reinterpret_cast<float4*>(xb + i) = v;
      }
    }
    if (staged) *reinterpret_cast<float4*>(Xin + i) = v;
  }
  __syncthreads();
  if (!staged) Xin = xb;
  // (2) per-group patch MLP: Linear(C*g, Hs, no bias) -> GELU -> Linear(Hs, D) + b   (:108-111)
  for (int g = 0; g < a.n_groups; ++g) {
    const int Kin = a.g_count[g] * a.C;
    const float* xin = Xin + a.g_first[g] * a.C;
    for (int c0 = 0; c0 < a.Hs; c0 += kChunk) {
      const int ch = min(kChunk, a.Hs - c0);
      if ((a.C & 3) == 0) linear64<false>(xin, FC, Kin, a.enc_w1[g] + static_cast<long long>(c0) * Kin, Kin, nullptr, ch, BIG, big_w, true);
      else linear64_scalar(xin, FC, Kin, a.enc_w1[g] + static_cast<long long>(c0) * Kin, Kin, ch, BIG, big_w, true);
      __syncthreads();
      if (c0 == 0) linear64<false>(BIG, big_w, ch, a.enc_w2[g] + c0, a.Hs, a.enc_b2[g], a.D, Z + g * a.D, Es, false);
      else linear64<true>(BIG, big_w, ch, a.enc_w2[g] + c0, a.Hs, nullptr, a.D, Z + g * a.D, Es, false);
      __syncthreads();
    }
  }
  // (3) + sinusoidal positional encoding over the PATCH axis (:114, base_blocks.py:355-372)
  for (int i = threadIdx.x; i < P * Es; i += kThreads) Z[i] += __ldg(a.pe + i);
  __syncthreads();
  // (4) encoder blocks (base_blocks.py:123-138)
  for (int l = 0; l < a.num_layers; ++l) {
    const LayerDev& L = a.layers[l];
    layernorm64(Z, Es, Es, L.ln1_w, nullptr, Nn, Es, false);
    __syncthreads();
    linear64<false>(Nn, Es, Es, L.q_w, Es, L.q_b, Es, BIG, 3 * Es, false);
    linear64<false>(Nn, Es, Es, L.k_w, Es, L.k_b, Es, BIG + Es, 3 * Es, false);
    linear64<false>(Nn, Es, Es, L.v_w, Es, L.v_b, Es, BIG + 2 * Es, 3 * Es, false);
    __syncthreads();
    attention_dispatch(BIG, Es, a.n_heads, Nn);
    __syncthreads();
    linear64<true>(Nn, Es, Es, L.proj_w, Es, nullptr, Es, Z, Es, false);
    __syncthreads();
    layernorm64(Z, Es, Es, L.ln2_w, nullptr, Nn, Es, false);
    __syncthreads();
    linear64<false>(Nn, Es, Es, L.mlp0_w, Es, L.mlp0_b, 4 * Es, BIG, 4 * Es, false);
    __syncthreads();
    layernorm64(BIG, 4 * Es, 4 * Es, L.mlp_ln_w, L.mlp_ln_b, BIG, 4 * Es, true);
    __syncthreads();
    linear64<true>(BIG, 4 * Es, 4 * Es, L.mlp3_w, 4 * Es, L.mlp3_b, Es, Z, Es, false);
    __syncthreads();
  }
  // (5) final nn.LayerNorm(Es) and the latent store (optionally already in the temporal layout)
  layernorm64(Z, Es, Es, a.ln_w, a.ln_b, Nn, Es, false);
  __syncthreads();
  for (int i = threadIdx.x; i < P * Es; i += kThreads) {
    const int p = i / Es, c = i - p * Es, g = c / a.D, d = c - g * a.D;
    z[latent_index(b, p, g, d, a.n_groups, a.D, layout)] = Nn[i];
  }
}

// ------------------------------------------------------------------------------------ decoder
__global__ void __launch_bounds__(kThreads) spatial_decode_kernel(const SpatialDev a, const float* __restrict__ z,
                                                                  float* __restrict__ out, int layout, int staged) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  extern __shared__ __align__(16) float sm[];
  const int FC = a.n_fields * a.C, Es = a.n_groups * a.D;
  const int b = blockIdx.x;
  float* ob = out + static_cast<long long>(b) * P * FC;
  float* Zin = sm;                 // [64][Es]
  float* BIG = Zin + P * Es;       // [64][kChunk]
  float* Out = staged ? BIG + P * kChunk : ob;   // [64][FC]; accumulated in place when too wide for smem
  for (int i = threadIdx.x; i < P * Es; i += kThreads) {
    const int p = i / Es, c = i - p * Es, g = c / a.D, d = c - g * a.D;
    Zin[i] = z[latent_index(b, p, g, d, a.n_groups, a.D, layout)];
  }
  __syncthreads();
  // per group: Linear(D, Hs, no bias) -> GELU -> Linear(Hs, C*g) + b   (encoder_decoder.py:140-143)
  for (int g = 0; g < a.n_groups; ++g) {
    const int Nout = a.g_count[g] * a.C;
    float* og = Out + a.g_first[g] * a.C;
    for (int c0 = 0; c0 < a.Hs; c0 += kChunk) {
      const int ch = min(kChunk, a.Hs - c0);
      linear64<false>(Zin + g * a.D, Es, a.D, a.dec_w1[g] + static_cast<long long>(c0) * a.D, a.D, nullptr, ch, BIG, kChunk, true);
      __syncthreads();
      if (c0 == 0) linear64<false>(BIG, kChunk, ch, a.dec_w2[g] + c0, a.Hs, a.dec_b2[g], Nout, og, FC, false);
      else linear64<true>(BIG, kChunk, ch, a.dec_w2[g] + c0, a.Hs, nullptr, Nout, og, FC, false);
      __syncthreads();
    }
  }
  if (!staged) return;
  for (int i = threadIdx.x * 4; i < P * FC; i += kThreads * 4)
    *reinterpret_cast<float4*>(ob + i) = *reinterpret_cast<const float4*>(Out + i);
}

int fill_dev(const sea_spatial_desc* d, SpatialDev& a) {
  if (!d || d->n_groups < 1 || d->n_groups > 4 || d->num_layers < 0 || d->num_layers > 16) return SEA_ERR_UNSUPPORTED;
  if (d->n_patches != P) return SEA_ERR_UNSUPPORTED;
  if (d->n_inp < 1 || (d->embed_dim % 4) || (d->mlp_hidden % 4) || d->n_heads < 1) return SEA_ERR_UNSUPPORTED;
  const int Es = d->n_groups * d->embed_dim;
  if (Es % d->n_heads) return SEA_ERR_UNSUPPORTED;
  const int hd = Es / d->n_heads;
  if (hd != 2 && hd != 4 && hd != 8 && hd != 16) return SEA_ERR_UNSUPPORTED;
  a.n_groups = d->n_groups; a.n_fields = d->n_fields; a.C = d->n_inp; a.Hs = d->mlp_hidden;
  a.D = d->embed_dim; a.n_heads = d->n_heads; a.num_layers = d->num_layers;
  for (int g = 0; g < d->n_groups; ++g) {
    a.g_first[g] = d->group_first_field[g]; a.g_count[g] = d->group_num_fields[g];
    if (a.g_count[g] < 1 || a.g_first[g] < 0 || a.g_first[g] + a.g_count[g] > d->n_fields) return SEA_ERR_INVALID;
    a.enc_w1[g] = d->enc_w1[g]; a.enc_w2[g] = d->enc_w2[g]; a.enc_b2[g] = d->enc_b2[g];
    a.dec_w1[g] = d->dec_w1[g]; a.dec_w2[g] = d->dec_w2[g]; a.dec_b2[g] = d->dec_b2[g];
  }
  a.ln_w = d->ln_w; a.ln_b = d->ln_b; a.pe = d->pe;
  for (int l = 0; l < d->num_layers; ++l) {
    const sea_spatial_layer& s = d->layers[l];
    a.layers[l] = LayerDev{s.ln1_w, s.q_w, s.q_b, s.k_w, s.k_b, s.v_w, s.v_b, s.proj_w, s.ln2_w,
                           s.mlp0_w, s.mlp0_b, s.mlp_ln_w, s.mlp_ln_b, s.mlp3_w, s.mlp3_b};
  }
  return SEA_OK;
}

}  // namespace
}  // namespace sea

using namespace sea;

extern "C" int sea_spatial_encode(const sea_spatial_desc* d, float* x, float* z, int B, int latent_layout,
                                  float pad_idx, int fix_pad, sea_stream_t stream) {
  if (!x || !z || B <= 0) return SEA_ERR_INVALID;
  SpatialDev a{};
  int rc = fill_dev(d, a);
  if (rc) return rc;
  for (int g = 0; g < a.n_groups; ++g)
    if (!a.enc_w1[g] || !a.enc_w2[g] || !a.enc_b2[g]) return SEA_ERR_INVALID;
  if (!a.ln_w || !a.ln_b || !a.pe || (a.num_layers > 0 && !d->layers)) return SEA_ERR_INVALID;
  const int FC = a.n_fields * a.C, Es = a.n_groups * a.D;
  const int big_w = 4 * Es > kChunk ? 4 * Es : kChunk;
  size_t smem = sizeof(float) * (static_cast<size_t>(P) * FC + 2 * P * Es + static_cast<size_t>(P) * big_w);
  const int staged = smem <= 220 * 1024;
  if (!staged) smem -= sizeof(float) * static_cast<size_t>(P) * FC;
  if (smem > 220 * 1024) return SEA_ERR_UNSUPPORTED;
  SEA_CUDA_OK(cudaFuncSetAttribute(spatial_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  SEA_LAUNCH(spatial_encode_kernel, B, kThreads, smem, reinterpret_cast<cudaStream_t>(stream), a, x, z, latent_layout, pad_idx, fix_pad, staged);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int sea_spatial_decode(const sea_spatial_desc* d, const float* z, float* out, int B,
                                  int latent_layout, sea_stream_t stream) {
  if (!z || !out || B <= 0) return SEA_ERR_INVALID;
  SpatialDev a{};
  int rc = fill_dev(d, a);
  if (rc) return rc;
  for (int g = 0; g < a.n_groups; ++g)
    if (!a.dec_w1[g] || !a.dec_w2[g] || !a.dec_b2[g]) return SEA_ERR_INVALID;
  const int FC = a.n_fields * a.C, Es = a.n_groups * a.D;
  size_t smem = sizeof(float) * (static_cast<size_t>(P) * Es + static_cast<size_t>(P) * kChunk + static_cast<size_t>(P) * FC);
  const int staged = smem <= 220 * 1024;
  if (!staged) smem -= sizeof(float) * static_cast<size_t>(P) * FC;
  if (smem > 220 * 1024) return SEA_ERR_UNSUPPORTED;
  SEA_CUDA_OK(cudaFuncSetAttribute(spatial_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  SEA_LAUNCH(spatial_decode_kernel, B, kThreads, smem, reinterpret_cast<cudaStream_t>(stream), a, z, out, latent_layout, staged);
  return static_cast<int>(cudaGetLastError());
}

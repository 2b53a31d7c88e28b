// K4/K5/K6 + operand packing — the HBM-bound kernels of the temporal path.
// All are one pass over their tensors (each element read once, written once), 16-byte vector
// accesses, fp32 statistics with warp-shuffle / block reductions.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/sea_b200.h"
#include "internal.h"
#include "ptx.cuh"

namespace sea {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float silu(float x) { return __fdividef(x, 1.0f + __expf(-x)); }   // MUFU.RCP + FMUL (2 ulp); IEEE "/" is ~10 instructions + a guarded slow path

// ------------------------------------------------------------------ AdaLN cond hidden layer
// h[m, j] = SiLU(sum_c w1[j, c] * ib[m, c] + b1[j])     models/base_blocks.py:337-339, 344
__global__ void adaln_hidden_kernel(const float* __restrict__ ib, long long ld_ib, int M, int ib_num,
                                    const float* __restrict__ w1, const float* __restrict__ b1,
                                    int n, __nv_bfloat16* __restrict__ out_bf16,
                                    float* __restrict__ out_f32) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const long long total = static_cast<long long>(M) * (n / 2);
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int m = static_cast<int>(idx / (n / 2));
    const int j = static_cast<int>(idx - static_cast<long long>(m) * (n / 2)) * 2;
    float a0 = b1[j], a1 = b1[j + 1];
    for (int c = 0; c < ib_num; ++c) {
      const float v = ib[static_cast<long long>(m) * ld_ib + c];
      a0 = fmaf(w1[j * ib_num + c], v, a0);
      a1 = fmaf(w1[(j + 1) * ib_num + c], v, a1);
    }
    a0 = silu(a0);
    a1 = silu(a1);
    if (out_bf16) *reinterpret_cast<uint32_t*>(out_bf16 + static_cast<long long>(m) * n + j) = ptx::pack_bf16(a0, a1);
    if (out_f32) *reinterpret_cast<float2*>(out_f32 + static_cast<long long>(m) * n + j) = make_float2(a0, a1);
  }
}

// ------------------------------------------------------------------------- TIPI hidden layer
// g[m, :] = GELU(LayerNorm_hid(W0 ib[m] + b0))     models/base_blocks.py:22-25 (MLP(ib_num, ...))
// Also stores the pre-LN values and (mean, rstd) when asked (backward).
constexpr int kMaxTipiHid = 64;
__global__ void tipi_hidden_kernel(const float* __restrict__ ib, long long ld_ib, int M, int ib_num,
                                   const float* __restrict__ w0, const float* __restrict__ b0,
                                   const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                   int hid, float* __restrict__ g_out, float* __restrict__ pre_out,
                                   float* __restrict__ stats_out) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float u[kMaxTipiHid];
  float mean = 0.f;
  for (int k = 0; k < hid; ++k) {
    float a = b0[k];
    for (int c = 0; c < ib_num; ++c) a = fmaf(w0[k * ib_num + c], ib[static_cast<long long>(m) * ld_ib + c], a);
    u[k] = a;
    mean += a;
  }
  mean /= hid;
  float var = 0.f;
  for (int k = 0; k < hid; ++k) var += (u[k] - mean) * (u[k] - mean);
  const float rstd = rsqrtf(var / hid + 1e-5f);
  for (int k = 0; k < hid; ++k) {
    const float n = (u[k] - mean) * rstd * ln_w[k] + ln_b[k];
    g_out[static_cast<long long>(m) * hid + k] = ptx::gelu_erf(n);
    if (pre_out) pre_out[static_cast<long long>(m) * hid + k] = u[k];  // pre-LayerNorm, for backward
  }
  if (stats_out) {
    stats_out[2 * m] = mean;
    stats_out[2 * m + 1] = rstd;
  }
}

// ------------------------------------------------------------------ row norm (LN / AdaLN)
// One warp per row, the row lives in registers (CH chunks of 128 columns, d <= 2048).  Every
// global load of a row (x, the AdaLN condition row, the per-trajectory TIPI row) is issued before
// the first use so one DRAM round trip covers them; statistics are two-pass in registers.
//   x' = x + add_rows[m / add_div]            (optional; TIPI precomputed per trajectory)
//   x' = x + W3 g[m] + b3                     (optional; general TIPI, models/temporal.py:140-142)
//   y  = (x' - mean) * rstd * gamma + beta    gamma/beta from weight/bias (+ cond row m / cond_div)
struct NormDev {
  const float* x; long long ldx;
  int M, d, kind;
  const float* weight; const float* bias;
  const float* cond; long long ldc; int cond_div;
  const float* add_rows; long long ld_add; int add_div;
  const float* tipi_g; int tipi_hid; const float* tipi_w; const float* tipi_b;
  float* x_out; long long ldxo;
  float* y_f32; long long ldy_f32;
  __nv_bfloat16* y_bf16; long long ldy_bf16;
  float* stats;
  int cond_folded;                 // cond rows already hold gamma | beta (sea_adaln_fold)
  unsigned long long drop_seed; uint32_t drop_thresh, drop_site; float drop_scale;   // TIPI-term dropout (0 = off)
  int x_rows; long long x_bs;      // x_rows > 0: row m of x lives at x + (m / x_rows) * x_bs + (m % x_rows) * ldx
  int rows_per_cta;                // generic kernel: rows per CTA (multiple of 8; 8 rows in flight, one per warp)
  int tipi_stage;                  // generic kernel: W3^T | b3 of the TIPI layer staged in shared memory
};

constexpr int kNormMaxChunks = 16;  // 16 chunks x 32 lanes x 4 floats = 2048

// Up to SEA_MAX_STREAMS same-shape rows-norms (the V field streams) share a launch: blockIdx.y.
struct NormGroup { NormDev it[SEA_MAX_STREAMS]; };

template <int CH, bool ADALN>
__global__ void __launch_bounds__(256, (CH <= 8) ? 3 : 2) norm_fwd_kernel(const __grid_constant__ NormGroup grp) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const NormDev& a = grp.it[blockIdx.y];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Per-token TIPI term (training / eager forward): out[col] = b3[col] + sum_k W3[col][k] g[m][k].  W3 is [E][8]
  // row-major, so reading it per column is a 32-sector access per load; the CTA stages W3^T (row pitch d + 4:
  // conflict-free transposing stores, 16-byte aligned rows) and b3 once and serves its rows from shared memory.
  extern __shared__ __align__(16) float tsm[];
  const int ldt = a.d + 4;
  if (a.tipi_stage) {
    for (int i = threadIdx.x; i < a.d * 8; i += 256) tsm[(i & 7) * ldt + (i >> 3)] = __ldg(a.tipi_w + i);
    for (int i = threadIdx.x; i < a.d; i += 256) tsm[8 * ldt + i] = __ldg(a.tipi_b + i);
    __syncthreads();
  }
  const int row_end = min(a.M, (static_cast<int>(blockIdx.x) + 1) * a.rows_per_cta);
  for (int m = blockIdx.x * a.rows_per_cta + warp; m < row_end; m += 8) {
  const float* xr = a.x_rows > 0 ? a.x + static_cast<long long>(m / a.x_rows) * a.x_bs + static_cast<long long>(m % a.x_rows) * a.ldx
                                 : a.x + static_cast<long long>(m) * a.ldx;
  const float* cr = ADALN ? a.cond + static_cast<long long>(m / a.cond_div) * a.ldc : nullptr;
  const float* ar = a.add_rows ? a.add_rows + static_cast<long long>(m / a.add_div) * a.ld_add : nullptr;
  // Occupancy over prefetch: the grid of these small kernels should be resident in ONE wave
  // (<= 64 registers -> 32 warps / SM), so only x is staged in registers; the condition row and
  // the affine vectors are L2 hits consumed in the last loop.
  float4 v[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int col = c * 128 + lane * 4;
    if (col < a.d) v[c] = *reinterpret_cast<const float4*>(xr + col);
  }
  if (ar != nullptr) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int col = c * 128 + lane * 4;
      if (col < a.d) {
        const float4 t = *reinterpret_cast<const float4*>(ar + col);
        v[c].x += t.x; v[c].y += t.y; v[c].z += t.z; v[c].w += t.w;
        *reinterpret_cast<float4*>(a.x_out + static_cast<long long>(m) * a.ldxo + col) = v[c];
      }
    }
  } else if (a.tipi_g != nullptr) {
    float gk[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) gk[k] = (k < a.tipi_hid) ? a.tipi_g[static_cast<long long>(m) * a.tipi_hid + k] : 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int col = c * 128 + lane * 4;
      if (col < a.d) {
        float add[4];
        if (a.tipi_stage) {
          // same operation order as tipi_rows_kernel (bias first, one fma per hidden unit): the per-token and the
          // per-trajectory TIPI paths give bit-identical rows
          float4 acc4 = *reinterpret_cast<const float4*>(tsm + 8 * ldt + col);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 w = *reinterpret_cast<const float4*>(tsm + k * ldt + col);
            acc4.x = fmaf(w.x, gk[k], acc4.x); acc4.y = fmaf(w.y, gk[k], acc4.y);
            acc4.z = fmaf(w.z, gk[k], acc4.z); acc4.w = fmaf(w.w, gk[k], acc4.w);
          }
          add[0] = acc4.x; add[1] = acc4.y; add[2] = acc4.z; add[3] = acc4.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float s = a.tipi_b[col + e];
            for (int k = 0; k < a.tipi_hid; ++k)
              s = fmaf(a.tipi_w[(col + e) * a.tipi_hid + k],
                       a.tipi_g[static_cast<long long>(m) * a.tipi_hid + k], s);
            add[e] = s;
          }
        }
        if (a.drop_thresh != 0u) {   // the ib-MLP's own nn.Dropout (models/base_blocks.py:47)
#pragma unroll
          for (int e = 0; e < 4; ++e)
            add[e] *= ptx::drop_mult(a.drop_seed, a.drop_site, static_cast<unsigned long long>(m) * a.d + col + e,
                                     a.drop_thresh, a.drop_scale);
        }
        v[c].x += add[0]; v[c].y += add[1]; v[c].z += add[2]; v[c].w += add[3];
        *reinterpret_cast<float4*>(a.x_out + static_cast<long long>(m) * a.ldxo + col) = v[c];
      }
    }
  }
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c)
    if (c * 128 + lane * 4 < a.d) sum += v[c].x + v[c].y + v[c].z + v[c].w;
  const float mean = warp_sum(sum) / a.d;
  float sq = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    if (c * 128 + lane * 4 < a.d) {
      const float dx = v[c].x - mean, dy = v[c].y - mean, dz = v[c].z - mean, dw = v[c].w - mean;
      sq += dx * dx + dy * dy + dz * dz + dw * dw;
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / a.d + 1e-5f);
  if (a.stats && lane == 0) {
    a.stats[2 * m] = mean;
    a.stats[2 * m + 1] = rstd;
  }
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int col = c * 128 + lane * 4;
    if (col < a.d) {
      float4 w = __ldg(reinterpret_cast<const float4*>(a.weight + col));
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a.bias) b = __ldg(reinterpret_cast<const float4*>(a.bias + col));
      if (ADALN) {
        // gamma = weight + (w_cond + 1), beta = bias + b_cond   (models/base_blocks.py:345-350)
        const float4 cw = *reinterpret_cast<const float4*>(cr + col);
        const float4 cb = *reinterpret_cast<const float4*>(cr + a.d + col);
        w.x += cw.x + 1.f; w.y += cw.y + 1.f; w.z += cw.z + 1.f; w.w += cw.w + 1.f;
        b.x += cb.x; b.y += cb.y; b.z += cb.z; b.w += cb.w;
      }
      float4 y;
      y.x = (v[c].x - mean) * rstd * w.x + b.x;
      y.y = (v[c].y - mean) * rstd * w.y + b.y;
      y.z = (v[c].z - mean) * rstd * w.z + b.z;
      y.w = (v[c].w - mean) * rstd * w.w + b.w;
      if (a.y_f32) *reinterpret_cast<float4*>(a.y_f32 + static_cast<long long>(m) * a.ldy_f32 + col) = y;
      if (a.y_bf16) {
        uint2 o;
        o.x = ptx::pack_bf16(y.x, y.y);
        o.y = ptx::pack_bf16(y.z, y.w);
        *reinterpret_cast<uint2*>(a.y_bf16 + static_cast<long long>(m) * a.ldy_bf16 + col) = o;
      }
    }
  }
  }
}

// Inference variant for the two cases where the affine vectors are just two rows (plain LayerNorm:
// weight only; AdaLN with the per-trajectory gamma | beta already folded by adaln_fold_kernel):
// every global load of the row — x, the TIPI row, gamma, beta — is issued up front, so the warp pays
// ONE memory round trip instead of two (the generic kernel loads the affine vectors after the
// statistics to stay within 64 registers).  96-128 registers, 16 warps / SM, 3x the bytes in flight.
template <int CH, bool FOLDED>
__global__ void __launch_bounds__(256, (CH <= 8) ? 2 : 1) norm_fwd_prefetch_kernel(const __grid_constant__ NormGroup grp) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const NormDev& a = grp.it[blockIdx.y];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * (blockDim.x >> 5) + warp;
  if (m >= a.M) return;
  const float* xr = a.x_rows > 0 ? a.x + static_cast<long long>(m / a.x_rows) * a.x_bs + static_cast<long long>(m % a.x_rows) * a.ldx
                                 : a.x + static_cast<long long>(m) * a.ldx;
  const float* gr = FOLDED ? a.cond + static_cast<long long>(m / a.cond_div) * a.ldc : a.weight;
  const float* br = FOLDED ? gr + a.d : nullptr;
  const float* ar = a.add_rows ? a.add_rows + static_cast<long long>(m / a.add_div) * a.ld_add : nullptr;
  float4 v[CH], gm[CH], bt[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int col = c * 128 + lane * 4;
    if (col < a.d) {
      v[c] = *reinterpret_cast<const float4*>(xr + col);
      gm[c] = __ldg(reinterpret_cast<const float4*>(gr + col));
      if (FOLDED) bt[c] = __ldg(reinterpret_cast<const float4*>(br + col));
    }
  }
  if (ar != nullptr) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int col = c * 128 + lane * 4;
      if (col < a.d) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(ar + col));
        v[c].x += t.x; v[c].y += t.y; v[c].z += t.z; v[c].w += t.w;
        *reinterpret_cast<float4*>(a.x_out + static_cast<long long>(m) * a.ldxo + col) = v[c];
      }
    }
  }
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c)
    if (c * 128 + lane * 4 < a.d) sum += v[c].x + v[c].y + v[c].z + v[c].w;
  const float mean = warp_sum(sum) / a.d;
  float sq = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    if (c * 128 + lane * 4 < a.d) {
      const float dx = v[c].x - mean, dy = v[c].y - mean, dz = v[c].z - mean, dw = v[c].w - mean;
      sq += dx * dx + dy * dy + dz * dz + dw * dw;
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / a.d + 1e-5f);
  if (a.stats && lane == 0) {
    a.stats[2 * m] = mean;
    a.stats[2 * m + 1] = rstd;
  }
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int col = c * 128 + lane * 4;
    if (col < a.d) {
      const float4 w = gm[c];
      const float4 b = FOLDED ? bt[c] : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 y;
      y.x = (v[c].x - mean) * rstd * w.x + b.x;
      y.y = (v[c].y - mean) * rstd * w.y + b.y;
      y.z = (v[c].z - mean) * rstd * w.z + b.z;
      y.w = (v[c].w - mean) * rstd * w.w + b.w;
      if (a.y_f32) *reinterpret_cast<float4*>(a.y_f32 + static_cast<long long>(m) * a.ldy_f32 + col) = y;
      if (a.y_bf16) {
        uint2 o;
        o.x = ptx::pack_bf16(y.x, y.y);
        o.y = ptx::pack_bf16(y.z, y.w);
        *reinterpret_cast<uint2*>(a.y_bf16 + static_cast<long long>(m) * a.ldy_bf16 + col) = o;
      }
    }
  }
}

// cond[r, :d] = weight + cond[r, :d] + 1 ;  cond[r, d:] = bias + cond[r, d:]   (in place, once per trajectory):
// AdaLN's effective gamma | beta (models/base_blocks.py:345-350) for the rows of a condition cache.
__global__ void __launch_bounds__(256) adaln_fold_kernel(float* __restrict__ cond, long long ldc, int R, int d,
                                                         const float* __restrict__ weight,
                                                         const float* __restrict__ bias) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (i >= d || r >= R) return;
  float* row = cond + static_cast<long long>(r) * ldc;
  row[i] = weight[i] + row[i] + 1.f;
  row[d + i] = bias[i] + row[d + i];
}

// rows[r, n] = W3[n, :] . g[r, :] + b3[n]   — the TIPI term for `R` distinct conditions
__global__ void __launch_bounds__(256) tipi_rows_kernel(const float* __restrict__ g, long long ldg, int R, int E,
                                                        int hid, const float* __restrict__ w3,
                                                        const float* __restrict__ b3, float* __restrict__ out) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (n >= E || r >= R) return;
  float s = b3[n];
  for (int k = 0; k < hid; ++k) s = fmaf(w3[n * hid + k], g[static_cast<long long>(r) * ldg + k], s);
  out[static_cast<long long>(r) * E + n] = s;
}

// --------------------------------------------------------- MLP inner LayerNorm(H) + GELU (K6)
// Each CTA owns a group of rows; a thread keeps the LayerNorm weight/bias of ITS columns in
// registers for the whole group (the fp32 affine vectors are 4x the bytes of a bf16 row, so
// re-reading them per row would dominate L2 traffic) and prefetches the next row while it works.
// models/base_blocks.py:23-25 (nn.LayerNorm(scaled_dim) with affine weight+bias, then nn.GELU()).
template <typename TIn, typename TOut, int CH>
__global__ void __launch_bounds__(256) ln_gelu_fwd_kernel(const TIn* __restrict__ h, long long ldh,
                                                          int M, int H,
                                                          const float* __restrict__ weight,
                                                          const float* __restrict__ bias,
                                                          TOut* __restrict__ g, long long ldg,
                                                          float* __restrict__ stats, int rows_per_cta) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  __shared__ float red[8];
  __shared__ float bcast;
  const int tid = threadIdx.x;
  const int row_begin = blockIdx.x * rows_per_cta;
  const int row_end = min(M, row_begin + rows_per_cta);
  float ww[CH][8], bb[CH][8];
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int col = (c * 256 + tid) * 8;
    if (col < H) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(weight + col));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(weight + col + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + col));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + col + 4));
      ww[c][0] = w0.x; ww[c][1] = w0.y; ww[c][2] = w0.z; ww[c][3] = w0.w;
      ww[c][4] = w1.x; ww[c][5] = w1.y; ww[c][6] = w1.z; ww[c][7] = w1.w;
      bb[c][0] = b0.x; bb[c][1] = b0.y; bb[c][2] = b0.z; bb[c][3] = b0.w;
      bb[c][4] = b1.x; bb[c][5] = b1.y; bb[c][6] = b1.z; bb[c][7] = b1.w;
    }
  }
  auto block_sum = [&](float x) -> float {
    x = warp_sum(x);
    if ((tid & 31) == 0) red[tid >> 5] = x;
    __syncthreads();
    if (tid < 32) {
      float t = (tid < 8) ? red[tid] : 0.f;
      t = warp_sum(t);
      if (tid == 0) bcast = t;
    }
    __syncthreads();
    const float r = bcast;
    __syncthreads();
    return r;
  };
  auto load_row = [&](int m, float (&v)[CH][8]) {
    const TIn* hr = h + static_cast<long long>(m) * ldh;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int col = (c * 256 + tid) * 8;
      if (col < H) {
        if constexpr (sizeof(TIn) == 2) {
          const uint4 raw = *reinterpret_cast<const uint4*>(hr + col);
          const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 f = __bfloat1622float2(p[q]);
            v[c][2 * q] = f.x;
            v[c][2 * q + 1] = f.y;
          }
        } else {
          const float4 a = *reinterpret_cast<const float4*>(hr + col);
          const float4 b = *reinterpret_cast<const float4*>(hr + col + 4);
          v[c][0] = a.x; v[c][1] = a.y; v[c][2] = a.z; v[c][3] = a.w;
          v[c][4] = b.x; v[c][5] = b.y; v[c][6] = b.z; v[c][7] = b.w;
        }
      }
    }
  };
  float v[CH][8], nxt[CH][8];
  if (row_begin < row_end) load_row(row_begin, nxt);
  for (int m = row_begin; m < row_end; ++m) {
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
      for (int e = 0; e < 8; ++e) v[c][e] = nxt[c][e];
    if (m + 1 < row_end) load_row(m + 1, nxt);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c)
      if ((c * 256 + tid) * 8 < H)
#pragma unroll
        for (int e = 0; e < 8; ++e) sum += v[c][e];
    const float mean = block_sum(sum) / H;
    float sq = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c)
      if ((c * 256 + tid) * 8 < H)
#pragma unroll
        for (int e = 0; e < 8; ++e) sq += (v[c][e] - mean) * (v[c][e] - mean);
    const float rstd = rsqrtf(block_sum(sq) / H + 1e-5f);
    if (stats && tid == 0) {
      stats[2 * m] = mean;
      stats[2 * m + 1] = rstd;
    }
    TOut* gr = g + static_cast<long long>(m) * ldg;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int col = (c * 256 + tid) * 8;
      if (col < H) {
        float y[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) y[e] = ptx::gelu_erf((v[c][e] - mean) * rstd * ww[c][e] + bb[c][e]);
        if constexpr (sizeof(TOut) == 2) {
          uint4 o;
          o.x = ptx::pack_bf16(y[0], y[1]);
          o.y = ptx::pack_bf16(y[2], y[3]);
          o.z = ptx::pack_bf16(y[4], y[5]);
          o.w = ptx::pack_bf16(y[6], y[7]);
          *reinterpret_cast<uint4*>(gr + col) = o;
        } else {
          *reinterpret_cast<float4*>(gr + col) = make_float4(y[0], y[1], y[2], y[3]);
          *reinterpret_cast<float4*>(gr + col + 4) = make_float4(y[4], y[5], y[6], y[7]);
        }
      }
    }
  }
}

// Shared-memory-staged variant (the one that is launched): a CTA pulls R whole rows into smem with
// 1-D bulk copies (cp.async.bulk + mbarrier: no registers tied up, up to 64 KB in flight per CTA,
// 3 CTAs / SM), takes the statistics with a warp-pair per row, then sweeps column-wise so each
// thread loads the fp32 affine vectors of its 8 columns ONCE for all R rows.
struct LnGeluItem { const void* h; const float* weight; const float* bias; void* g; float* stats; };
struct LnGeluGroup { LnGeluItem it[SEA_MAX_STREAMS]; };

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) ln_gelu_fwd_smem_kernel(const __grid_constant__ LnGeluGroup grp, long long ldh,
                                                               int M, int H, long long ldg, int R) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const LnGeluItem& item = grp.it[blockIdx.y];
  const TIn* __restrict__ h = static_cast<const TIn*>(item.h);
  const float* __restrict__ weight = item.weight;
  const float* __restrict__ bias = item.bias;
  TOut* __restrict__ g = static_cast<TOut*>(item.g);
  float* __restrict__ stats = item.stats;
  extern __shared__ __align__(128) uint8_t ln_smem[];
  __shared__ uint64_t bar;
  __shared__ float red[8];
  __shared__ float mean_s[8], rstd_s[8];
  TIn* rows = reinterpret_cast<TIn*>(ln_smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * R;
  const int nrows = min(R, M - m0);
  if (tid == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0) {
    ptx::mbar_expect_tx(&bar, static_cast<uint32_t>(nrows * H * sizeof(TIn)));
    for (int r = 0; r < nrows; ++r)
      ptx::bulk_load_1d(rows + static_cast<size_t>(r) * H, h + static_cast<long long>(m0 + r) * ldh,
                        static_cast<uint32_t>(H * sizeof(TIn)), &bar);
  }
  ptx::mbar_wait(&bar, 0);
  // statistics: warp w handles row (w % R), slice (w / R) of 8/R equal column slices
  const int parts = 8 / R;
  const int r_mine = warp % R, part = warp / R;
  const int c0 = part * (H / parts), c1 = c0 + H / parts;
  const TIn* myrow = rows + static_cast<size_t>(r_mine) * H;
  auto ldv = [&](int col, float (&o)[8]) {
    if constexpr (sizeof(TIn) == 2) {
      const uint4 raw = *reinterpret_cast<const uint4*>(myrow + col);
      const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = __bfloat1622float2(p[q]);
        o[2 * q] = f.x; o[2 * q + 1] = f.y;
      }
    } else {
      const float4 a = *reinterpret_cast<const float4*>(myrow + col);
      const float4 b = *reinterpret_cast<const float4*>(myrow + col + 4);
      o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
    }
  };
  float acc = 0.f;
  if (r_mine < nrows)
    for (int col = c0 + lane * 8; col < c1; col += 256) {
      float o[8];
      ldv(col, o);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc += o[e];
    }
  acc = warp_sum(acc);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (tid < R) {
    float t = 0.f;
    for (int p2 = 0; p2 < parts; ++p2) t += red[p2 * R + tid];
    mean_s[tid] = t / H;
  }
  __syncthreads();
  const float mu = mean_s[r_mine];
  acc = 0.f;
  if (r_mine < nrows)
    for (int col = c0 + lane * 8; col < c1; col += 256) {
      float o[8];
      ldv(col, o);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc += (o[e] - mu) * (o[e] - mu);
    }
  acc = warp_sum(acc);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (tid < R) {
    float t = 0.f;
    for (int p2 = 0; p2 < parts; ++p2) t += red[p2 * R + tid];
    const float rs = rsqrtf(t / H + 1e-5f);
    rstd_s[tid] = rs;
    if (stats && tid < nrows) {
      stats[2 * (m0 + tid)] = mean_s[tid];
      stats[2 * (m0 + tid) + 1] = rs;
    }
  }
  __syncthreads();
  // normalise + affine + GELU, column-wise
  for (int col = tid * 8; col < H; col += 2048) {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(weight + col));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(weight + col + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + col));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + col + 4));
    const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    for (int r = 0; r < nrows; ++r) {
      const float m_r = mean_s[r], rs = rstd_s[r];
      float o[8], y[8];
      const TIn* rp = rows + static_cast<size_t>(r) * H + col;
      if constexpr (sizeof(TIn) == 2) {
        const uint4 raw = *reinterpret_cast<const uint4*>(rp);
        const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = __bfloat1622float2(p[q]);
          o[2 * q] = f.x; o[2 * q + 1] = f.y;
        }
      } else {
        const float4 a = *reinterpret_cast<const float4*>(rp);
        const float4 b = *reinterpret_cast<const float4*>(rp + 4);
        o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
      }
      TOut* gr = g + static_cast<long long>(m0 + r) * ldg + col;
      if constexpr (sizeof(TOut) == 2) {
const float a1 = rs, c1 = -m_r * rs;
#pragma unroll
        for (int e = 0; e < 8; ++e) y[e] = ptx::gelu_bf16(fmaf(fmaf(o[e], a1, c1), ww[e], bb[e]));
        uint4 pk;
        pk.x = ptx::pack_bf16(y[0], y[1]); pk.y = ptx::pack_bf16(y[2], y[3]);
        pk.z = ptx::pack_bf16(y[4], y[5]); pk.w = ptx::pack_bf16(y[6], y[7]);
        *reinterpret_cast<uint4*>(gr) = pk;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) y[e] = ptx::gelu_erf((o[e] - m_r) * rs * ww[e] + bb[e]);
        *reinterpret_cast<float4*>(gr) = make_float4(y[0], y[1], y[2], y[3]);
        *reinterpret_cast<float4*>(gr + 4) = make_float4(y[4], y[5], y[6], y[7]);
      }
    }
  }
}

// Grouped AdaLN hidden layers: up to 8 (w1, b1, out) items of the same width in one launch.
struct AdalnGroupDev {
  const float* w1[8]; const float* b1[8]; __nv_bfloat16* out_bf16[8]; float* out_f32[8];
  const float* ib; long long ld_ib; int M, ib_num, n, items;
};
__global__ void __launch_bounds__(256) adaln_hidden_group_kernel(const AdalnGroupDev a) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const int it = blockIdx.y;
  const long long total = static_cast<long long>(a.M) * (a.n / 2);
  const float* w1 = a.w1[it];
  const float* b1 = a.b1[it];
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int m = static_cast<int>(idx / (a.n / 2));
    const int j = static_cast<int>(idx - static_cast<long long>(m) * (a.n / 2)) * 2;
    float a0 = b1[j], a1 = b1[j + 1];
    for (int c = 0; c < a.ib_num; ++c) {
      const float v = a.ib[static_cast<long long>(m) * a.ld_ib + c];
      a0 = fmaf(w1[j * a.ib_num + c], v, a0);
      a1 = fmaf(w1[(j + 1) * a.ib_num + c], v, a1);
    }
    a0 = silu(a0);
    a1 = silu(a1);
    if (a.out_bf16[it]) *reinterpret_cast<uint32_t*>(a.out_bf16[it] + static_cast<long long>(m) * a.n + j) = ptx::pack_bf16(a0, a1);
    if (a.out_f32[it]) *reinterpret_cast<float2*>(a.out_f32[it] + static_cast<long long>(m) * a.n + j) = make_float2(a0, a1);
  }
}

// ----------------------------------------------------------------------------- operand packing
// src [R, C] (fp32 or bf16, pitch ld) -> bf16 dst, plain or transposed, optionally as the
// 3-way bf16 split used for fp32-accurate products on the bf16 tensor cores:
//   x = x1 + x2 + x3 (x1 = bf16(x), x2 = bf16(x - x1), x3 = bf16(x - x1 - x2));
//   A-side pattern along K: [x3 x2 x1 x2 x1 x1], B-side: [y1 y2 y3 y1 y2 y1]  (6 products, all
//   terms down to 2^-24 relative; the dropped x2*y3, x3*y2, x3*y3 are below fp32 epsilon).
// dst has (transpose ? C : R) rows of (split ? 6 : 1) * (transpose ? R : C) columns.
__device__ __forceinline__ void split3(float x, __nv_bfloat16& p1, __nv_bfloat16& p2, __nv_bfloat16& p3) {
  p1 = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(p1);
  p2 = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(p2);
  p3 = __float2bfloat16_rn(r2);
}

template <typename TIn>
__global__ void __launch_bounds__(256) pack_kernel(const TIn* __restrict__ src, long long ld, int R,
                                                   int C, int transpose, int split, int act, int split_inner,
                                                   __nv_bfloat16* __restrict__ dst, long long ldd) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  // 32x32 tile through shared memory so both the read and the (possibly transposed) write are
  // coalesced.
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    float v = 0.f;
    if (r < R && c < C) {
      if constexpr (sizeof(TIn) == 2) v = __bfloat162float(src[static_cast<long long>(r) * ld + c]);
      else v = src[static_cast<long long>(r) * ld + c];
      if (act == SEA_ACT_GELU) v = ptx::gelu_erf(v);
    }
    tile[i][tx] = v;
  }
  __syncthreads();
  const int inner = split_inner > 0 ? split_inner : (transpose ? R : C);  // segment pitch
  for (int i = ty; i < 32; i += 8) {
    int orow, ocol;
    float v;
    if (transpose) {
      orow = c0 + i; ocol = r0 + tx; v = tile[tx][i];
      if (orow >= C || ocol >= R) continue;
    } else {
      orow = r0 + i; ocol = c0 + tx; v = tile[i][tx];
      if (orow >= R || ocol >= C) continue;
    }
    __nv_bfloat16* drow = dst + static_cast<long long>(orow) * ldd;
    if (split == 0) {
      drow[ocol] = __float2bfloat16_rn(v);
    } else {
      __nv_bfloat16 p1, p2, p3;
      split3(v, p1, p2, p3);
      if (split == 1) {  // A pattern  (smallest products first, x1*y1 last)
        drow[ocol] = p3; drow[inner + ocol] = p2; drow[2 * inner + ocol] = p1;
        drow[3 * inner + ocol] = p2; drow[4 * inner + ocol] = p1; drow[5 * inner + ocol] = p1;
      } else {           // B pattern
        drow[ocol] = p1; drow[inner + ocol] = p2; drow[2 * inner + ocol] = p3;
        drow[3 * inner + ocol] = p1; drow[4 * inner + ocol] = p2; drow[5 * inner + ocol] = p1;
      }
    }
  }
}

// out[n] = sum_m src[m, n] (bias gradients).  One warp per 32 columns x row-slab, atomics into out.
struct ColsumGroup { const float* f32[SEA_MAX_STREAMS]; const __nv_bfloat16* b16[SEA_MAX_STREAMS]; float* out[SEA_MAX_STREAMS]; };
__global__ void colsum_kernel(const ColsumGroup grp, long long ld, int M, int N) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const float* __restrict__ src_f32 = grp.f32[blockIdx.z];
  const __nv_bfloat16* __restrict__ src_bf16 = grp.b16[blockIdx.z];
  float* __restrict__ out = grp.out[blockIdx.z];
  if (out == nullptr) return;
  const int n = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int r_begin = blockIdx.y * rows_per;
  const int r_end = min(M, r_begin + rows_per);
  float acc = 0.f;
  if (n < N) {
    for (int r = r_begin + (threadIdx.x >> 5); r < r_end; r += (blockDim.x >> 5)) {
      acc += src_f32 ? src_f32[static_cast<long long>(r) * ld + n]
                     : __bfloat162float(src_bf16[static_cast<long long>(r) * ld + n]);
    }
  }
  __shared__ float red[8][32];
  red[threadIdx.x >> 5][threadIdx.x & 31] = acc;
  __syncthreads();
  if (threadIdx.x < 32 && n < N) {
    float t = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w][threadIdx.x];
    atomicAdd(out + n, t);
  }
}


// keep decisions of elements [0, n) of one dropout site (test / oracle helper)
__global__ void __launch_bounds__(256) dropout_mask_kernel(unsigned long long seed, uint32_t site, long long n,
                                                           uint32_t thresh, unsigned char* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = ptx::drop_mult(seed, site, static_cast<unsigned long long>(i), thresh, 1.0f) != 0.f ? 1 : 0;
}

// dst[m, n] = src[m, n] * mask(m * N + n) / (1 - p): the gradient that enters a dropped branch
__global__ void __launch_bounds__(256) dropout_apply_kernel(const float* __restrict__ src, long long ld_src, int M, int N,
                                                            unsigned long long seed, uint32_t site, uint32_t thresh,
                                                            float scale, float* __restrict__ dst_f32, long long ld_f32,
                                                            __nv_bfloat16* __restrict__ dst_b16, long long ld_b16) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const long long pairs_per_row = N >> 1;
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= pairs_per_row * M) return;
  const int m = static_cast<int>(t / pairs_per_row);
  const int n = static_cast<int>(t - static_cast<long long>(m) * pairs_per_row) * 2;
  const float2 v = *reinterpret_cast<const float2*>(src + static_cast<long long>(m) * ld_src + n);
  const uint2 h = ptx::drop_hash(seed, site, (static_cast<unsigned long long>(m) * N + n) >> 1);
  const float a = h.x >= thresh ? v.x * scale : 0.f, b = h.y >= thresh ? v.y * scale : 0.f;
  if (dst_f32) *reinterpret_cast<float2*>(dst_f32 + static_cast<long long>(m) * ld_f32 + n) = make_float2(a, b);
  if (dst_b16) *reinterpret_cast<uint32_t*>(dst_b16 + static_cast<long long>(m) * ld_b16 + n) = ptx::pack_bf16(a, b);
}

// Vectorised variant (N % 8 == 0, 16-byte aligned rows): a thread owns 8 consecutive columns (one
// 16-byte load per bf16 row), a warp 256, the 8 warps of a CTA stride over the rows of a slab with four
// rows in flight each; partials meet in shared memory, one atomicAdd per column and CTA.
template <bool BF16>
__global__ void __launch_bounds__(256) colsum_vec_kernel(const ColsumGroup grp, long long ld, int M, int N) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  float* __restrict__ out = grp.out[blockIdx.z];
  if (out == nullptr) return;
  const __nv_bfloat16* __restrict__ sb = grp.b16[blockIdx.z];
  const float* __restrict__ sf = grp.f32[blockIdx.z];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = blockIdx.x * 256 + lane * 8;
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int r_begin = blockIdx.y * rows_per, r_end = min(M, r_begin + rows_per);
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (col < N) {
#pragma unroll 4
    for (int r = r_begin + warp; r < r_end; r += 8) {
      if (BF16) {
        const uint4 raw = *reinterpret_cast<const uint4*>(sb + static_cast<long long>(r) * ld + col);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int q = 0; q < 4; ++q) { const float2 f = __bfloat1622float2(h[q]); acc[2 * q] += f.x; acc[2 * q + 1] += f.y; }
      } else {
        const float4 a = *reinterpret_cast<const float4*>(sf + static_cast<long long>(r) * ld + col);
        const float4 b = *reinterpret_cast<const float4*>(sf + static_cast<long long>(r) * ld + col + 4);
        acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w; acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
      }
    }
  }
  __shared__ float red[8][256 + 8];
#pragma unroll
  for (int e = 0; e < 8; ++e) red[warp][lane * 8 + e] = acc[e];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < N) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    atomicAdd(out + c, t);
  }
}

}  // namespace
}  // namespace sea

using namespace sea;

extern "C" int sea_adaln_hidden(const float* ib, int64_t ld_ib, int M, int ib_num, const float* w1,
                                const float* b1, int n, void* out_bf16, float* out_f32,
                                sea_stream_t stream) {
  if (!ib || !w1 || !b1 || M <= 0 || n <= 0 || (n % 2) || ib_num <= 0) return SEA_ERR_INVALID;
  if (!out_bf16 && !out_f32) return SEA_ERR_INVALID;
  const long long total = static_cast<long long>(M) * (n / 2);
  long long nblk = (total + 255) / 256;
  if (nblk > 148LL * 16) nblk = 148LL * 16;
  const int grid = static_cast<int>(nblk);
  SEA_LAUNCH(adaln_hidden_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream), ib, ld_ib > 0 ? ld_ib : ib_num, M, ib_num, w1, b1, n, static_cast<__nv_bfloat16*>(out_bf16), out_f32);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int sea_adaln_hidden_group(int items, const float* const* w1, const float* const* b1,
                                      void* const* out_bf16, float* const* out_f32, const float* ib,
                                      int64_t ld_ib, int M, int ib_num, int n, sea_stream_t stream) {
  if (items < 1 || items > 8 || !w1 || !b1 || !ib || M <= 0 || n <= 0 || (n % 2) || ib_num <= 0) return SEA_ERR_INVALID;
  AdalnGroupDev d{};
  for (int i = 0; i < items; ++i) {
    d.w1[i] = w1[i]; d.b1[i] = b1[i];
    d.out_bf16[i] = out_bf16 ? static_cast<__nv_bfloat16*>(out_bf16[i]) : nullptr;
    d.out_f32[i] = out_f32 ? out_f32[i] : nullptr;
    if (!d.w1[i] || !d.b1[i] || (!d.out_bf16[i] && !d.out_f32[i])) return SEA_ERR_INVALID;
  }
  d.ib = ib; d.ld_ib = ld_ib > 0 ? ld_ib : ib_num; d.M = M; d.ib_num = ib_num; d.n = n; d.items = items;
  const long long total = static_cast<long long>(M) * (n / 2);
  long long nblk = (total + 255) / 256;
  if (nblk > 148LL * 8) nblk = 148LL * 8;
  dim3 grid(static_cast<unsigned>(nblk), items);
  SEA_LAUNCH(adaln_hidden_group_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream), d);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int sea_tipi_hidden(const float* ib, int64_t ld_ib, int M, int ib_num, const float* w0, const float* b0,
                               const float* ln_w, const float* ln_b, int hid, float* g_out,
                               float* pre_out, float* stats_out, sea_stream_t stream) {
  if (!ib || !w0 || !b0 || !ln_w || !ln_b || !g_out || M <= 0 || ib_num <= 0) return SEA_ERR_INVALID;
  if (hid <= 0 || hid > kMaxTipiHid) return SEA_ERR_UNSUPPORTED;
  SEA_LAUNCH(tipi_hidden_kernel, (M + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream), ib, ld_ib > 0 ? ld_ib : ib_num, M, ib_num, w0, b0, ln_w, ln_b, hid, g_out, pre_out, stats_out);
  return static_cast<int>(cudaGetLastError());
}

template <int CH>
static int launch_norm(const NormGroup& g, int n, cudaStream_t s) {
  const NormDev& d = g.it[0];
  const int rows_per_cta = 8;
  const dim3 grid((d.M + rows_per_cta - 1) / rows_per_cta, n);
  bool prefetch = d.tipi_g == nullptr;   // the general (per-token) TIPI path stays in the generic kernel
  for (int i = 0; i < n; ++i)
    prefetch = prefetch && g.it[i].tipi_g == nullptr && g.it[i].cond_folded == d.cond_folded &&
               (d.kind != SEA_NORM_ADALN || g.it[i].cond_folded);
  if (prefetch && d.kind == SEA_NORM_ADALN) {
    SEA_LAUNCH((norm_fwd_prefetch_kernel<CH, true>), grid, rows_per_cta * 32, 0, s, g);
    return static_cast<int>(cudaGetLastError());
  }
  if (prefetch && d.kind == SEA_NORM_LN) {
    SEA_LAUNCH((norm_fwd_prefetch_kernel<CH, false>), grid, rows_per_cta * 32, 0, s, g);
    return static_cast<int>(cudaGetLastError());
  }
  // generic kernel; with the per-token TIPI term the CTA stages W3^T | b3 (9 x (d + 4) floats) and, once there
  // is more than a wave of 8-row CTAs, takes more rows so that the staging is amortised
  NormGroup gg = g;
  bool stage = true;
  for (int i = 0; i < n; ++i) stage = stage && g.it[i].tipi_g != nullptr && g.it[i].tipi_hid == 8;
  int rpc = rows_per_cta;
  if (stage && d.M > 8 * 2 * 148) rpc = ((d.M + 2 * 148 - 1) / (2 * 148) + 7) / 8 * 8;
  for (int i = 0; i < n; ++i) { gg.it[i].rows_per_cta = rpc; gg.it[i].tipi_stage = stage ? 1 : 0; }
  const size_t smem = stage ? sizeof(float) * 9 * (d.d + 4) : 0;
  const dim3 ggrid((d.M + rpc - 1) / rpc, n);
  static bool attr_set[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 16 && !attr_set[dev]) {
    SEA_CUDA_OK(cudaFuncSetAttribute(norm_fwd_kernel<CH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 9 * (CH * 128 + 4) * 4));
    SEA_CUDA_OK(cudaFuncSetAttribute(norm_fwd_kernel<CH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 9 * (CH * 128 + 4) * 4));
    attr_set[dev] = true;
  }
  if (d.kind == SEA_NORM_ADALN) SEA_LAUNCH((norm_fwd_kernel<CH, true>), ggrid, 256, smem, s, gg);
  else SEA_LAUNCH((norm_fwd_kernel<CH, false>), ggrid, 256, smem, s, gg);
  return static_cast<int>(cudaGetLastError());
}

static int fill_norm(const sea_norm_args* a, NormDev& d) {
  if (!a->x || !a->weight || a->M <= 0 || a->d <= 0) return SEA_ERR_INVALID;
  if ((a->d % 4) || a->d > kNormMaxChunks * 128) return SEA_ERR_UNSUPPORTED;
  if (!a->y_f32 && !a->y_bf16) return SEA_ERR_INVALID;
  if (a->kind == SEA_NORM_ADALN && !a->cond) return SEA_ERR_INVALID;
  if ((a->ldx % 4) || (a->y_f32 && (a->ldy_f32 % 4)) || (a->y_bf16 && (a->ldy_bf16 % 4)) ||
      (a->cond && (a->ldc % 4)))
    return SEA_ERR_INVALID;
  if (a->tipi_g && (!a->tipi_w || !a->tipi_b || !a->x_out || a->tipi_hid <= 0 || (a->ldxo % 4)))
    return SEA_ERR_INVALID;
  if (a->add_rows && (!a->x_out || (a->ld_add % 4) || (a->ldxo % 4))) return SEA_ERR_INVALID;
  d.x = a->x; d.ldx = a->ldx; d.M = a->M; d.d = a->d; d.kind = a->kind;
  d.weight = a->weight; d.bias = a->bias; d.cond = a->cond; d.ldc = a->ldc;
  d.cond_div = a->cond_div > 0 ? a->cond_div : 1;
  d.add_rows = a->add_rows; d.ld_add = a->ld_add; d.add_div = a->add_div > 0 ? a->add_div : 1;
  d.tipi_g = a->tipi_g; d.tipi_hid = a->tipi_hid; d.tipi_w = a->tipi_w; d.tipi_b = a->tipi_b;
  d.x_out = a->x_out; d.ldxo = a->ldxo;
  d.y_f32 = a->y_f32; d.ldy_f32 = a->ldy_f32;
  d.y_bf16 = static_cast<__nv_bfloat16*>(a->y_bf16); d.ldy_bf16 = a->ldy_bf16;
  d.stats = a->stats;
  d.cond_folded = a->cond_folded;
  if (a->tipi_dropout_p < 0.f || a->tipi_dropout_p >= 1.f) return SEA_ERR_INVALID;
  if (a->tipi_dropout_p > 0.f && !a->tipi_g) return SEA_ERR_UNSUPPORTED;   // only the per-token TIPI path drops
  d.drop_seed = a->tipi_dropout_seed; d.drop_site = a->tipi_dropout_site;
  d.drop_thresh = a->tipi_dropout_p > 0.f ? static_cast<uint32_t>(static_cast<double>(a->tipi_dropout_p) * 4294967296.0) : 0u;
  d.drop_scale = 1.0f / (1.0f - a->tipi_dropout_p);
  d.x_rows = a->x_rows_per_batch; d.x_bs = a->x_batch_stride;
  if (a->cond_folded && a->kind != SEA_NORM_ADALN) return SEA_ERR_INVALID;
  if (a->x_rows_per_batch < 0 || (a->x_rows_per_batch > 0 && (a->x_batch_stride % 4))) return SEA_ERR_INVALID;
  return SEA_OK;
}

extern "C" int sea_adaln_fold(float* cond, int64_t ldc, int R, int d, const float* weight, const float* bias,
                              sea_stream_t stream) {
  if (!cond || !weight || !bias || R <= 0 || d <= 0 || ldc < 2LL * d) return SEA_ERR_INVALID;
  dim3 grid((d + 255) / 256, R);
  SEA_LAUNCH(adaln_fold_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream), cond, static_cast<long long>(ldc), R, d, weight, bias);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int sea_norm_fwd_group(int n, const sea_norm_args* a, sea_stream_t stream) {
  if (!a || n < 1 || n > SEA_MAX_STREAMS) return SEA_ERR_INVALID;
  NormGroup g;
  for (int i = 0; i < n; ++i) {
    int rc = fill_norm(&a[i], g.it[i]);
    if (rc) return rc;
    if (a[i].M != a[0].M || a[i].d != a[0].d || a[i].kind != a[0].kind) return SEA_ERR_INVALID;
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (a->d <= 512) return launch_norm<4>(g, n, s);
  if (a->d <= 1024) return launch_norm<8>(g, n, s);
  return launch_norm<16>(g, n, s);
}

extern "C" int sea_norm_fwd(const sea_norm_args* a, sea_stream_t stream) {
  return sea_norm_fwd_group(1, a, stream);
}

extern "C" int sea_tipi_rows(const float* g, int64_t ldg, int R, int E, int hid, const float* w3,
                             const float* b3, float* out, sea_stream_t stream) {
  if (!g || !w3 || !b3 || !out || R <= 0 || E <= 0 || hid <= 0) return SEA_ERR_INVALID;
  dim3 grid((E + 255) / 256, R);
  SEA_LAUNCH(tipi_rows_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream), g, ldg, R, E, hid, w3, b3, out);
  return static_cast<int>(cudaGetLastError());
}

template <typename TIn, typename TOut>
static int launch_ln_gelu_smem(int n, const sea_ln_gelu_args* a, cudaStream_t s) {
  // R rows per CTA, power of two <= 8, at most 64 KB of rows in shared memory
  int R = 8;
  while (R > 1 && static_cast<size_t>(R) * a->H * sizeof(TIn) > 64 * 1024) R >>= 1;
  while (R > 1 && static_cast<long long>((a->M + R - 1) / R) * n < 2 * 148) R >>= 1;   // keep the machine full for small M
  if ((a->H % (8 * (8 / R) )) != 0) R = 1;
  const size_t smem = static_cast<size_t>(R) * a->H * sizeof(TIn);
  static bool attr_set[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 16 && !attr_set[dev]) {
    SEA_CUDA_OK(cudaFuncSetAttribute(ln_gelu_fwd_smem_kernel<TIn, TOut>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr_set[dev] = true;
  }
  LnGeluGroup grp;
  for (int i = 0; i < n; ++i) {
    const bool bf = sizeof(TIn) == 2;
    grp.it[i] = LnGeluItem{bf ? a[i].h_bf16 : static_cast<const void*>(a[i].h_f32), a[i].weight, a[i].bias,
                           bf ? a[i].g_bf16 : static_cast<void*>(a[i].g_f32), a[i].stats};
  }
  const dim3 grid((a->M + R - 1) / R, n);
  SEA_LAUNCH((ln_gelu_fwd_smem_kernel<TIn, TOut>), grid, 256, smem, s, grp, a->ldh, a->M, a->H, a->ldg, R);
  return static_cast<int>(cudaGetLastError());
}

template <typename TIn, typename TOut>
static int launch_ln_gelu(const sea_ln_gelu_args* a, const TIn* h, TOut* g, cudaStream_t s) {
  // rows per CTA: amortise the affine vectors, but keep >= ~2 CTAs per SM in flight
  int rows = a->M / (2 * 148);
  rows = rows < 1 ? 1 : (rows > 8 ? 8 : rows);
  const int grid = (a->M + rows - 1) / rows;
  if (a->H <= 4096)
    SEA_LAUNCH((ln_gelu_fwd_kernel<TIn, TOut, 2>), grid, 256, 0, s, h, a->ldh, a->M, a->H, a->weight, a->bias, g, a->ldg, a->stats, rows);
  else if (a->H <= 8192)
    SEA_LAUNCH((ln_gelu_fwd_kernel<TIn, TOut, 4>), grid, 256, 0, s, h, a->ldh, a->M, a->H, a->weight, a->bias, g, a->ldg, a->stats, rows);
  else
    SEA_LAUNCH((ln_gelu_fwd_kernel<TIn, TOut, 8>), grid, 256, 0, s, h, a->ldh, a->M, a->H, a->weight, a->bias, g, a->ldg, a->stats, rows);
  return static_cast<int>(cudaGetLastError());
}

template <typename T>
static bool ln_gelu_smem_ok(const sea_ln_gelu_args* a, const void* h) {
  return static_cast<size_t>(a->H) * sizeof(T) <= 64 * 1024 && (a->H % 64) == 0 &&
         (reinterpret_cast<uintptr_t>(h) % 16) == 0 && ((a->ldh * sizeof(T)) % 16) == 0;
}

extern "C" int sea_ln_gelu_fwd_group(int n, const sea_ln_gelu_args* a, sea_stream_t stream) {
  if (!a || n < 1 || n > SEA_MAX_STREAMS) return SEA_ERR_INVALID;
  bool all_bf = true, all_f32 = true, smem_ok = true;
  for (int i = 0; i < n; ++i) {
    const sea_ln_gelu_args* x = &a[i];
    if (x->M <= 0 || x->H <= 0 || !x->weight || !x->bias) return SEA_ERR_INVALID;
    if ((x->H % 8) || x->H > 16384 || (x->ldh % 8) || (x->ldg % 8)) return SEA_ERR_UNSUPPORTED;
    if (x->M != a->M || x->H != a->H || x->ldh != a->ldh || x->ldg != a->ldg) return SEA_ERR_INVALID;
    const bool bf = x->h_bf16 && x->g_bf16, f32 = x->h_f32 && x->g_f32;
    if (!bf && !f32) return SEA_ERR_INVALID;
    all_bf = all_bf && bf; all_f32 = all_f32 && !bf && f32;
    smem_ok = smem_ok && (bf ? ln_gelu_smem_ok<__nv_bfloat16>(x, x->h_bf16) : ln_gelu_smem_ok<float>(x, x->h_f32));
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (smem_ok && all_bf) return launch_ln_gelu_smem<__nv_bfloat16, __nv_bfloat16>(n, a, s);
  if (smem_ok && all_f32) return launch_ln_gelu_smem<float, float>(n, a, s);
  for (int i = 0; i < n; ++i) {
    const sea_ln_gelu_args* x = &a[i];
    int rc;
    if (x->h_bf16 && x->g_bf16)
      rc = launch_ln_gelu(x, static_cast<const __nv_bfloat16*>(x->h_bf16), static_cast<__nv_bfloat16*>(x->g_bf16), s);
    else
      rc = launch_ln_gelu(x, x->h_f32, x->g_f32, s);
    if (rc) return rc;
  }
  return SEA_OK;
}

extern "C" int sea_ln_gelu_fwd(const sea_ln_gelu_args* a, sea_stream_t stream) {
  return sea_ln_gelu_fwd_group(1, a, stream);
}

extern "C" int sea_pack_operand(const sea_pack_args* a, sea_stream_t stream) {
  if (!a || !a->dst || a->R <= 0 || a->C <= 0 || (!a->src_f32 && !a->src_bf16)) return SEA_ERR_INVALID;
  if (a->split < 0 || a->split > 2) return SEA_ERR_INVALID;
  dim3 grid((a->C + 31) / 32, (a->R + 31) / 32);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (a->src_f32)
    SEA_LAUNCH((pack_kernel<float>), grid, 256, 0, s, a->src_f32, a->ld, a->R, a->C, a->transpose, a->split, a->act, a->split_inner, static_cast<__nv_bfloat16*>(a->dst), a->ld_dst);
  else
    SEA_LAUNCH((pack_kernel<__nv_bfloat16>), grid, 256, 0, s, static_cast<const __nv_bfloat16*>(a->src_bf16), a->ld, a->R, a->C, a->transpose, a->split, a->act, a->split_inner, static_cast<__nv_bfloat16*>(a->dst), a->ld_dst);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int sea_dropout_mask(uint64_t seed, uint32_t site, int64_t n, float p, uint8_t* out, sea_stream_t stream) {
  if (!out || n <= 0 || p < 0.f || p >= 1.f) return SEA_ERR_INVALID;
  const uint32_t thresh = static_cast<uint32_t>(static_cast<double>(p) * 4294967296.0);
  SEA_LAUNCH(dropout_mask_kernel, static_cast<unsigned>((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream),
             static_cast<unsigned long long>(seed), site, static_cast<long long>(n), thresh, out);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int sea_dropout_apply(const float* src, int64_t ld_src, int M, int N, uint64_t seed, uint32_t site, float p,
                                 float* dst_f32, int64_t ld_f32, void* dst_bf16, int64_t ld_bf16, sea_stream_t stream) {
  if (!src || (!dst_f32 && !dst_bf16) || M <= 0 || N <= 0 || p < 0.f || p >= 1.f) return SEA_ERR_INVALID;
  if ((N % 2) || (ld_src % 2) || (dst_f32 && (ld_f32 % 2)) || (dst_bf16 && (ld_bf16 % 2))) return SEA_ERR_UNSUPPORTED;
  const uint32_t thresh = static_cast<uint32_t>(static_cast<double>(p) * 4294967296.0);
  const long long work = static_cast<long long>(M) * (N / 2);
  SEA_LAUNCH(dropout_apply_kernel, static_cast<unsigned>((work + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream),
             src, static_cast<long long>(ld_src), M, N, static_cast<unsigned long long>(seed), site, thresh,
             1.0f / (1.0f - p), dst_f32, static_cast<long long>(ld_f32), static_cast<__nv_bfloat16*>(dst_bf16),
             static_cast<long long>(ld_bf16));
  return static_cast<int>(cudaGetLastError());
}

extern "C" int sea_colsum_accumulate_group(int n, const float* const* src_f32, const void* const* src_bf16,
                                           int64_t ld, int M, int N, float* const* out, sea_stream_t stream) {
  if (n < 1 || n > SEA_MAX_STREAMS || !out || M <= 0 || N <= 0) return SEA_ERR_INVALID;
  ColsumGroup g;
  bool any = false;
  for (int i = 0; i < n; ++i) {
    g.f32[i] = src_f32 ? src_f32[i] : nullptr;
    g.b16[i] = src_bf16 ? static_cast<const __nv_bfloat16*>(src_bf16[i]) : nullptr;
    g.out[i] = out[i];
    if (out[i] && !g.f32[i] && !g.b16[i]) return SEA_ERR_INVALID;
    any = any || out[i] != nullptr;
  }
  if (!any) return SEA_OK;
  bool all_b16 = true, all_f32 = true, aligned = (N % 8) == 0 && (ld % 8) == 0;
  for (int i = 0; i < n; ++i) {
    if (!out[i]) continue;
    all_b16 = all_b16 && g.b16[i] != nullptr && g.f32[i] == nullptr;
    all_f32 = all_f32 && g.f32[i] != nullptr;
    aligned = aligned && ((reinterpret_cast<uintptr_t>(g.b16[i]) | reinterpret_cast<uintptr_t>(g.f32[i])) & 15) == 0;
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (aligned && (all_b16 || all_f32)) {
    const int colblocks = (N + 255) / 256;
    int slabs = (3 * 148 + colblocks * n - 1) / (colblocks * n);   // ~3 CTAs per SM in total
    if (slabs > (M + 31) / 32) slabs = (M + 31) / 32;              // at least 32 rows per CTA
    if (slabs < 1) slabs = 1;
    dim3 grid(colblocks, slabs, n);
    if (all_b16) SEA_LAUNCH(colsum_vec_kernel<true>, grid, 256, 0, s, g, static_cast<long long>(ld), M, N);
    else SEA_LAUNCH(colsum_vec_kernel<false>, grid, 256, 0, s, g, static_cast<long long>(ld), M, N);
    return static_cast<int>(cudaGetLastError());
  }
  int slabs = M / 64;
  if (slabs < 1) slabs = 1;
  if (slabs > 64) slabs = 64;
  dim3 grid((N + 31) / 32, slabs, n);
  SEA_LAUNCH(colsum_kernel, grid, 256, 0, s, g, static_cast<long long>(ld), M, N);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int sea_colsum_accumulate(const float* src_f32, const void* src_bf16, int64_t ld, int M,
                                     int N, float* out, sea_stream_t stream) {
  if (!out) return SEA_ERR_INVALID;
  return sea_colsum_accumulate_group(1, src_f32 ? &src_f32 : nullptr, src_bf16 ? &src_bf16 : nullptr, ld, M, N, &out, stream);
}

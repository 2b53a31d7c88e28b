// Backward of the HBM-bound kernels (the reference relies on autograd for all of these).
// Row-wise quantities are recomputed from the saved fp32 inputs + (mean, rstd); parameter
// gradients are column reductions over M, accumulated per CTA in registers / shared memory and
// flushed with one atomicAdd per column per CTA.
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

// ------------------------------------------------------------------------ LN / AdaLN backward
struct NormBwdDev {
  const float* dy; long long lddy;
  const float* x; long long ldx;
  const float* stats;
  const float* weight;
  const float* cond; long long ldc;
  const float* dres; long long lddres;
  float* dx; long long lddx;
  __nv_bfloat16* dx_bf16; long long lddxb;
  float* dweight; float* dbias;
  float* dcond; long long lddc; int dcond_accumulate;
  __nv_bfloat16* dcond_b16;   // optional bf16 copy of dcond (row pitch 2d)
  float* dweight2; float* dbias2;
  int M, d, kind, rows_per_cta;
};

constexpr int kNormMaxChunks = 16;

struct NormBwdGroup { NormBwdDev it[SEA_MAX_STREAMS]; };  // the V field streams share a launch (blockIdx.y)

// WPR warps share a row (each owns CH chunks of 128 columns): wide rows (d = 2048) would otherwise need 128
// registers of row state per thread and 128 KB of per-warp partial sums, i.e. 8 warps per SM; with WPR = 2 the
// kernel fits twice per SM.  The row sums of the WPR warps meet in shared memory behind a 64-thread named barrier.
template <int CH, int WPR>
__global__ void __launch_bounds__(256, WPR) norm_bwd_kernel(const __grid_constant__ NormBwdGroup grp) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const NormBwdDev& a = grp.it[blockIdx.y];
  constexpr int SLOTS = 8 / WPR;  // rows in flight per CTA
  extern __shared__ float red[];  // [SLOTS][d] x2 (dweight, dbias partials; each lane owns its columns)
  __shared__ float xs[2][SLOTS][WPR][2];   // row-sum exchange, double-buffered by row parity
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = warp / WPR, part = warp % WPR;
  const int col0 = part * CH * 128;
  float* rw = red;
  float* rb = red + SLOTS * a.d;
  const bool want_param = (a.dweight != nullptr) || (a.dbias != nullptr) || (a.dweight2 != nullptr) || (a.dbias2 != nullptr);
  if (want_param) {
    for (int i = threadIdx.x; i < 2 * SLOTS * a.d; i += 256) red[i] = 0.f;
    __syncthreads();
  }
  const int row_begin = blockIdx.x * a.rows_per_cta;
  const int row_end = min(a.M, row_begin + a.rows_per_cta);
  const float inv_d = 1.0f / a.d;
  int it = 0;
  for (int m = row_begin + slot; m < row_end; m += SLOTS, ++it) {
    const float mean = a.stats[2 * m], rstd = a.stats[2 * m + 1];
    const float* xr = a.x + static_cast<long long>(m) * a.ldx;
    const float* dyr = a.dy + static_cast<long long>(m) * a.lddy;
    float4 xh[CH], g[CH];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int col = col0 + c * 128 + lane * 4;
      if (col < a.d) {
        const float4 xv = *reinterpret_cast<const float4*>(xr + col);
        const float4 dv = *reinterpret_cast<const float4*>(dyr + col);
        float4 w = __ldg(reinterpret_cast<const float4*>(a.weight + col));
        if (a.kind == SEA_NORM_ADALN) {
          const float4 cw = *reinterpret_cast<const float4*>(a.cond + static_cast<long long>(m) * a.ldc + col);
          w.x += cw.x + 1.f; w.y += cw.y + 1.f; w.z += cw.z + 1.f; w.w += cw.w + 1.f;
        }
        float4 h;
        h.x = (xv.x - mean) * rstd; h.y = (xv.y - mean) * rstd;
        h.z = (xv.z - mean) * rstd; h.w = (xv.w - mean) * rstd;
        float4 gg;
        gg.x = dv.x * w.x; gg.y = dv.y * w.y; gg.z = dv.z * w.z; gg.w = dv.w * w.w;
        xh[c] = h; g[c] = gg;
        s1 += gg.x + gg.y + gg.z + gg.w;
        s2 += gg.x * h.x + gg.y * h.y + gg.z * h.z + gg.w * h.w;
        // d gamma_eff = dy * xhat, d beta_eff = dy
        const float4 dgam = make_float4(dv.x * h.x, dv.y * h.y, dv.z * h.z, dv.w * h.w);
        if (want_param) {
          float4* pw = reinterpret_cast<float4*>(rw + slot * a.d + col);
          float4* pb = reinterpret_cast<float4*>(rb + slot * a.d + col);
          float4 t = *pw;
          t.x += dgam.x; t.y += dgam.y; t.z += dgam.z; t.w += dgam.w;
          *pw = t;
          t = *pb;
          t.x += dv.x; t.y += dv.y; t.z += dv.z; t.w += dv.w;
          *pb = t;
        }
        if (a.dcond) {
          float* dc = a.dcond + static_cast<long long>(m) * a.lddc;
          float4 o1 = dgam, o2 = dv;
          if (a.dcond_accumulate) {
            const float4 p1 = *reinterpret_cast<const float4*>(dc + col);
            const float4 p2 = *reinterpret_cast<const float4*>(dc + a.d + col);
            o1.x += p1.x; o1.y += p1.y; o1.z += p1.z; o1.w += p1.w;
            o2.x += p2.x; o2.y += p2.y; o2.z += p2.z; o2.w += p2.w;
          }
          *reinterpret_cast<float4*>(dc + col) = o1;
          *reinterpret_cast<float4*>(dc + a.d + col) = o2;
        }
        if (a.dcond_b16) {   // (scale | shift) gradient as the bf16 operand of the cond_mlp[2] backward GEMMs
          __nv_bfloat16* dcb = a.dcond_b16 + static_cast<long long>(m) * (2LL * a.d);
          uint2 q1, q2;
          q1.x = ptx::pack_bf16(dgam.x, dgam.y); q1.y = ptx::pack_bf16(dgam.z, dgam.w);
          q2.x = ptx::pack_bf16(dv.x, dv.y); q2.y = ptx::pack_bf16(dv.z, dv.w);
          *reinterpret_cast<uint2*>(dcb + col) = q1;
          *reinterpret_cast<uint2*>(dcb + a.d + col) = q2;
        }
      }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (WPR > 1) {
      if (lane == 0) { xs[it & 1][slot][part][0] = s1; xs[it & 1][slot][part][1] = s2; }
      asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "r"(32 * WPR) : "memory");
      s1 = 0.f; s2 = 0.f;
#pragma unroll
      for (int w = 0; w < WPR; ++w) { s1 += xs[it & 1][slot][w][0]; s2 += xs[it & 1][slot][w][1]; }
    }
    const float c1 = s1 * inv_d, c2 = s2 * inv_d;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int col = col0 + c * 128 + lane * 4;
      if (col < a.d) {
        float4 o;
        o.x = rstd * (g[c].x - c1 - xh[c].x * c2);
        o.y = rstd * (g[c].y - c1 - xh[c].y * c2);
        o.z = rstd * (g[c].z - c1 - xh[c].z * c2);
        o.w = rstd * (g[c].w - c1 - xh[c].w * c2);
        if (a.dres) {
          const float4 r = *reinterpret_cast<const float4*>(a.dres + static_cast<long long>(m) * a.lddres + col);
          o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        if (a.dx) *reinterpret_cast<float4*>(a.dx + static_cast<long long>(m) * a.lddx + col) = o;
        if (a.dx_bf16) {
          uint2 pk;
          pk.x = ptx::pack_bf16(o.x, o.y);
          pk.y = ptx::pack_bf16(o.z, o.w);
          *reinterpret_cast<uint2*>(a.dx_bf16 + static_cast<long long>(m) * a.lddxb + col) = pk;
        }
      }
    }
  }
  // cross-warp reduction of the column partials, then one atomic per column per CTA
  if (!want_param) return;
  __syncthreads();
  for (int col = threadIdx.x; col < a.d; col += blockDim.x) {
    float sw = 0.f, sb = 0.f;
#pragma unroll
    for (int w = 0; w < SLOTS; ++w) {
      sw += rw[w * a.d + col];
      sb += rb[w * a.d + col];
    }
    if (a.dweight) atomicAdd(a.dweight + col, sw);
    if (a.dbias) atomicAdd(a.dbias + col, sb);
    if (a.dweight2) atomicAdd(a.dweight2 + col, sw);
    if (a.dbias2) atomicAdd(a.dbias2 + col, sb);
  }
}

// ------------------------------------------------------------- LayerNorm(H)+GELU backward (K6)
// dh = LN'(GELU'(u) * dg);  dweight/dbias of the inner nn.LayerNorm accumulate in shared memory.
struct LnGeluBwdItem {
  const __nv_bfloat16 *dg, *h; const float *stats, *weight, *bias; __nv_bfloat16* dh; float *dweight, *dbias;
};
struct LnGeluBwdGroup { LnGeluBwdItem it[SEA_MAX_STREAMS]; };

// 512 threads per CTA, one row at a time, thread t owns columns (c*512 + t)*8 .. +8 of every row (c < 4: H <= 16384).
// The next row's h / dg are already in flight (registers) while the current row is being processed, x-hat is
// recomputed from the raw bf16 h in the second pass instead of being kept, and the per-CTA dweight / dbias
// partial sums live in shared memory (each thread is the only writer of its columns).
constexpr int kLgbThreads = 512;
constexpr int kLgbChunks = 4;

template <int kDummy>
__global__ void __launch_bounds__(kLgbThreads, 1) ln_gelu_bwd_kernel(const __grid_constant__ LnGeluBwdGroup grp, long long lddg,
                                                                      long long ldh, long long lddh, int M, int H,
                                                                      int rows_per_cta) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const LnGeluBwdItem& item = grp.it[blockIdx.y];
  const __nv_bfloat16* __restrict__ dg = item.dg;
  const __nv_bfloat16* __restrict__ h = item.h;
  const float* __restrict__ stats = item.stats;
  const float* __restrict__ weight = item.weight;
  const float* __restrict__ bias = item.bias;
  __nv_bfloat16* __restrict__ dh = item.dh;
  float* __restrict__ dweight = item.dweight;
  float* __restrict__ dbias = item.dbias;
  extern __shared__ float acc[];  // [2][H]
  __shared__ float red[2][kLgbThreads / 32];
  __shared__ float bc[2];
  const int tid = threadIdx.x;
  float* aw = acc;
  float* ab = acc + H;
  for (int i = tid; i < 2 * H; i += kLgbThreads) acc[i] = 0.f;
  __syncthreads();
  const int row_begin = blockIdx.x * rows_per_cta;
  const int row_end = min(M, row_begin + rows_per_cta);
  const float inv_h = 1.0f / H;
  uint4 nh[kLgbChunks], ng[kLgbChunks];   // next row, in flight
  float2 nst = make_float2(0.f, 1.f);
  auto fetch = [&](int m) {
    nst = *reinterpret_cast<const float2*>(stats + 2 * m);
#pragma unroll
    for (int c = 0; c < kLgbChunks; ++c) {
      const int col = (c * kLgbThreads + tid) * 8;
      if (col < H) {
        nh[c] = *reinterpret_cast<const uint4*>(h + static_cast<long long>(m) * ldh + col);
        ng[c] = *reinterpret_cast<const uint4*>(dg + static_cast<long long>(m) * lddg + col);
      }
    }
  };
  if (row_begin < row_end) fetch(row_begin);
  for (int m = row_begin; m < row_end; ++m) {
    const float mean = nst.x, rstd = nst.y;
    uint4 ch[kLgbChunks];
    float dhh[kLgbChunks][8];
    float s1 = 0.f, s2 = 0.f;
    uint4 cg[kLgbChunks];
#pragma unroll
    for (int c = 0; c < kLgbChunks; ++c) { ch[c] = nh[c]; cg[c] = ng[c]; }
    if (m + 1 < row_end) fetch(m + 1);
#pragma unroll
    for (int c = 0; c < kLgbChunks; ++c) {
      const int col = (c * kLgbThreads + tid) * 8;
      if (col < H) {
        const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&ch[c]);
        const __nv_bfloat162* gp = reinterpret_cast<const __nv_bfloat162*>(&cg[c]);
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(weight + col));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(weight + col + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + col));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + col + 4));
        const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float daw[8], dab[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 hv = __bfloat1622float2(hp[q]);
          const float2 gv = __bfloat1622float2(gp[q]);
          const float hv2[2] = {hv.x, hv.y}, gv2[2] = {gv.x, gv.y};
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int k = 2 * q + e;
            const float xhat = (hv2[e] - mean) * rstd;
            const float u = xhat * ww[k] + bb[k];
            const float dgu = gv2[e] * ptx::gelu_erf_grad_fast(u);
            daw[k] = dgu * xhat;
            dab[k] = dgu;
            const float dxh = dgu * ww[k];
            dhh[c][k] = dxh;
            s1 += dxh;
            s2 += dxh * xhat;
          }
        }
        // partial sums: this thread is the only writer of its 8 columns.  Layout [plane][c*512 + tid] of float4
        // (plane = low / high four columns): consecutive threads touch consecutive 16-byte words, so the
        // read-modify-write is conflict-free (column-major [H] floats were an 8-way bank conflict and the
        // whole kernel's bottleneck)
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) {
          float4* pw = reinterpret_cast<float4*>(aw) + pl * (H / 8) + c * kLgbThreads + tid;
          float4* pb = reinterpret_cast<float4*>(ab) + pl * (H / 8) + c * kLgbThreads + tid;
          float4 vw = *pw, vb = *pb;
          vw.x += daw[4 * pl]; vw.y += daw[4 * pl + 1]; vw.z += daw[4 * pl + 2]; vw.w += daw[4 * pl + 3];
          vb.x += dab[4 * pl]; vb.y += dab[4 * pl + 1]; vb.z += dab[4 * pl + 2]; vb.w += dab[4 * pl + 3];
          *pw = vw; *pb = vb;
        }
      }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if ((tid & 31) == 0) { red[0][tid >> 5] = s1; red[1][tid >> 5] = s2; }
    __syncthreads();
    if (tid < 32) {
      float t1 = tid < kLgbThreads / 32 ? red[0][tid] : 0.f, t2 = tid < kLgbThreads / 32 ? red[1][tid] : 0.f;
      t1 = warp_sum(t1); t2 = warp_sum(t2);
      if (tid == 0) { bc[0] = t1 * inv_h; bc[1] = t2 * inv_h; }
    }
    __syncthreads();
    const float c1 = bc[0], c2 = bc[1];
#pragma unroll
    for (int c = 0; c < kLgbChunks; ++c) {
      const int col = (c * kLgbThreads + tid) * 8;
      if (col < H) {
        const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&ch[c]);
        float o[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 hv = __bfloat1622float2(hp[q]);
          o[2 * q] = rstd * (dhh[c][2 * q] - c1 - (hv.x - mean) * rstd * c2);
          o[2 * q + 1] = rstd * (dhh[c][2 * q + 1] - c1 - (hv.y - mean) * rstd * c2);
        }
        uint4 pk;
        pk.x = ptx::pack_bf16(o[0], o[1]); pk.y = ptx::pack_bf16(o[2], o[3]);
        pk.z = ptx::pack_bf16(o[4], o[5]); pk.w = ptx::pack_bf16(o[6], o[7]);
        *reinterpret_cast<uint4*>(dh + static_cast<long long>(m) * lddh + col) = pk;
      }
    }
    // red[] / bc[] are rewritten only after the next row's first barrier has been passed by every thread
    // that read them here; one barrier closes the hazard on bc[] (written by thread 0 after barrier 1)
    __syncthreads();
  }
  // every thread flushes the words it accumulated itself (no barrier needed): word (plane, c, tid) holds
  // columns (c*512 + tid)*8 + plane*4 .. +4
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(dweight) | reinterpret_cast<uintptr_t>(dbias)) & 15) == 0;
#pragma unroll
  for (int c = 0; c < kLgbChunks; ++c) {
    const int col = (c * kLgbThreads + tid) * 8;
    if (col < H) {
#pragma unroll
      for (int pl = 0; pl < 2; ++pl) {
        const float4 vw = reinterpret_cast<const float4*>(aw)[pl * (H / 8) + c * kLgbThreads + tid];
        const float4 vb = reinterpret_cast<const float4*>(ab)[pl * (H / 8) + c * kLgbThreads + tid];
        const int i = col + 4 * pl;
        if (vec_ok) {
          ptx::red_add_v4(dweight + i, vw.x, vw.y, vw.z, vw.w);
          ptx::red_add_v4(dbias + i, vb.x, vb.y, vb.z, vb.w);
        } else {
          atomicAdd(dweight + i, vw.x); atomicAdd(dweight + i + 1, vw.y);
          atomicAdd(dweight + i + 2, vw.z); atomicAdd(dweight + i + 3, vw.w);
          atomicAdd(dbias + i, vb.x); atomicAdd(dbias + i + 1, vb.y);
          atomicAdd(dbias + i + 2, vb.z); atomicAdd(dbias + i + 3, vb.w);
        }
      }
    }
  }
}

// --------------------------------------------------------- AdaLN cond_mlp.0 + SiLU backward
// dpre = dh * silu'(w1*ib + b1);  dw1[j,c] += sum_m dpre*ib[m,c];  db1[j] += sum_m dpre
__global__ void __launch_bounds__(256) adaln_hidden_bwd_kernel(const float* __restrict__ dh, long long lddh,
                                                               const float* __restrict__ ib, int M, int ib_num,
                                                               const float* __restrict__ w1,
                                                               const float* __restrict__ b1, int n,
                                                               float* __restrict__ dw1, float* __restrict__ db1,
                                                               int rows_per_cta) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  float wj[4], aw[4] = {0.f, 0.f, 0.f, 0.f}, ab = 0.f;
  for (int c = 0; c < ib_num && c < 4; ++c) wj[c] = w1[j * ib_num + c];
  const float bj = b1[j];
  for (int m = r0; m < r1; ++m) {
    float pre = bj;
    for (int c = 0; c < ib_num && c < 4; ++c) pre = fmaf(wj[c], ib[static_cast<long long>(m) * ib_num + c], pre);
    const float sig = __fdividef(1.0f, 1.0f + __expf(-pre));
    const float dsilu = sig * (1.0f + pre * (1.0f - sig));
    const float dp = dh[static_cast<long long>(m) * lddh + j] * dsilu;
    ab += dp;
    for (int c = 0; c < ib_num && c < 4; ++c) aw[c] = fmaf(dp, ib[static_cast<long long>(m) * ib_num + c], aw[c]);
  }
  atomicAdd(db1 + j, ab);
  for (int c = 0; c < ib_num && c < 4; ++c) atomicAdd(dw1 + j * ib_num + c, aw[c]);
}

// ------------------------------------------------------------------------------ TIPI backward
// (A) dW3[n,k] += sum_m dx[m,n] g[m,k];  db3[n] += sum_m dx[m,n]        (thread per column n)
__global__ void __launch_bounds__(256) tipi_bwd_w_kernel(const float* __restrict__ dx, long long lddx,
                                                         const float* __restrict__ g, int M, int E, int hid,
                                                         float* __restrict__ dw3, float* __restrict__ db3,
                                                         int rows_per_cta) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= E) return;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  float aw[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, ab = 0.f;
  for (int m = r0; m < r1; ++m) {
    const float v = dx[static_cast<long long>(m) * lddx + n];
    ab += v;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < hid) aw[k] = fmaf(v, g[static_cast<long long>(m) * hid + k], aw[k]);
  }
  atomicAdd(db3 + n, ab);
  for (int k = 0; k < hid && k < 8; ++k) atomicAdd(dw3 + n * hid + k, aw[k]);
}

// (B) dg[m,k] = sum_streams sum_n dx_s[m,n] W3[n,k]; then GELU'/LayerNorm' over hid (<= 8) and
//     the gradients of Linear(ib_num, hid) + LayerNorm(hid).  One warp per row; `u` is the saved
//     pre-LayerNorm activation, `stats` its (mean, rstd).
struct TipiBwdDev {
  const float* dx[SEA_MAX_STREAMS]; long long lddx; int n_streams;
  const float* w3; const float* u; const float* stats; const float* ib;
  const float* ln_w; const float* ln_b;
  float *dw0, *db0, *dlnw, *dlnb;
  int M, E, hid, ib_num;
};
__global__ void __launch_bounds__(256) tipi_bwd_g_kernel(const TipiBwdDev a) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  // CTA-level accumulators: db0[8] | dlnw[8] | dlnb[8] | dw0[8][4]
  __shared__ float acc[24 + 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * 8 + warp;
  if (threadIdx.x < 56) acc[threadIdx.x] = 0.f;
  __syncthreads();
  if (m < a.M) {
    float dg[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int s = 0; s < a.n_streams; ++s) {
      const float* dxr = a.dx[s] + static_cast<long long>(m) * a.lddx;
      if (a.hid == 8) {
        // W3 row n = 8 consecutive floats: two 16-byte loads instead of eight 4-byte loads with a 32-byte lane stride
#pragma unroll 4
        for (int n = lane; n < a.E; n += 32) {
          const float v = dxr[n];
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(a.w3 + n * 8));
          const float4 w1 = __ldg(reinterpret_cast<const float4*>(a.w3 + n * 8 + 4));
          dg[0] = fmaf(v, w0.x, dg[0]); dg[1] = fmaf(v, w0.y, dg[1]); dg[2] = fmaf(v, w0.z, dg[2]); dg[3] = fmaf(v, w0.w, dg[3]);
          dg[4] = fmaf(v, w1.x, dg[4]); dg[5] = fmaf(v, w1.y, dg[5]); dg[6] = fmaf(v, w1.z, dg[6]); dg[7] = fmaf(v, w1.w, dg[7]);
        }
      } else {
        for (int n = lane; n < a.E; n += 32) {
          const float v = dxr[n];
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k < a.hid) dg[k] = fmaf(v, __ldg(a.w3 + n * a.hid + k), dg[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) dg[k] = warp_sum(dg[k]);
    if (lane == 0) {
      const float mean = a.stats[2 * m], rstd = a.stats[2 * m + 1];
      float xh[8], dxh[8], s1 = 0.f, s2 = 0.f;
      for (int k = 0; k < a.hid; ++k) {
        xh[k] = (a.u[static_cast<long long>(m) * a.hid + k] - mean) * rstd;
        const float nrm = xh[k] * a.ln_w[k] + a.ln_b[k];
        const float dgu = dg[k] * ptx::gelu_erf_grad(nrm);
        atomicAdd(&acc[8 + k], dgu * xh[k]);
        atomicAdd(&acc[16 + k], dgu);
        dxh[k] = dgu * a.ln_w[k];
        s1 += dxh[k];
        s2 += dxh[k] * xh[k];
      }
      const float c1 = s1 / a.hid, c2 = s2 / a.hid;
      for (int k = 0; k < a.hid; ++k) {
        const float du = rstd * (dxh[k] - c1 - xh[k] * c2);
        atomicAdd(&acc[k], du);
        for (int c = 0; c < a.ib_num && c < 4; ++c)
          atomicAdd(&acc[24 + k * 4 + c], du * a.ib[static_cast<long long>(m) * a.ib_num + c]);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < a.hid) {
    const int k = threadIdx.x;
    atomicAdd(a.db0 + k, acc[k]);
    atomicAdd(a.dlnw + k, acc[8 + k]);
    atomicAdd(a.dlnb + k, acc[16 + k]);
    for (int c = 0; c < a.ib_num && c < 4; ++c) atomicAdd(a.dw0 + k * a.ib_num + c, acc[24 + k * 4 + c]);
  }
}

// ---------------------------------------------------------- fp32-parity mode (precision="fp32")
// The reference trains in fp32 (train/train_temporal.py:252-258, no AMP).  These are the accurate, plain
// counterparts of the bf16 kernels above: fp32 I/O, erff / expf (no MUFU approximations); speed is secondary.
__global__ void __launch_bounds__(256) ln_gelu_bwd_f32_kernel(const float* __restrict__ dg, long long lddg,
                                                              const float* __restrict__ h, long long ldh,
                                                              const float* __restrict__ stats,
                                                              const float* __restrict__ weight,
                                                              const float* __restrict__ bias, float* __restrict__ dh,
                                                              long long lddh, float* __restrict__ dweight,
                                                              float* __restrict__ dbias, int M, int H, int rows_per_cta) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  extern __shared__ float acc32[];   // [2][H]; column c is only ever touched by thread c % 256
  __shared__ float red[2][8];
  __shared__ float bc[2];
  const int tid = threadIdx.x;
  for (int i = tid; i < 2 * H; i += 256) acc32[i] = 0.f;
  const int row_begin = blockIdx.x * rows_per_cta, row_end = min(M, row_begin + rows_per_cta);
  const float inv_h = 1.0f / H;
  for (int m = row_begin; m < row_end; ++m) {
    const float mean = stats[2 * m], rstd = stats[2 * m + 1];
    const float* hr = h + static_cast<long long>(m) * ldh;
    const float* gr = dg + static_cast<long long>(m) * lddg;
    float s1 = 0.f, s2 = 0.f;
    for (int c = tid; c < H; c += 256) {
      const float xhat = (hr[c] - mean) * rstd;
      const float dgu = gr[c] * ptx::gelu_erf_grad(xhat * weight[c] + bias[c]);
      acc32[c] += dgu * xhat;
      acc32[H + c] += dgu;
      const float dxh = dgu * weight[c];
      s1 += dxh;
      s2 += dxh * xhat;
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if ((tid & 31) == 0) { red[0][tid >> 5] = s1; red[1][tid >> 5] = s2; }
    __syncthreads();
    if (tid < 32) {
      float t1 = tid < 8 ? red[0][tid] : 0.f, t2 = tid < 8 ? red[1][tid] : 0.f;
      t1 = warp_sum(t1); t2 = warp_sum(t2);
      if (tid == 0) { bc[0] = t1 * inv_h; bc[1] = t2 * inv_h; }
    }
    __syncthreads();
    const float c1 = bc[0], c2 = bc[1];
    float* dr = dh + static_cast<long long>(m) * lddh;
    for (int c = tid; c < H; c += 256) {
      const float xhat = (hr[c] - mean) * rstd;
      const float dxh = gr[c] * ptx::gelu_erf_grad(xhat * weight[c] + bias[c]) * weight[c];
      dr[c] = rstd * (dxh - c1 - xhat * c2);
    }
    __syncthreads();
  }
  for (int c = tid; c < H; c += 256) {
    atomicAdd(dweight + c, acc32[c]);
    atomicAdd(dbias + c, acc32[H + c]);
  }
}

// d[m, n] *= gelu'(pre[m, n])   (the exchange branch: cross_up(GELU(projection(.))), models/temporal.py:185)
__global__ void __launch_bounds__(256) gelu_grad_mul_f32_kernel(float* __restrict__ d, long long ldd,
                                                                const float* __restrict__ pre, long long ldp, int M, int N) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= static_cast<long long>(M) * N) return;
  const int m = static_cast<int>(i / N), n = static_cast<int>(i - static_cast<long long>(m) * N);
  d[m * ldd + n] *= ptx::gelu_erf_grad(pre[m * ldp + n]);
}

}  // namespace
}  // namespace sea

using namespace sea;

static int rows_per_cta_for(int M, int granule) {
  int r = (M + 2 * 148 - 1) / (2 * 148);
  r = ((r + granule - 1) / granule) * granule;
  return r < granule ? granule : r;
}

static int fill_norm_bwd(const sea_norm_bwd_args* a, NormBwdDev& d) {
  if (!a->dy || !a->x || !a->stats || !a->weight || a->M <= 0 || a->d <= 0) return SEA_ERR_INVALID;
  if ((a->d % 4) || a->d > kNormMaxChunks * 128) return SEA_ERR_UNSUPPORTED;
  if (a->kind == SEA_NORM_ADALN && !a->cond) return SEA_ERR_INVALID;
  if (!a->dx && !a->dx_bf16) return SEA_ERR_INVALID;
  if ((a->lddy % 4) || (a->ldx % 4) || (a->dx && (a->lddx % 4)) || (a->dres && (a->lddres % 4)))
    return SEA_ERR_INVALID;
  d.dy = a->dy; d.lddy = a->lddy; d.x = a->x; d.ldx = a->ldx; d.stats = a->stats;
  d.weight = a->weight; d.cond = a->cond; d.ldc = a->ldc;
  d.dres = a->dres; d.lddres = a->lddres;
  d.dx = a->dx; d.lddx = a->lddx;
  d.dx_bf16 = static_cast<__nv_bfloat16*>(a->dx_bf16); d.lddxb = a->lddx_bf16;
  d.dweight = a->dweight; d.dbias = a->dbias;
  d.dcond = a->dcond; d.lddc = a->lddcond; d.dcond_accumulate = a->dcond_accumulate;
  d.dcond_b16 = a->kind == SEA_NORM_ADALN ? static_cast<__nv_bfloat16*>(a->dcond_bf16) : nullptr;
  if (d.dcond_b16 && a->dcond_accumulate) return SEA_ERR_INVALID;
  d.dweight2 = a->dweight2; d.dbias2 = a->dbias2;
  d.M = a->M; d.d = a->d; d.kind = a->kind;
  d.rows_per_cta = rows_per_cta_for(a->M, 8);
  return SEA_OK;
}

extern "C" int sea_norm_bwd_group(int n, const sea_norm_bwd_args* a, sea_stream_t stream) {
  if (!a || n < 1 || n > SEA_MAX_STREAMS) return SEA_ERR_INVALID;
  NormBwdGroup g;
  for (int i = 0; i < n; ++i) {
    int rc = fill_norm_bwd(&a[i], g.it[i]);
    if (rc) return rc;
    if (a[i].M != a->M || a[i].d != a->d || a[i].kind != a->kind) return SEA_ERR_INVALID;
  }
  const NormBwdDev& d = g.it[0];
  const dim3 grid((a->M + d.rows_per_cta - 1) / d.rows_per_cta, n);
  const bool wide = a->d > 1024;
  const size_t smem = sizeof(float) * (wide ? 8 : 16) * a->d;
  static bool attr_set[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 16 && !attr_set[dev]) {
    SEA_CUDA_OK(cudaFuncSetAttribute((norm_bwd_kernel<4, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 2048 * 4));
    SEA_CUDA_OK(cudaFuncSetAttribute((norm_bwd_kernel<8, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 2048 * 4));
    SEA_CUDA_OK(cudaFuncSetAttribute((norm_bwd_kernel<8, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2048 * 4));
    attr_set[dev] = true;
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (a->d <= 512) SEA_LAUNCH((norm_bwd_kernel<4, 1>), grid, 256, smem, s, g);
  else if (a->d <= 1024) SEA_LAUNCH((norm_bwd_kernel<8, 1>), grid, 256, smem, s, g);
  else SEA_LAUNCH((norm_bwd_kernel<8, 2>), grid, 256, smem, s, g);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int sea_norm_bwd(const sea_norm_bwd_args* a, sea_stream_t stream) {
  return sea_norm_bwd_group(1, a, stream);
}

extern "C" int sea_ln_gelu_bwd_group(int n, const sea_ln_gelu_bwd_args* a, sea_stream_t stream) {
  if (!a || n < 1 || n > SEA_MAX_STREAMS) return SEA_ERR_INVALID;
  if (a->prec == SEA_PREC_FP32) {   // parity mode: fp32 dg / h / dh, one plain launch per item
    for (int i = 0; i < n; ++i) {
      const sea_ln_gelu_bwd_args* x = &a[i];
      if (!x->dg || !x->h || !x->stats || !x->weight || !x->bias || !x->dh || !x->dweight || !x->dbias || x->prec != SEA_PREC_FP32)
        return SEA_ERR_INVALID;
      if (x->M <= 0 || x->H <= 0 || x->H > 24576) return SEA_ERR_UNSUPPORTED;
      int rows = (x->M + 147) / 148;
      const size_t smem = sizeof(float) * 2 * x->H;
      static bool attr32[16] = {};
      int dev = 0;
      cudaGetDevice(&dev);
      if (dev < 16 && !attr32[dev]) {
        SEA_CUDA_OK(cudaFuncSetAttribute(ln_gelu_bwd_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 24576 * 4));
        attr32[dev] = true;
      }
      SEA_LAUNCH(ln_gelu_bwd_f32_kernel, (x->M + rows - 1) / rows, 256, smem, reinterpret_cast<cudaStream_t>(stream),
                 static_cast<const float*>(x->dg), static_cast<long long>(x->lddg), static_cast<const float*>(x->h),
                 static_cast<long long>(x->ldh), x->stats, x->weight, x->bias, static_cast<float*>(x->dh),
                 static_cast<long long>(x->lddh), x->dweight, x->dbias, x->M, x->H, rows);
    }
    return static_cast<int>(cudaGetLastError());
  }
  LnGeluBwdGroup g;
  for (int i = 0; i < n; ++i) {
    const sea_ln_gelu_bwd_args* x = &a[i];
    if (!x->dg || !x->h || !x->stats || !x->weight || !x->bias || !x->dh || !x->dweight || !x->dbias)
      return SEA_ERR_INVALID;
    if (x->M <= 0 || x->H <= 0 || (x->H % 8) || x->H > 16384) return SEA_ERR_UNSUPPORTED;
    if ((x->lddg % 8) || (x->ldh % 8) || (x->lddh % 8)) return SEA_ERR_INVALID;
    if (x->M != a->M || x->H != a->H || x->lddg != a->lddg || x->ldh != a->ldh || x->lddh != a->lddh) return SEA_ERR_INVALID;
    g.it[i] = LnGeluBwdItem{static_cast<const __nv_bfloat16*>(x->dg), static_cast<const __nv_bfloat16*>(x->h), x->stats,
                            x->weight, x->bias, static_cast<__nv_bfloat16*>(x->dh), x->dweight, x->dbias};
  }
  // one wave of CTAs (each keeps [2][H] partial sums in shared memory and ends with 2H global reductions)
  int rows = (a->M * n + 147) / 148;
  if (rows < 1) rows = 1;
  const dim3 grid((a->M + rows - 1) / rows, n);
  const size_t smem = sizeof(float) * 2 * a->H;
  static bool attr_set[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 16 && !attr_set[dev]) {
    SEA_CUDA_OK(cudaFuncSetAttribute(ln_gelu_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 16384 * 4));
    attr_set[dev] = true;
  }
  SEA_LAUNCH((ln_gelu_bwd_kernel<0>), grid, kLgbThreads, smem, reinterpret_cast<cudaStream_t>(stream), g, a->lddg, a->ldh,
             a->lddh, a->M, a->H, rows);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int sea_gelu_grad_mul_f32(float* d, int64_t ldd, const float* pre, int64_t ldp, int M, int N, sea_stream_t stream) {
  if (!d || !pre || M <= 0 || N <= 0) return SEA_ERR_INVALID;
  const long long work = static_cast<long long>(M) * N;
  SEA_LAUNCH(gelu_grad_mul_f32_kernel, static_cast<unsigned>((work + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream),
             d, static_cast<long long>(ldd), pre, static_cast<long long>(ldp), M, N);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int sea_ln_gelu_bwd(const sea_ln_gelu_bwd_args* a, sea_stream_t stream) {
  return sea_ln_gelu_bwd_group(1, a, stream);
}

extern "C" int sea_adaln_hidden_bwd(const float* dh, int64_t lddh, const float* ib, int M, int ib_num,
                                    const float* w1, const float* b1, int n, float* dw1, float* db1,
                                    sea_stream_t stream) {
  if (!dh || !ib || !w1 || !b1 || !dw1 || !db1 || M <= 0 || n <= 0) return SEA_ERR_INVALID;
  if (ib_num < 1 || ib_num > 4) return SEA_ERR_UNSUPPORTED;
  const int rows = rows_per_cta_for(M, 8) * 4;
  dim3 grid((n + 255) / 256, (M + rows - 1) / rows);
  SEA_LAUNCH(adaln_hidden_bwd_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream), dh, lddh, ib, M, ib_num, w1, b1, n, dw1, db1, rows);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int sea_tipi_bwd(const sea_tipi_bwd_args* a, sea_stream_t stream) {
  if (!a || a->n_streams < 1 || a->n_streams > SEA_MAX_STREAMS || a->M <= 0 || a->E <= 0) return SEA_ERR_INVALID;
  if (a->hid < 1 || a->hid > 8 || a->ib_num < 1 || a->ib_num > 4) return SEA_ERR_UNSUPPORTED;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int rows = rows_per_cta_for(a->M, 8) * 4;
  dim3 grid((a->E + 255) / 256, (a->M + rows - 1) / rows);
  TipiBwdDev d{};
  for (int i = 0; i < a->n_streams; ++i) {
    if (!a->dx[i]) return SEA_ERR_INVALID;
    d.dx[i] = a->dx[i];
    SEA_LAUNCH(tipi_bwd_w_kernel, grid, 256, 0, s, a->dx[i], a->lddx, a->g, a->M, a->E, a->hid, a->dw3, a->db3, rows);
  }
  d.lddx = a->lddx; d.n_streams = a->n_streams;
  d.w3 = a->w3; d.u = a->u; d.stats = a->stats; d.ib = a->ib; d.ln_w = a->ln_w; d.ln_b = a->ln_b;
  d.dw0 = a->dw0; d.db0 = a->db0; d.dlnw = a->dlnw; d.dlnb = a->dlnb;
  d.M = a->M; d.E = a->E; d.hid = a->hid; d.ib_num = a->ib_num;
  SEA_LAUNCH(tipi_bwd_g_kernel, (a->M + 7) / 8, 256, 0, s, d);
  return static_cast<int>(cudaGetLastError());
}

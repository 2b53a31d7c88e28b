// Pipe-throughput probe for the attention softmax inner loop (B200): one warp per SM sub-partition,
// clock64 around 64-element register loops.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 sfu_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float a, float b) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float fl = x + 12582912.0f;
  const float f = x - (fl - 12582912.0f);
  float p = fmaf(0.0551714599f, f, 0.2426108569f);
  p = fmaf(p, f, 0.6932609677f);
  p = fmaf(p, f, 0.9999281168f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(fl) << 23));
}

template <int MODE>
__global__ void probe(const float* in, float* out, long long* cyc, float scale, float nm) {
  float r[64];
  for (int e = 0; e < 64; ++e) r[e] = in[threadIdx.x * 64 + e];
  float l0 = 0.f, l1 = 0.f;
  uint32_t pk[32];
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 16; ++it) {
#pragma unroll
    for (int e = 0; e < 64; e += 2) {
      float p0, p1;
      if (MODE == 0) { p0 = ex2(r[e]); p1 = ex2(r[e + 1]); }                                   // MUFU only
      if (MODE == 1) { p0 = ex2(fmaf(r[e], scale, nm)); p1 = ex2(fmaf(r[e + 1], scale, nm)); } // + FFMA
      if (MODE >= 2 && MODE != 7) { p0 = ex2(fmaf(r[e], scale, nm)); p1 = ex2(fmaf(r[e + 1], scale, nm)); l0 += p0; l1 += p1; }
      if (MODE == 4 || MODE == 5) {   // 3 of 8 (MODE 4) or 4 of 8 (MODE 5) on the FMA pipe
        const uint32_t mask = MODE == 4 ? 0x49u : 0x55u;
        const float x0 = fmaf(r[e], scale, nm), x1 = fmaf(r[e + 1], scale, nm);
        p0 = ((mask >> (e & 7)) & 1) ? ex2_poly(x0) : ex2(x0);
        p1 = ((mask >> ((e + 1) & 7)) & 1) ? ex2_poly(x1) : ex2(x1);
        l0 += p0; l1 += p1;
      }
      if (MODE == 6) { p0 = ex2_poly(fmaf(r[e], scale, nm)); p1 = ex2_poly(fmaf(r[e + 1], scale, nm)); l0 += p0; l1 += p1; }
      if (MODE == 7) {   // packed fma.rn.f32x2 / add.f32x2 (two scores per FMA-pipe instruction)
        unsigned long long xx, rr, ss, nn, ll;
        asm("mov.b64 %0, {%1, %2};" : "=l"(rr) : "f"(r[e]), "f"(r[e + 1]));
        asm("mov.b64 %0, {%1, %1};" : "=l"(ss) : "f"(scale));
        asm("mov.b64 %0, {%1, %1};" : "=l"(nn) : "f"(nm));
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(xx) : "l"(rr), "l"(ss), "l"(nn));
        float x0, x1;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(xx));
        p0 = ex2(x0); p1 = ex2(x1);
        asm("mov.b64 %0, {%1, %2};" : "=l"(xx) : "f"(p0), "f"(p1));
        asm("mov.b64 %0, {%1, %2};" : "=l"(ll) : "f"(l0), "f"(l1));
        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(ll) : "l"(ll), "l"(xx));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(l0), "=f"(l1) : "l"(ll));
      }
      if (MODE >= 3) pk[e >> 1] = pack(p0, p1); else { pk[e >> 1] = __float_as_uint(p0) ^ __float_as_uint(p1); }
    }
#pragma unroll
    for (int e = 0; e < 32; ++e) r[e] = __uint_as_float((pk[e] & 0x007fffffu) | 0xbf000000u);   // feed back, keeps values sane
    nm += l0 * 1e-30f;
  }
  const long long t1 = clock64();
  float acc = l0 + l1;
  for (int e = 0; e < 64; ++e) acc += r[e];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads, const float* in, float* out, long long* cyc) {
  probe<MODE><<<1, threads>>>(in, out, cyc, 0.125f, -1.0f);
  probe<MODE><<<1, threads>>>(in, out, cyc, 0.125f, -1.0f);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-44s warps/SMSP=%d : %7.1f cycles per 64 elements (%.2f / element)\n", name, threads / 128, c / 16.0, c / 16.0 / 64);
}

int main() {
  float *in, *out; long long* cyc;
  cudaMalloc(&in, 1024 * 64 * 4); cudaMalloc(&out, 1024 * 4); cudaMalloc(&cyc, 64);
  cudaMemset(in, 0, 1024 * 64 * 4);
  for (int threads : {128, 256, 512}) {
    run<0>("MUFU.EX2 only", threads, in, out, cyc);
    run<1>("FFMA + MUFU", threads, in, out, cyc);
    run<2>("FFMA + MUFU + FADD", threads, in, out, cyc);
    run<3>("FFMA + MUFU + FADD + F2FP.BF16x2", threads, in, out, cyc);
    run<4>("same, 3 of 8 exponentials as FMA polynomial", threads, in, out, cyc);
    run<5>("same, 4 of 8 exponentials as FMA polynomial", threads, in, out, cyc);
    run<6>("same, all exponentials as FMA polynomial", threads, in, out, cyc);
    run<7>("FFMA2 + 2 MUFU + FADD2 + F2FP (packed f32x2)", threads, in, out, cyc);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

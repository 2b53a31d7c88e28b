// Mesh patchify / unpatch on the GPU: the reference's DataPartitioner2D
// (utils/data_processors.py:9-111) — bucketize the cell coordinates into an (m-1) x (n-1) grid of
// patches, list the cells of every patch in ascending cell order, pad every list to the longest one —
// as four kernels of HBM-bound integer / index work:
//   patch_bucketize   one thread per cell: torch.bucketize(right=True) against the (tiny) boundary
//                     vectors, clamp to [1, m-1] / [1, n-1], patch = (ix-1)*(n-1) + (iy-1); histogram;
//   patch_index_map   one CTA per patch: a stable block-wide compaction over all cells (ascending
//                     cell index, exactly mask.nonzero() of the reference), padded with pad_id;
//   patch_gather      out[s, p, c, f] = fields[f][s, index[p, c]]  (pad -> pad_field_value);
//                     optional [s, p, f, c] layout = the SpatialModel input (models/encoder_decoder.py:105-110);
//   patch_scatter     the inverse (inverse_partition, :90-111): recon[s, index[p, c], f] = part[s, p, c, f].
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sea_b200.h"
#include "internal.h"
#include "ptx.cuh"

namespace sea {
namespace {

// number of boundaries <= v  (torch.bucketize(v, b, right=True))
__device__ __forceinline__ int bucket_right(const float* __restrict__ b, int n, float v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (b[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256) patch_bucketize_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                              int n_cells, const float* __restrict__ xb, int m,
                                                              const float* __restrict__ yb, int n,
                                                              int32_t* __restrict__ patch_id, int32_t* __restrict__ counts) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  __shared__ float sxb[64], syb[64];
  for (int i = threadIdx.x; i < m; i += blockDim.x) sxb[i] = xb[i];
  for (int i = threadIdx.x; i < n; i += blockDim.x) syb[i] = yb[i];
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  int ix = bucket_right(sxb, m, x[c]);
  int iy = bucket_right(syb, n, y[c]);
  ix = min(max(ix, 1), m - 1);
  iy = min(max(iy, 1), n - 1);
  const int p = (ix - 1) * (n - 1) + (iy - 1);
  patch_id[c] = p;
  atomicAdd(counts + p, 1);
}

// DataPartitioner3D (utils/data_processors.py:114-165): patch = ((ix-1)*(n-1) + (iy-1))*(k-1) + (iz-1)
__global__ void __launch_bounds__(256) patch_bucketize3d_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                                const float* __restrict__ z, int n_cells,
                                                                const float* __restrict__ xb, int m, const float* __restrict__ yb,
                                                                int n, const float* __restrict__ zb, int k,
                                                                int32_t* __restrict__ patch_id, int32_t* __restrict__ counts) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  __shared__ float sxb[64], syb[64], szb[64];
  for (int i = threadIdx.x; i < m; i += blockDim.x) sxb[i] = xb[i];
  for (int i = threadIdx.x; i < n; i += blockDim.x) syb[i] = yb[i];
  for (int i = threadIdx.x; i < k; i += blockDim.x) szb[i] = zb[i];
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  const int ix = min(max(bucket_right(sxb, m, x[c]), 1), m - 1);
  const int iy = min(max(bucket_right(syb, n, y[c]), 1), n - 1);
  const int iz = min(max(bucket_right(szb, k, z[c]), 1), k - 1);
  const int p = ((ix - 1) * (n - 1) + (iy - 1)) * (k - 1) + (iz - 1);
  patch_id[c] = p;
  atomicAdd(counts + p, 1);
}

__global__ void __launch_bounds__(1024) patch_index_map_kernel(const int32_t* __restrict__ patch_id, int n_cells,
                                                               int capacity, long long pad_id,
                                                               long long* __restrict__ index_map) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  __shared__ int warp_tot[32];
  __shared__ int base_s;
  const int p = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  long long* row = index_map + static_cast<long long>(p) * capacity;
  if (tid == 0) base_s = 0;
  __syncthreads();
  for (int c0 = 0; c0 < n_cells; c0 += blockDim.x) {
    const int c = c0 + tid;
    const bool hit = c < n_cells && patch_id[c] == p;
    const unsigned ballot = __ballot_sync(0xffffffffu, hit);
    const int in_warp = __popc(ballot & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[warp] = __popc(ballot);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) {
      const int t = warp_tot[w];
      if (w < warp) before += t;
      total += t;
    }
    const int base = base_s;
    if (hit) {
      const int pos = base + before + in_warp;
      if (pos < capacity) row[pos] = c;
    }
    __syncthreads();
    if (tid == 0) base_s = base + total;
    __syncthreads();
  }
  for (int k = base_s + tid; k < capacity; k += blockDim.x) row[k] = pad_id;
}

struct FieldPtrs { const float* f[8]; };

// grid: (ceil(capacity*F / 256), P, S)
__global__ void __launch_bounds__(256) patch_gather_kernel(const FieldPtrs fields, long long ld_field,
                                                           const long long* __restrict__ index_map, int P,
                                                           int capacity, int F, float pad_value, int layout_pfc,
                                                           float* __restrict__ out) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;   // element inside one [C, F] (or [F, C]) patch block
  if (e >= capacity * F) return;
  const int p = blockIdx.y, s = blockIdx.z;
  int c, f;
  if (layout_pfc) { f = e / capacity; c = e - f * capacity; }   // [F, C]: consecutive threads = consecutive cells
  else { c = e / F; f = e - c * F; }                              // [C, F]: the reference's stacked layout
  const long long idx = index_map[static_cast<long long>(p) * capacity + c];
  const float v = idx >= 0 ? fields.f[f][static_cast<long long>(s) * ld_field + idx] : pad_value;
  out[(static_cast<long long>(s) * P + p) * capacity * F + e] = v;
}

__global__ void __launch_bounds__(256) patch_scatter_kernel(const float* __restrict__ part, const long long* __restrict__ index_map,
                                                            int P, int capacity, int F, int n_cells, int layout_pfc,
                                                            float* __restrict__ out) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= capacity * F) return;
  const int p = blockIdx.y, s = blockIdx.z;
  int c, f;
  if (layout_pfc) { f = e / capacity; c = e - f * capacity; }
  else { c = e / F; f = e - c * F; }
  const long long idx = index_map[static_cast<long long>(p) * capacity + c];
  if (idx < 0) return;
  out[(static_cast<long long>(s) * n_cells + idx) * F + f] = part[(static_cast<long long>(s) * P + p) * capacity * F + e];
}


// MinMaxScaler.transform / inverse_transform (utils/data_processors.py:245-252, :258-272) in the reference's own
// operation order, one IEEE fp32 operation per torch op (no contraction into FMAs): bit-identical to the eager result.
struct FieldScalers { sea_field_scaler f[8]; };
__device__ __forceinline__ float scale_fwd(float v, const sea_field_scaler& sc) {
  if (!sc.enabled) return v;
  const float std_ = __fdiv_rn(__fsub_rn(v, sc.min_val), __fsub_rn(sc.max_val, sc.min_val));
  return __fadd_rn(__fmul_rn(std_, sc.range), sc.lo);
}
__device__ __forceinline__ float scale_inv(float v, const sea_field_scaler& sc) {
  if (!sc.enabled) return v;
  const float std_ = __fdiv_rn(__fsub_rn(v, sc.lo), sc.range);
  return __fadd_rn(__fmul_rn(std_, __fsub_rn(sc.max_val, sc.min_val)), sc.min_val);
}

// fields [S, n_cells, F] (interleaved, as the reference holds them) -> scaled, padded patches [S, P, F, C] / [S, P, C, F]
__global__ void __launch_bounds__(256) patch_gather_scaled_kernel(const float* __restrict__ fields, int n_cells,
                                                                  const FieldScalers sc,
                                                                  const long long* __restrict__ index_map, int P,
                                                                  int capacity, int F, float pad_value, int layout_pfc,
                                                                  float* __restrict__ out) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= capacity * F) return;
  const int p = blockIdx.y, s = blockIdx.z;
  int c, f;
  if (layout_pfc) { f = e / capacity; c = e - f * capacity; }
  else { c = e / F; f = e - c * F; }
  const long long idx = index_map[static_cast<long long>(p) * capacity + c];
  float v = pad_value;
  if (idx >= 0) v = scale_fwd(fields[(static_cast<long long>(s) * n_cells + idx) * F + f], sc.f[f]);
  out[(static_cast<long long>(s) * P + p) * capacity * F + e] = v;
}

// padded patches -> unscaled fields [S, n_cells, F]
__global__ void __launch_bounds__(256) patch_scatter_scaled_kernel(const float* __restrict__ part,
                                                                   const long long* __restrict__ index_map, int P,
                                                                   int capacity, int F, int n_cells, int layout_pfc,
                                                                   const FieldScalers sc, float* __restrict__ out) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= capacity * F) return;
  const int p = blockIdx.y, s = blockIdx.z;
  int c, f;
  if (layout_pfc) { f = e / capacity; c = e - f * capacity; }
  else { c = e / F; f = e - c * F; }
  const long long idx = index_map[static_cast<long long>(p) * capacity + c];
  if (idx < 0) return;
  out[(static_cast<long long>(s) * n_cells + idx) * F + f] =
      scale_inv(part[(static_cast<long long>(s) * P + p) * capacity * F + e], sc.f[f]);
}

}  // namespace
}  // namespace sea

extern "C" int sea_patch_bucketize(const float* x, const float* y, int n_cells, const float* x_boundary, int m,
                                   const float* y_boundary, int n, int32_t* patch_id, int32_t* counts,
                                   sea_stream_t stream) {
  using namespace sea;
  if (!x || !y || !x_boundary || !y_boundary || !patch_id || !counts || n_cells <= 0) return SEA_ERR_INVALID;
  if (m < 2 || n < 2 || m > 64 || n > 64) return SEA_ERR_UNSUPPORTED;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  SEA_CUDA_OK(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (m - 1) * (n - 1), s));
  SEA_LAUNCH(patch_bucketize_kernel, (n_cells + 255) / 256, 256, 0, s, x, y, n_cells, x_boundary, m, y_boundary, n,
             patch_id, counts);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int sea_patch_bucketize3d(const float* x, const float* y, const float* z, int n_cells, const float* x_boundary,
                                     int m, const float* y_boundary, int n, const float* z_boundary, int k,
                                     int32_t* patch_id, int32_t* counts, sea_stream_t stream) {
  using namespace sea;
  if (!x || !y || !z || !x_boundary || !y_boundary || !z_boundary || !patch_id || !counts || n_cells <= 0) return SEA_ERR_INVALID;
  if (m < 2 || n < 2 || k < 2 || m > 64 || n > 64 || k > 64) return SEA_ERR_UNSUPPORTED;
  if (static_cast<long long>(m - 1) * (n - 1) * (k - 1) > 65535) return SEA_ERR_UNSUPPORTED;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  SEA_CUDA_OK(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (m - 1) * (n - 1) * (k - 1), s));
  SEA_LAUNCH(patch_bucketize3d_kernel, (n_cells + 255) / 256, 256, 0, s, x, y, z, n_cells, x_boundary, m, y_boundary, n,
             z_boundary, k, patch_id, counts);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int sea_patch_index_map(const int32_t* patch_id, int n_cells, int n_patches, int capacity,
                                   int64_t pad_id, int64_t* index_map, sea_stream_t stream) {
  using namespace sea;
  if (!patch_id || !index_map || n_cells <= 0 || n_patches <= 0 || capacity <= 0 || pad_id >= 0) return SEA_ERR_INVALID;
  SEA_LAUNCH(patch_index_map_kernel, n_patches, 1024, 0, reinterpret_cast<cudaStream_t>(stream), patch_id, n_cells,
             capacity, static_cast<long long>(pad_id), reinterpret_cast<long long*>(index_map));
  return static_cast<int>(cudaGetLastError());
}

extern "C" int sea_patch_gather(const float* const* fields, int n_fields, int64_t ld_field, const int64_t* index_map,
                                int n_snapshots, int n_patches, int capacity, float pad_value, int layout_pfc,
                                float* out, sea_stream_t stream) {
  using namespace sea;
  if (!fields || !index_map || !out || n_fields < 1 || n_fields > 8) return SEA_ERR_INVALID;
  if (n_snapshots <= 0 || n_patches <= 0 || capacity <= 0 || n_snapshots > 65535 || n_patches > 65535) return SEA_ERR_INVALID;
  FieldPtrs fp{};
  for (int i = 0; i < n_fields; ++i) {
    if (!fields[i]) return SEA_ERR_INVALID;
    fp.f[i] = fields[i];
  }
  dim3 grid((capacity * n_fields + 255) / 256, n_patches, n_snapshots);
  SEA_LAUNCH(patch_gather_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream), fp, static_cast<long long>(ld_field),
             reinterpret_cast<const long long*>(index_map), n_patches, capacity, n_fields, pad_value, layout_pfc, out);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int sea_patch_scatter(const float* part, const int64_t* index_map, int n_snapshots, int n_patches,
                                 int capacity, int n_fields, int n_cells, int layout_pfc, float* out,
                                 sea_stream_t stream) {
  using namespace sea;
  if (!part || !index_map || !out || n_fields < 1 || n_cells <= 0) return SEA_ERR_INVALID;
  if (n_snapshots <= 0 || n_patches <= 0 || capacity <= 0 || n_snapshots > 65535 || n_patches > 65535) return SEA_ERR_INVALID;
  dim3 grid((capacity * n_fields + 255) / 256, n_patches, n_snapshots);
  SEA_LAUNCH(patch_scatter_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream), part,
             reinterpret_cast<const long long*>(index_map), n_patches, capacity, n_fields, n_cells, layout_pfc, out);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int sea_patch_gather_scaled(const float* fields, int n_cells, int n_fields, const sea_field_scaler* scalers,
                                       const int64_t* index_map, int n_snapshots, int n_patches, int capacity,
                                       float pad_value, int layout_pfc, float* out, sea_stream_t stream) {
  using namespace sea;
  if (!fields || !index_map || !out || n_fields < 1 || n_fields > 8 || n_cells <= 0) return SEA_ERR_INVALID;
  if (n_snapshots <= 0 || n_patches <= 0 || capacity <= 0 || n_snapshots > 65535 || n_patches > 65535) return SEA_ERR_INVALID;
  FieldScalers sc{};
  for (int i = 0; i < n_fields; ++i)
    if (scalers) sc.f[i] = scalers[i];
  dim3 grid((capacity * n_fields + 255) / 256, n_patches, n_snapshots);
  SEA_LAUNCH(patch_gather_scaled_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream), fields, n_cells, sc,
             reinterpret_cast<const long long*>(index_map), n_patches, capacity, n_fields, pad_value, layout_pfc, out);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int sea_patch_scatter_scaled(const float* part, const int64_t* index_map, int n_snapshots, int n_patches,
                                        int capacity, int n_fields, int n_cells, int layout_pfc,
                                        const sea_field_scaler* scalers, float* out, sea_stream_t stream) {
  using namespace sea;
  if (!part || !index_map || !out || n_fields < 1 || n_fields > 8 || n_cells <= 0) return SEA_ERR_INVALID;
  if (n_snapshots <= 0 || n_patches <= 0 || capacity <= 0 || n_snapshots > 65535 || n_patches > 65535) return SEA_ERR_INVALID;
  FieldScalers sc{};
  for (int i = 0; i < n_fields; ++i)
    if (scalers) sc.f[i] = scalers[i];
  dim3 grid((capacity * n_fields + 255) / 256, n_patches, n_snapshots);
  SEA_LAUNCH(patch_scatter_scaled_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream), part,
             reinterpret_cast<const long long*>(index_map), n_patches, capacity, n_fields, n_cells, layout_pfc, sc, out);
  return static_cast<int>(cudaGetLastError());
}

// Library state: device properties and the driver entry point needed for TMA descriptors.
#include <cuda.h>
#include <cuda_runtime.h>

#include <mutex>

#include "../../include/sea_b200.h"
#include "internal.h"

namespace sea {
namespace {
std::mutex g_mu;
bool g_ready = false;
}  // namespace
extern bool g_pdl;
namespace {
int g_sms = 0;
TensorMapEncodeFn g_encode = nullptr;
}  // namespace

bool g_pdl = true;
bool pdl_enabled() { return g_pdl; }
namespace { thread_local bool g_fence_next = false; }
void pdl_fence_next() { g_fence_next = true; }
bool pdl_take_fence() { const bool f = g_fence_next; g_fence_next = false; return f; }
int num_sms() { return g_sms; }
TensorMapEncodeFn tensor_map_encoder() { return g_encode; }

int ensure_init() {
  if (g_ready) return SEA_OK;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return SEA_ERR_NO_DEVICE;
  return sea_init(dev);
}

int make_tmap_bf16_3d(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t mid,
                      uint64_t outer, uint64_t ld_mid, uint64_t ld_outer, uint32_t box_inner,
                      uint32_t box_mid) {
  if (g_encode == nullptr) return SEA_ERR_NO_DEVICE;
  cuuint64_t dims[3] = {inner, mid, outer};
  cuuint64_t strides[2] = {ld_mid * 2, ld_outer * 2};
  cuuint32_t box[3] = {box_inner, box_mid, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? SEA_OK : SEA_ERR_INVALID;
}
}  // namespace sea

extern "C" int sea_init(int device) {
  using namespace sea;
  std::lock_guard<std::mutex> lock(g_mu);
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return SEA_ERR_NO_DEVICE;
  if (prop.major != 10) return SEA_ERR_NO_DEVICE;  // sm_100a only: no other code path exists
  g_sms = prop.multiProcessorCount;
  if (g_encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult st;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &st);
    if (e != cudaSuccess || st != cudaDriverEntryPointSuccess || fn == nullptr)
      return SEA_ERR_NO_DEVICE;
    g_encode = reinterpret_cast<TensorMapEncodeFn>(fn);
  }
  g_ready = true;
  return SEA_OK;
}

extern "C" int sea_version(void) { return 100; }
extern "C" void sea_set_pdl(int on) { sea::g_pdl = on != 0; }
extern "C" int sea_num_sms(void) { return sea::g_sms; }

// Event plumbing for callers without a CUDA runtime binding (the Python host side): lets the data-parallel
// gradient exchange on a second stream wait for a point INSIDE sea_temporal_backward.
extern "C" int sea_event_create(void** out) {
  if (!out) return SEA_ERR_INVALID;
  cudaEvent_t ev = nullptr;
  const cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  *out = ev;
  return static_cast<int>(e);
}
extern "C" int sea_event_destroy(void* ev) {
  return ev ? static_cast<int>(cudaEventDestroy(static_cast<cudaEvent_t>(ev))) : SEA_OK;
}
extern "C" int sea_stream_wait_event(sea_stream_t stream, void* ev) {
  if (!ev) return SEA_ERR_INVALID;
  return static_cast<int>(cudaStreamWaitEvent(reinterpret_cast<cudaStream_t>(stream), static_cast<cudaEvent_t>(ev), 0));
}

extern "C" int sea_copy_rows_to_host(void* dst_host, size_t dst_pitch, const void* src_dev, size_t src_pitch,
                                     size_t row_bytes, size_t rows, sea_stream_t stream) {
  if (!dst_host || !src_dev || row_bytes > dst_pitch || row_bytes > src_pitch) return SEA_ERR_INVALID;
  if (rows == 0 || row_bytes == 0) return SEA_OK;
  return static_cast<int>(cudaMemcpy2DAsync(dst_host, dst_pitch, src_dev, src_pitch, row_bytes, rows,
                                            cudaMemcpyDeviceToHost, reinterpret_cast<cudaStream_t>(stream)));
}

extern "C" const char* sea_strerror(int code) {
  switch (code) {
    case SEA_OK: return "ok";
    case SEA_ERR_INVALID: return "sea_b200: invalid argument (null/misaligned pointer, bad size or stride)";
    case SEA_ERR_UNSUPPORTED: return "sea_b200: shape outside the supported set";
    case SEA_ERR_NO_DEVICE: return "sea_b200: no sm_100 device or driver entry point unavailable";
    case SEA_ERR_WORKSPACE: return "sea_b200: workspace too small";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "sea_b200: unknown error";
}

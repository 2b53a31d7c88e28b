// Internal host-side helpers shared by the translation units of libsea_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sea {

typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                      const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                      const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int ensure_init();                       // lazily runs sea_init(current device)
int num_sms();                           // SM count of the initialised device
TensorMapEncodeFn tensor_map_encoder();  // cuTensorMapEncodeTiled via cudaGetDriverEntryPoint

// 2-D bf16 tensor map, 128B swizzle: dims {inner, outer}, row pitch ld_elems, box {bi, bo}.
int make_tmap_bf16_2d(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer,
                      uint64_t ld_elems, uint32_t box_inner, uint32_t box_outer);
// 3-D bf16 tensor map {inner, mid, outer} with pitches in elements; box {bi, bm, 1}.
int make_tmap_bf16_3d(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t mid,
                      uint64_t outer, uint64_t ld_mid, uint64_t ld_outer, uint32_t box_inner,
                      uint32_t box_mid);

// Kernel launch with the programmatic-stream-serialization attribute (PDL).  All kernels of this
// library call griddepcontrol.wait before touching global memory, so they may be launched early.
bool pdl_enabled();
// The next SEA_LAUNCH on this thread is issued WITHOUT the PDL attribute (ordinary stream order):
// nothing launched after it can overlap anything launched before it.  Executors call this on entry
// so that weights packed by an earlier call are "static" for every kernel of this call.
void pdl_fence_next();
bool pdl_take_fence();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_enabled() && !pdl_take_fence()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define SEA_LAUNCH(kernel, grid, block, smem, stream, ...) \
  (void)::sea::launch_pdl(kernel, dim3(grid), dim3(block), smem, stream, __VA_ARGS__)

#define SEA_CUDA_OK(expr)                                  \
  do {                                                     \
    cudaError_t _e = (expr);                               \
    if (_e != cudaSuccess) return static_cast<int>(_e);    \
  } while (0)

}  // namespace sea

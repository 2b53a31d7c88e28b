// K2 dispatch: picks the tcgen05 kernel (bf16, head_dim in {64,128,256}) or the CUDA-core kernel.
#include <cuda_runtime.h>

#include "../../include/sea_b200.h"
#include "internal.h"

namespace sea {
int attention_fwd_simt(const sea_attn_args* a, cudaStream_t s);
int attention_fwd_tc(const sea_attn_args* a, cudaStream_t s);  // attention_tc.cu
bool attention_tc_supported(const sea_attn_args* a);
int g_force_simt = 0;  // test hook: route bf16 attention (fwd and bwd) to the CUDA-core kernels
}  // namespace sea

extern "C" void sea_attention_force_simt(int on) { sea::g_force_simt = on; }

extern "C" int sea_attention_fwd(const sea_attn_args* a, sea_stream_t stream) {
  using namespace sea;
  if (!a || !a->q || !a->k || !a->v || !a->o) return SEA_ERR_INVALID;
  if (a->B <= 0 || a->T <= 0 || a->n_heads <= 0 || a->head_dim <= 0) return SEA_ERR_INVALID;
  if (a->prec != SEA_PREC_BF16 && a->prec != SEA_PREC_FP32) return SEA_ERR_INVALID;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (!g_force_simt && attention_tc_supported(a)) return attention_fwd_tc(a, s);
  return attention_fwd_simt(a, s);
}

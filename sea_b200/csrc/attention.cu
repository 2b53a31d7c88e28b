// K2 dispatch: picks the tcgen05 kernel (bf16, head_dim in {64,128,256}) or the CUDA-core kernel.
#include <cuda_runtime.h>

#include "../../include/sea_b200.h"
#include "internal.h"

namespace sea {
int attention_fwd_simt(const sea_attn_args* a, cudaStream_t s);
int attention_fwd_tc(int n, const sea_attn_args* a, cudaStream_t s);  // attention_tc.cu
int attention_fwd_small(int n, const sea_attn_args* a, cudaStream_t s);  // attention_small.cu: T <= 128, mma.sync
bool attention_small_supported(const sea_attn_args* a);
bool attention_tc_supported(const sea_attn_args* a);
int g_force_simt = 0;  // test hook: route bf16 attention (fwd and bwd) to the CUDA-core kernels

static int validate(const sea_attn_args* a) {
  if (!a->q || !a->k || !a->v || !a->o) return SEA_ERR_INVALID;
  if (a->B <= 0 || a->T <= 0 || a->n_heads <= 0 || a->head_dim <= 0) return SEA_ERR_INVALID;
  if (a->prec != SEA_PREC_BF16 && a->prec != SEA_PREC_FP32) return SEA_ERR_INVALID;
  return SEA_OK;
}
}  // namespace sea

namespace sea { extern int g_attn_two_tiles; void attention_set_trace(void*); }
extern "C" void sea_attention_debug_trace(void* dev_buf) { sea::attention_set_trace(dev_buf); }
extern "C" void sea_attention_force_simt(int on) { sea::g_force_simt = on; }
extern "C" void sea_attention_two_tiles(int on) { sea::g_attn_two_tiles = on; }
namespace sea { extern int g_attn_bwd_probe; extern int g_attn_bwd_wide; }
extern "C" void sea_attention_bwd_probe(int mode) { sea::g_attn_bwd_probe = mode; }
extern "C" void sea_attention_bwd_wide(int on) { sea::g_attn_bwd_wide = on; }

extern "C" int sea_attention_fwd(const sea_attn_args* a, sea_stream_t stream) {
  return sea_attention_fwd_group(1, a, stream);
}

extern "C" int sea_attention_fwd_group(int n, const sea_attn_args* a, sea_stream_t stream) {
  using namespace sea;
  if (!a || n < 1 || n > SEA_MAX_STREAMS) return SEA_ERR_INVALID;
  bool tc = !g_force_simt;
  for (int i = 0; i < n; ++i) {
    int rc = validate(&a[i]);
    if (rc) return rc;
    if (a[i].B != a[0].B || a[i].T != a[0].T || a[i].n_heads != a[0].n_heads || a[i].head_dim != a[0].head_dim ||
        a[i].src_len != a[0].src_len || a[i].scale != a[0].scale || a[i].prec != a[0].prec || a[i].ldo != a[0].ldo ||
        a[i].dropout_p != a[0].dropout_p || a[i].dropout_seed != a[0].dropout_seed)
      return SEA_ERR_INVALID;
    tc = tc && attention_tc_supported(&a[i]);
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (a[0].dropout_p < 0.f || a[0].dropout_p >= 1.f) return SEA_ERR_INVALID;
  if (tc && attention_small_supported(&a[0])) return attention_fwd_small(n, a, s);   // same shape for all n (checked above)
  if (tc) return attention_fwd_tc(n, a, s);
  if (a[0].dropout_p > 0.f) return SEA_ERR_UNSUPPORTED;   // probability dropout lives in the tensor-core kernels only
  for (int i = 0; i < n; ++i) {
    int rc = attention_fwd_simt(&a[i], s);
    if (rc) return rc;
  }
  return SEA_OK;
}

// placeholder, replaced below by the tcgen05 attention kernel
#include "common.cuh"
#include "kernels.h"
namespace v2m {
int attn_fwd_bf16_tc(const AttnParams& p, cudaStream_t stream) {
  set_last_error("attn_fwd_bf16_tc: not built yet");
  return kUnsupported;
}
}

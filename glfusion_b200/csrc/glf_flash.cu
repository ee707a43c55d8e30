// glf_flash.cu — mode='embedded' (softmax) attention, flash-style.  Placeholder entry points until the kernel lands.
#include "glf_internal.h"

namespace glf {

int flash_fwd(const bf16*, bf16*, float*, int, int, int, cudaStream_t) {
  return set_error(GLF_ERR_UNSUPPORTED, "mode='embedded' forward kernel not available in this build");
}
int flash_bwd(const bf16*, const bf16*, const bf16*, const float*, bf16*, float*, int, int, int, cudaStream_t) {
  return set_error(GLF_ERR_UNSUPPORTED, "mode='embedded' backward kernel not available in this build");
}

}  // namespace glf

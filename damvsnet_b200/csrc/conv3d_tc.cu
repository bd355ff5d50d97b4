// tcgen05 / TMEM implicit-GEMM 3x3x3 convolution (placeholder until the kernel lands).
#include "common.cuh"

namespace damvs {

int conv3d_tc_launch(const damvs_conv3d_desc*, const void*, const void*, const float*, const float*, const void*,
                     void*, cudaStream_t) {
  return set_error(DAMVS_ERR_UNSUPPORTED, "conv3d: tcgen05 implementation not built yet");
}
size_t conv3d_tc_packed_bytes(const damvs_conv3d_desc*) { return 0; }
int conv3d_tc_pack(const damvs_conv3d_desc*, const float*, void*, cudaStream_t) {
  return set_error(DAMVS_ERR_UNSUPPORTED, "conv3d: tcgen05 implementation not built yet");
}

}  // namespace damvs

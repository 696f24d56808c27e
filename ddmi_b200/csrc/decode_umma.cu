// tcgen05 path placeholder (replaced by the real kernels; keeps the library linkable)
#include "common.cuh"
namespace ddmi {
int launch_image_umma(const PlaneSet&, int, int, const float*, const float*, long long, const void*, size_t,
                      const float*, size_t, float*, cudaStream_t) {
  set_error("tcgen05 image kernel not built yet");
  return DDMI_ERR_UNSUPPORTED;
}
int launch_selftest_umma(const float*, const float*, float*, int, int, cudaStream_t) {
  set_error("tcgen05 selftest not built yet");
  return DDMI_ERR_UNSUPPORTED;
}
}  // namespace ddmi

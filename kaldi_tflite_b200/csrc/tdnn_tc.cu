// tcgen05 / TMEM implicit-GEMM engine for the TDNN affine layers (KTF_PREC_BF16).
// Placeholder until the tensor-core kernel lands: creation fails loudly, nothing falls back.
#include "tdnn_internal.cuh"

namespace ktf {

int affine_tc_prepare(ktf_affine*, const float*) {
  set_error("KTF_PREC_BF16 (tcgen05 engine) is not available in this build");
  return KTF_EINVAL;
}
void affine_tc_release(ktf_affine*) {}
int affine_tc_forward(const ktf_affine*, const float*, const int64_t*, const int64_t*, int64_t, int64_t,
                      int64_t, float*, float*, cudaStream_t) {
  set_error("KTF_PREC_BF16 (tcgen05 engine) is not available in this build");
  return KTF_EINVAL;
}

}  // namespace ktf

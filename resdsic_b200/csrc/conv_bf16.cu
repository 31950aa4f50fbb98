// bf16 tensor-core implicit-GEMM convolution (tcgen05 / TMEM) -- placeholder until
// the sm_100a kernel lands; fails loudly rather than falling back.
#include "common.cuh"

int rdsic_conv_forward_bf16(const rdsic_conv_desc* d, cudaStream_t stream) {
  (void)d; (void)stream;
  return RDSIC_E_UNSUPPORTED;
}

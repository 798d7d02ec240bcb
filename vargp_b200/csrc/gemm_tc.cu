// tcgen05 / TMA 3xTF32 GEMM (placeholder: lands after the SIMT path is parity-green).
#include "common.cuh"

int vargp_tc_init() { return 0; }

extern "C" int vargp_gemm_tc(const vargp_gemm_t* g, void* stream) {
  (void)g; (void)stream;
  return VARGP_ERR_UNSUPPORTED;
}

// Library bookkeeping: init, version, error strings, launch counter.
#include <cstdlib>

#include "common.cuh"

namespace vargp {
int64_t g_launches = 0;
int g_pdl = 2;
int g_device = -1;
}  // namespace vargp

using namespace vargp;

extern "C" const char* vargp_version(void) { return "vargp_sm100 0.1 (sm_100a)"; }

extern "C" const char* vargp_strerror(int code) {
  switch (code) {
    case 0: return "ok";
    case VARGP_ERR_ARG: return "invalid argument";
    case VARGP_ERR_UNSUPPORTED: return "problem shape not supported by this entry point";
    case VARGP_ERR_NOT_INIT: return "vargp_init() has not been called";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
  }
}

extern "C" int64_t vargp_launch_count(void) { return g_launches; }

extern "C" int vargp_set_pdl(int mode) {
  const int old = g_pdl;
  g_pdl = mode < 0 ? 0 : (mode > 2 ? 2 : mode);
  return old;
}

int vargp_tc_init();   // gemm_tc.cu

extern "C" int vargp_init(int device) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess) return (int)e;
  if (device < 0 || device >= ndev) return VARGP_ERR_ARG;
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return (int)e;
  if (prop.major != 10) return VARGP_ERR_UNSUPPORTED;   // sm_100a only: no other code path exists
  g_device = device;
  const char* pdl = getenv("VARGP_PDL");
  if (pdl) vargp_set_pdl(atoi(pdl));
  return vargp_tc_init();
}

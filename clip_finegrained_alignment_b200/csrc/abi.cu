// ABI bookkeeping for libcfa_b200.so
#include "common.cuh"

extern "C" int cfa_abi_version(void) { return CFA_ABI_VERSION; }

extern "C" const char* cfa_error_string(int code) {
  switch (code) {
    case CFA_OK: return "ok";
    case CFA_ERR_BAD_ARG: return "cfa: bad argument";
    case CFA_ERR_UNSUPPORTED: return "cfa: unsupported dtype or shape for the sm_100a kernels";
    case CFA_ERR_WORKSPACE: return "cfa: workspace missing or too small";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "cfa: unknown error";
}

// ABI bookkeeping entry points of libslate_b200.so.
#include "sqb_common.cuh"

extern "C" int sqb_abi_version(void) { return 1; }

extern "C" const char* sqb_error_string(int code) {
  switch (code) {
    case SQB_OK: return "ok";
    case SQB_E_BADARG: return "bad argument (null pointer, non-positive size or nd out of range)";
    case SQB_E_UNSUPPORTED: return "size outside what the sm_100a kernels implement";
    case SQB_E_WORKSPACE: return "workspace missing or too small";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown error";
}

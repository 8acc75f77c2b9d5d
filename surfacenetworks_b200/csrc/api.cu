// api.cu -- version / status strings of the C ABI (include/surfnet_b200.h).
#include "common.cuh"

SN_API int sn_version(void) { return SN_VERSION; }

SN_API const char* sn_status_string(int status) {
  switch (status) {
    case SN_OK: return "ok";
    case SN_ERR_ARG: return "invalid argument (null pointer, negative size or inconsistent shapes)";
    case SN_ERR_UNSUPPORTED: return "unsupported shape or alignment for this entry point";
    case SN_ERR_WORKSPACE: return "workspace missing or smaller than the *_ws_bytes() query";
    case SN_ERR_OVERFLOW: return "size does not fit the 32-bit index format / launch grid";
    default: break;
  }
  if (status > 0) return cudaGetErrorString((cudaError_t)status);
  return "unknown status";
}

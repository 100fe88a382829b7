// Status strings / version of the lws_b200 C ABI (include/lws.h).
#include "lws_common.cuh"

extern "C" const char* lws_status_string(int status) {
  switch (status) {
    case LWS_OK: return "LWS_OK";
    case LWS_ERR_BAD_SHAPE: return "LWS_ERR_BAD_SHAPE";
    case LWS_ERR_BAD_ALIGN: return "LWS_ERR_BAD_ALIGN";
    case LWS_ERR_NULL_PTR: return "LWS_ERR_NULL_PTR";
    case LWS_ERR_WORKSPACE_TOO_SMALL: return "LWS_ERR_WORKSPACE_TOO_SMALL";
    case LWS_ERR_UNSUPPORTED: return "LWS_ERR_UNSUPPORTED";
    default: break;
  }
  if (status > 0) return cudaGetErrorString((cudaError_t)status);
  return "LWS_ERR_UNKNOWN";
}

extern "C" const char* lws_version(void) { return "lws_b200 0.1.0 sm_100a"; }

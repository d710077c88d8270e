// Library-level entry points of libcatb200: version, error string, per-device initialisation.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace catb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return CATB_ERR_CUDA;
  }
  return CATB_OK;
}

int init_igemm_attributes();
int init_halo_attributes();
int init_simt_attributes();
int init_halo_wgrad_attributes();
int init_halo_persist_attributes();

}  // namespace catb

extern "C" const char* catb_version(void) { return "catb200 0.2 (sm_100a)"; }

extern "C" const char* catb_last_error_string(void) { return catb::g_err; }

extern "C" int catb_init(int device) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    catb::set_error("no CUDA device visible: libcatb200 has no CPU fallback");
    return CATB_ERR_NO_DEVICE;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    catb::set_error("cudaGetDeviceProperties(%d) failed", device);
    return CATB_ERR_CUDA;
  }
  if (prop.major != 10) {
    catb::set_error("device %d is sm_%d%d; libcatb200 is built for sm_100a only", device, prop.major, prop.minor);
    return CATB_ERR_NO_DEVICE;
  }
  if (cudaSetDevice(device) != cudaSuccess) {
    catb::set_error("cudaSetDevice(%d) failed", device);
    return CATB_ERR_CUDA;
  }
  if (int e = catb::init_igemm_attributes()) return e;
  if (int e = catb::init_halo_attributes()) return e;
  if (int e = catb::init_simt_attributes()) return e;
  if (int e = catb::init_halo_persist_attributes()) return e;
  return catb::init_halo_wgrad_attributes();
}

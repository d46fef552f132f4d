// core.cu -- error reporting and device checks shared by every entry point of libmarlb200.
#include "common.cuh"
#include <string.h>
#include <stdlib.h>

namespace marl {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// No CPU fallback: a missing / non-Blackwell device is an error, never a silent slow path.
int check_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_error("libmarlb200: no CUDA device available (%s); this library has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    return MARL_ECUDA;
  }
  if (device < 0 || device >= n) { set_error("libmarlb200: device %d out of range (0..%d)", device, n - 1); return MARL_EINVAL; }
  cudaDeviceProp p;
  MARL_CUDA_TRY(cudaGetDeviceProperties(&p, device));
  if (p.major != 10) {
    set_error("libmarlb200: device %d is sm_%d%d; this build targets sm_100a (B200) only", device, p.major, p.minor);
    return MARL_ECUDA;
  }
  MARL_CUDA_TRY(cudaSetDevice(device));
  return MARL_OK;
}

static int env_flag(const char* name, int dflt) { const char* v = getenv(name); return v && *v ? (v[0] != '0') : dflt; }
static int g_tc_forward = 1, g_tc_backward = 1, g_tc_pingpong = env_flag("MARL_TC_PINGPONG", 1);   // the environment variable only moves the default
int tc_pingpong_enabled() { return g_tc_pingpong; }
int tc_forward_enabled() { return g_tc_forward; }
int tc_backward_enabled() { return g_tc_backward; }

}  // namespace marl

extern "C" {
/* Process-wide options.  "tensor_core_forward": 1 = forward-only passes (act, target networks) run on tcgen05 with the
 * 3xTF32 split (default), 0 = FP32 FFMA kernels. */
int marl_set_option(const char* name, int32_t value) {
  if (name && strcmp(name, "tensor_core_forward") == 0) { marl::g_tc_forward = value ? 1 : 0; return MARL_OK; }
  if (name && strcmp(name, "tensor_core_backward") == 0) { marl::g_tc_backward = value ? 1 : 0; return MARL_OK; }
  /* 1 (default): tensor-core kernels with two accumulator buffers in TMEM -- the epilogue of one 128-row tile runs under the MMAs of the
   * next; 0: the round-1 kernels (one tile at a time, tensor and CUDA-core phases alternate) */
  if (name && strcmp(name, "tensor_core_pingpong") == 0) { marl::g_tc_pingpong = value ? 1 : 0; return MARL_OK; }
  marl::set_error("marl_set_option: unknown option '%s'", name ? name : "(null)");
  return MARL_EINVAL;
}
int marl_version(void) { return MARL_ABI_VERSION; }
const char* marl_last_error(void) { return marl::g_err; }
}

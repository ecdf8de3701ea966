// Error plumbing, launch accounting and device queries shared by all entry points.
#include "common.cuh"

namespace duo {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: CUDA error %d (%s)", what, static_cast<int>(e), cudaGetErrorString(e));
  return DUO_ERR_CUDA;
}

void count_launch() { ++g_launches; }

int device_sm_count() {
  static int cached_dev = -1;
  static int cached_sms = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
    cached_dev = dev;
    cached_sms = sms;
  }
  return cached_sms;
}

bool first_use_on_device(uint64_t& mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
  const uint64_t bit = uint64_t(1) << dev;
  if (mask & bit) return false;
  mask |= bit;
  return true;
}

}  // namespace duo

extern "C" const char* duo_last_error(void) { return duo::g_err; }
extern "C" int duo_abi_version(void) { return 4; }  // 3: duo_gemm_args statistics forwarding; 4: duo_conv2d / duo_stem_*
extern "C" int64_t duo_launch_count(void) { return duo::g_launches; }
extern "C" void duo_launch_count_reset(void) { duo::g_launches = 0; }

// Host-side status plumbing shared by every entry point of libsgv3d_b200.so.
#include <stdarg.h>

#include "common.cuh"

namespace sgv3d {
namespace {
thread_local char g_error[512] = "";
thread_local int64_t g_launches = 0;
}  // namespace

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int64_t &launch_counter() { return g_launches; }
}  // namespace sgv3d

extern "C" int sgv3d_abi_version(void) { return SGV3D_ABI_VERSION; }
extern "C" const char *sgv3d_last_error(void) { return sgv3d::g_error; }
extern "C" int64_t sgv3d_launch_count(int reset) {
  const int64_t v = sgv3d::g_launches;
  if (reset) sgv3d::g_launches = 0;
  return v;
}

// Host-side status plumbing shared by every entry point of libsgv3d_b200.so.
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

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

// ---- per-kernel event timing ---------------------------------------------------------------------
namespace {
struct Span { const char *name; cudaEvent_t a, b; };
struct Acc { long launches = 0; double ms = 0.0; };
thread_local bool g_prof_on = false;
thread_local std::vector<Span> *g_spans = nullptr;      // in flight since the last report
thread_local std::vector<cudaEvent_t> *g_pool = nullptr;  // recycled events
thread_local cudaStream_t g_stream = nullptr;
thread_local cudaEvent_t g_prev = nullptr;  // end of the previous kernel == start of the next one

cudaEvent_t take_event() {
  if (!g_pool) g_pool = new std::vector<cudaEvent_t>();
  if (!g_pool->empty()) {
    cudaEvent_t e = g_pool->back();
    g_pool->pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

void prof_begin(cudaStream_t stream) {
  g_stream = stream;
  if (!g_prof_on) return;
  g_prev = take_event();
  cudaEventRecord(g_prev, stream);
}

void prof_mark(const char *name) {
  if (!g_prof_on || !g_prev) return;
  if (!g_spans) g_spans = new std::vector<Span>();
  cudaEvent_t b = take_event();
  cudaEventRecord(b, g_stream);
  g_spans->push_back({name, g_prev, b});
  g_prev = b;
}
}  // namespace sgv3d

extern "C" int sgv3d_profile_enable(int on) {
  sgv3d::g_prof_on = on != 0;
  return SGV3D_OK;
}

// Waits for the recorded events, writes one "name,launches,total_ms" line per kernel into buf
// and clears the accumulators.  Returns the number of bytes needed (excluding the NUL).
extern "C" long sgv3d_profile_report(char *buf, size_t buflen) {
  using namespace sgv3d;
  std::map<std::string, Acc> acc;
  if (g_spans) {
    for (const Span &sp : *g_spans) {
      float ms = 0.f;
      cudaEventSynchronize(sp.b);
      cudaEventElapsedTime(&ms, sp.a, sp.b);
      Acc &a = acc[sp.name];
      a.launches += 1;
      a.ms += ms;
    }
    // consecutive spans share events (b of one == a of the next): recycle each event once
    std::vector<cudaEvent_t> evs;
    for (const Span &sp : *g_spans) { evs.push_back(sp.a); evs.push_back(sp.b); }
    std::sort(evs.begin(), evs.end());
    evs.erase(std::unique(evs.begin(), evs.end()), evs.end());
    for (cudaEvent_t e : evs) g_pool->push_back(e);
    g_spans->clear();
    g_prev = nullptr;
  }
  std::string out;
  char line[256];
  for (const auto &kv : acc) {
    snprintf(line, sizeof(line), "%s,%ld,%.6f\n", kv.first.c_str(), kv.second.launches, kv.second.ms);
    out += line;
  }
  if (buf && buflen > 0) {
    const size_t n = out.size() < buflen - 1 ? out.size() : buflen - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return (long)out.size();
}

extern "C" int sgv3d_abi_version(void) { return SGV3D_ABI_VERSION; }
extern "C" const char *sgv3d_last_error(void) { return sgv3d::g_error; }
extern "C" int64_t sgv3d_launch_count(int reset) {
  const int64_t v = sgv3d::g_launches;
  if (reset) sgv3d::g_launches = 0;
  return v;
}

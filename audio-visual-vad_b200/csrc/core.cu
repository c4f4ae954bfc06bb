// Error plumbing, version and launch accounting for libavvad.
#include <stdlib.h>

#include "common.cuh"

namespace avvad {
static thread_local std::string t_last_error;
std::atomic<uint64_t> g_launches{0};
void set_error(const std::string& msg) { t_last_error = msg; }
bool sync_debug() {
  static int v = [] {
    const char* e = getenv("AVVAD_SYNC_DEBUG");
    return (e && atoi(e) != 0) ? 1 : 0;
  }();
  return v != 0;
}
}  // namespace avvad

extern "C" const char* avvad_last_error(void) { return avvad::t_last_error.c_str(); }
extern "C" int avvad_version(void) { return 100; }
extern "C" uint64_t avvad_launch_count(void) { return avvad::g_launches.load(); }

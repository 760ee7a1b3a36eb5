// Error state + version for the C-ABI library.
#include <stdarg.h>
#include <atomic>
#include <string.h>

#include "common.cuh"
#include "../../include/b200gat.h"

namespace b200gat {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launches() { return g_launches.load(); }
}  // namespace b200gat

extern "C" const char* b200gat_last_error(void) { return b200gat::g_err; }
extern "C" int b200gat_abi_version(void) { return B200GAT_ABI_VERSION; }
extern "C" int64_t b200gat_launch_count(void) { return (int64_t)b200gat::launches(); }

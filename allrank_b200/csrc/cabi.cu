// C-ABI bookkeeping: thread-local error string, ABI version, launch counter.
#include <atomic>
#include <cstring>

#include "common.h"

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void arb_set_error(const char* msg) {
  std::strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
}
void arb_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" const char* arb_last_error(void) { return g_err; }
extern "C" int32_t arb_abi_version(void) { return 1; }
extern "C" int64_t arb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

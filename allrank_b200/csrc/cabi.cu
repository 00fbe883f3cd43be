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
extern "C" int32_t arb_abi_version(void) { return 3; }
extern "C" int64_t arb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// ---------------------------------------------------------------- per-launch timing
#include <mutex>
#include <vector>
namespace {
struct ProfRec { int cls; double work; double bytes; cudaEvent_t e0, e1; };
std::vector<ProfRec> g_recs;
std::mutex g_prof_mu;
bool g_prof_on = false;
constexpr size_t kMaxRecs = 1 << 16;
}  // namespace

ProfScope::ProfScope(int cls, double work, cudaStream_t s, double bytes) : idx(-1), st(s) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (g_recs.size() >= kMaxRecs) return;
  ProfRec r{cls, work, bytes, nullptr, nullptr};
  if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
  cudaEventRecord(r.e0, st);
  g_recs.push_back(r);
  idx = int(g_recs.size()) - 1;
}
ProfScope::~ProfScope() {
  if (idx < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEventRecord(g_recs[idx].e1, st);
}

extern "C" void arb_prof_enable(int32_t on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  g_recs.clear();
  g_prof_on = on != 0;
}
// Sums device time (ms), work units and launch count of one kernel class since arb_prof_enable(1).
static double g_last_bytes[ARB_PROF_CLASSES] = {0};
extern "C" double arb_prof_last_bytes(int32_t cls) { return (cls >= 0 && cls < ARB_PROF_CLASSES) ? g_last_bytes[cls] : 0.0; }

extern "C" int32_t arb_prof_collect(int32_t cls, double* total_ms, double* total_work, int64_t* launches) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  double ms = 0, work = 0, bytes = 0;
  long long n = 0;
  for (auto& r : g_recs) {
    if (r.cls != cls) continue;
    if (cudaEventSynchronize(r.e1) != cudaSuccess) { arb_set_error("arb_prof_collect: event sync failed"); return ARB_E_CUDA; }
    float t = 0;
    cudaEventElapsedTime(&t, r.e0, r.e1);
    ms += t; work += r.work; bytes += r.bytes; ++n;
  }
  if (cls >= 0 && cls < ARB_PROF_CLASSES) g_last_bytes[cls] = bytes;
  if (total_ms) *total_ms = ms;
  if (total_work) *total_work = work;
  if (launches) *launches = n;
  return ARB_OK;
}

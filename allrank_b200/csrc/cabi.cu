// C-ABI bookkeeping: thread-local error string, ABI version, launch counter.
#include <atomic>
#include <cstring>

#include "common.h"
#include "defaults.h"

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void arb_set_error(const char* msg) {
  std::strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
}
void arb_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" const char* arb_last_error(void) { return g_err; }
extern "C" int32_t arb_abi_version(void) { return 4; }
extern "C" int64_t arb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// programmatic dependent launch: on by default; arb_set_pdl(0) / ARB_PDL=0 launches every kernel fully serialised
static std::atomic<int> g_pdl{ARB_DEFAULT_PDL};
bool arb_pdl_enabled() { return g_pdl.load(std::memory_order_relaxed) != 0; }
extern "C" void arb_set_pdl(int32_t on) { g_pdl.store(on ? 1 : 0, std::memory_order_relaxed); }

// ---------------------------------------------------------------- per-launch timing
#include <mutex>
#include <vector>
namespace {
struct ProfRec { int cls; double work; double bytes; cudaEvent_t e0, e1; char name[56]; };
std::vector<ProfRec> g_recs;
std::mutex g_prof_mu;
bool g_prof_on = false;
constexpr size_t kMaxRecs = 1 << 16;
double g_row_frac = 1.0;
double g_attn_frac = 1.0;
}  // namespace

double arb_row_frac() { return g_row_frac; }
void arb_set_row_frac(double f) { g_row_frac = f; }
double arb_attn_frac() { return g_attn_frac; }
void arb_set_attn_frac(double f) { g_attn_frac = f; }
bool arb_prof_enabled() { return g_prof_on; }

ProfScope::ProfScope(int cls, double work, cudaStream_t s, double bytes, const char* name) : idx(-1), st(s) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (g_recs.size() >= kMaxRecs) return;
  ProfRec r{cls, work, bytes, nullptr, nullptr, {0}};
  std::strncpy(r.name, name ? name : "?", sizeof(r.name) - 1);
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

// Per-kernel table since arb_prof_enable(1): one line per distinct launch name,
//   name<TAB>class<TAB>launches<TAB>total_ms<TAB>total_work<TAB>total_bytes
// (work = flops for class 0, algorithmic bytes otherwise; bytes = algorithmic HBM bytes where the launcher states
// them).  Returns the number of bytes written (truncated to cap - 1), or ARB_E_CUDA.
#include <cstdio>
#include <map>
#include <string>
extern "C" int64_t arb_prof_report(char* buf, int64_t cap) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  struct Agg { int cls; long long n; double ms, work, bytes; };
  std::map<std::string, Agg> table;
  std::vector<std::string> order;
  for (auto& r : g_recs) {
    if (cudaEventSynchronize(r.e1) != cudaSuccess) { arb_set_error("arb_prof_report: event sync failed"); return ARB_E_CUDA; }
    float t = 0;
    cudaEventElapsedTime(&t, r.e0, r.e1);
    auto it = table.find(r.name);
    if (it == table.end()) { it = table.emplace(r.name, Agg{r.cls, 0, 0, 0, 0}).first; order.push_back(r.name); }
    it->second.n += 1; it->second.ms += t; it->second.work += r.work; it->second.bytes += r.bytes;
  }
  std::string out;
  char line[256];
  for (auto& k : order) {
    const Agg& a = table[k];
    std::snprintf(line, sizeof line, "%s\t%d\t%lld\t%.6f\t%.6e\t%.6e\n", k.c_str(), a.cls, a.n, a.ms, a.work, a.bytes);
    out += line;
  }
  if (!buf || cap <= 0) return int64_t(out.size());
  const int64_t n = std::min<int64_t>(cap - 1, int64_t(out.size()));
  std::memcpy(buf, out.data(), size_t(n));
  buf[n] = 0;
  return n;
}

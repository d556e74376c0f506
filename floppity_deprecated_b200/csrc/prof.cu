#include "prof.h"
#include "common.cuh"
#include <vector>
#include <mutex>

namespace fp {
namespace {
struct Rec { cudaEvent_t a, b; int tag; double work, bytes; };
std::mutex g_mu;
bool g_on = false;
std::vector<Rec> g_recs;
std::vector<cudaEvent_t> g_pool;
int64_t g_launches = 0;
}  // namespace

ProfScope::ProfScope(int tag, double work, cudaStream_t s, double bytes) : s_(s), slot_(-1) {
  std::lock_guard<std::mutex> lk(g_mu);
  ++g_launches;
  if (!g_on) return;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(s, &st);
  if (st != cudaStreamCaptureStatusNone) return;
  Rec r{};
  for (cudaEvent_t* e : {&r.a, &r.b}) {
    if (!g_pool.empty()) { *e = g_pool.back(); g_pool.pop_back(); }
    else cudaEventCreate(e);
  }
  r.tag = tag; r.work = work; r.bytes = bytes;
  cudaEventRecord(r.a, s);
  g_recs.push_back(r);
  slot_ = (int)g_recs.size() - 1;
}
ProfScope::~ProfScope() {
  if (slot_ < 0) return;
  std::lock_guard<std::mutex> lk(g_mu);
  cudaEventRecord(g_recs[slot_].b, s_);
}
}  // namespace fp

using namespace fp;
extern "C" {
int64_t fp_launch_count(void) { std::lock_guard<std::mutex> lk(g_mu); return g_launches; }
int fp_prof_enable(int32_t on) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_on = on != 0;
  return 0;
}
// Synchronises the device, then fills per-class totals: ms[NTAGS], launches[NTAGS], work[NTAGS]; clears the records.
static int collect(double* ms, int64_t* launches, double* work, double* bytes) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return (int)e;
  std::lock_guard<std::mutex> lk(g_mu);
  for (int i = 0; i < NTAGS; ++i) { ms[i] = 0; launches[i] = 0; work[i] = 0; if (bytes) bytes[i] = 0; }
  for (auto& r : g_recs) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
      ms[r.tag] += t; launches[r.tag] += 1; work[r.tag] += r.work;
      if (bytes) bytes[r.tag] += r.bytes;
    }
    g_pool.push_back(r.a); g_pool.push_back(r.b);
  }
  g_recs.clear();
  return 0;
}
int fp_prof_collect(double* ms, int64_t* launches, double* work) { return collect(ms, launches, work, nullptr); }
// same, plus the algorithmic HBM bytes of the GEMM classes (bytes[NTAGS])
int fp_prof_collect2(double* ms, int64_t* launches, double* work, double* bytes) { return collect(ms, launches, work, bytes); }
}

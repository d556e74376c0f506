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
static const char* const g_tag_names[NTAGS] = {"k_gemm", "k_gemm_tc_nt", "k_gemm_tc_tn", "k_cond_fwd", "k_rqs_apply_fwd", "k_rqs_apply_inv",
                                               "k_rqs_backward", "k_cond_small_fwd", "k_cond_small_bwd", "k_embed_conv", "k_small_wgrad",
                                               "other"};
int32_t fp_prof_ntags(void) { return NTAGS; }
const char* fp_prof_tag_name(int32_t tag) { return (tag >= 0 && tag < NTAGS) ? g_tag_names[tag] : ""; }
// Synchronises the device, then fills per-kernel totals (arrays of fp_prof_ntags() entries): elapsed ms (CUDA events on
// the launching stream), launches, algorithmic flops and algorithmic HBM bytes; clears the records.
int fp_prof_collect(double* ms, int64_t* launches, double* flops, double* bytes) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return (int)e;
  std::lock_guard<std::mutex> lk(g_mu);
  for (int i = 0; i < NTAGS; ++i) { ms[i] = 0; launches[i] = 0; flops[i] = 0; bytes[i] = 0; }
  for (auto& r : g_recs) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
      ms[r.tag] += t; launches[r.tag] += 1; flops[r.tag] += r.work; bytes[r.tag] += r.bytes;
    }
    g_pool.push_back(r.a); g_pool.push_back(r.b);
  }
  g_recs.clear();
  return 0;
}
}

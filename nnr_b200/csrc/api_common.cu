// Error channel + launch counter shared by all translation units.
#include "common.cuh"
#include "../../include/nnr_b200.h"
#include <atomic>

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void nnr_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void nnr_count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

extern "C" const char* nnr_last_error(void) { return g_err; }
extern "C" int nnr_abi_version(void) { return NNR_ABI_VERSION; }
extern "C" uint64_t nnr_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// ------------------------------------------------------------------------------------------------
// optional kernel-level timing of the composite nnr_gemm op (CUDA events on the launching stream)
// ------------------------------------------------------------------------------------------------
#include <vector>
#include <mutex>
struct ProfRec { int tag; cudaEvent_t e0, e1; double flops; };
static std::vector<ProfRec> g_prof;
static std::mutex g_prof_mu;
static std::atomic<int> g_prof_on{0};

int nnr_prof_enabled() { return g_prof_on.load(std::memory_order_relaxed); }
void* nnr_prof_begin(int tag, double flops, void* stream) {
  if (!nnr_prof_enabled()) return nullptr;
  ProfRec* r = new ProfRec;
  r->tag = tag; r->flops = flops;
  cudaEventCreate(&r->e0); cudaEventCreate(&r->e1);
  cudaEventRecord(r->e0, (cudaStream_t)stream);
  return r;
}
void nnr_prof_end(void* h, void* stream) {
  if (!h) return;
  ProfRec* r = (ProfRec*)h;
  cudaEventRecord(r->e1, (cudaStream_t)stream);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(*r);
  delete r;
}
extern "C" int nnr_profile_enable(int on) {
  g_prof_on.store(on ? 1 : 0);
  return 0;
}
// out[3*tag + 0] = total ms, [3*tag + 1] = launches, [3*tag + 2] = algorithmic flops; ntags entries; clears the log
extern "C" int nnr_profile_read(double* out, int ntags) {
  cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int i = 0; i < 3 * ntags; ++i) out[i] = 0.0;
  for (auto& r : g_prof) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    if (r.tag >= 0 && r.tag < ntags) { out[3 * r.tag] += ms; out[3 * r.tag + 1] += 1.0; out[3 * r.tag + 2] += r.flops; }
    cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
  }
  g_prof.clear();
  return 0;
}

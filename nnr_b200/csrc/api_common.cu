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

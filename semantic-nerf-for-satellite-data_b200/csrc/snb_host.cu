// Host-side plumbing shared by every entry point: error strings, device query.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "snb_common.cuh"

namespace snb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return (int)e;
}

int num_sms() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_error("no CUDA device");
    return -1;
  }
  if (dev != cached_dev) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
      set_error("cudaGetDeviceProperties failed");
      return -1;
    }
    if (p.major != 10) {
      set_error("libsnb is built for sm_100a only; device is sm_%d%d", p.major, p.minor);
      return -1;
    }
    cached = p.multiProcessorCount;
    cached_dev = dev;
  }
  return cached;
}

// ---- measurement hooks ------------------------------------------------------------------------------
static bool g_prof_time = false;
static long long g_launches = 0, g_gemm_launches = 0;
static double g_gemm_macs = 0.0;
static std::vector<cudaEvent_t> g_ev_pool;
static size_t g_ev_used = 0;

void count_launch() { ++g_launches; }

struct GemmRecord {
  int epi, M, N, K, cg, splits;
  double macs;
};
static std::vector<GemmRecord> g_records;

bool profile_gemm_begin(cudaStream_t st, double macs, int epi, int M, int N, int K, int cg, int splits) {
  ++g_gemm_launches;
  g_gemm_macs += macs;
  if (!g_prof_time) return false;
  g_records.push_back({epi, M, N, K, cg, splits, macs});
  if (g_ev_used + 2 > g_ev_pool.size()) {
    for (int i = 0; i < 256; ++i) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      g_ev_pool.push_back(e);
    }
  }
  cudaEventRecord(g_ev_pool[g_ev_used], st);
  return true;
}
void profile_gemm_end(cudaStream_t st) {
  cudaEventRecord(g_ev_pool[g_ev_used + 1], st);
  g_ev_used += 2;
}

}  // namespace snb

extern "C" void snb_profile_begin(int time_gemms) {
  snb::g_prof_time = time_gemms != 0;
  snb::g_launches = snb::g_gemm_launches = 0;
  snb::g_gemm_macs = 0.0;
  snb::g_ev_used = 0;
  snb::g_records.clear();
}

extern "C" int snb_profile_end(double* gemm_ms, int64_t* gemm_launches, int64_t* total_launches, double* gemm_macs) {
  using namespace snb;
  SNB_CUDA(cudaDeviceSynchronize());
  double ms = 0.0;
  // SNB_PROF_DUMP=<file>: per-launch list (epilogue, shape, SM-pair mode, split count, device time) for tools/
  FILE* dump = nullptr;
  if (const char* path = getenv("SNB_PROF_DUMP")) dump = fopen(path, "a");
  if (dump) fprintf(dump, "# idx epi M N K cta_group splits us tflops_executed\n");
  for (size_t i = 0; i + 1 < g_ev_used; i += 2) {
    float t = 0.f;
    SNB_CUDA(cudaEventElapsedTime(&t, g_ev_pool[i], g_ev_pool[i + 1]));
    ms += t;
    if (dump && i / 2 < g_records.size()) {
      const GemmRecord& r = g_records[i / 2];
      fprintf(dump, "%zu %d %d %d %d %d %d %.2f %.1f\n", i / 2, r.epi, r.M, r.N, r.K, r.cg, r.splits, t * 1e3,
              2.0 * r.macs / (t * 1e-3) / 1e12);
    }
  }
  if (dump) fclose(dump);
  if (gemm_ms) *gemm_ms = ms;
  if (gemm_launches) *gemm_launches = g_gemm_launches;
  if (total_launches) *total_launches = g_launches;
  if (gemm_macs) *gemm_macs = g_gemm_macs;
  g_prof_time = false;
  g_ev_used = 0;
  return 0;
}

extern "C" int64_t snb_profile_launch_count(void) { return snb::g_launches; }
extern "C" void snb_profile_add_launches(int64_t n) { snb::g_launches += n; }

extern "C" int snb_version(void) { return SNB_VERSION; }
extern "C" const char* snb_last_error(void) { return snb::g_err; }
extern "C" int snb_device_sms(void) { return snb::num_sms(); }

// Host-side plumbing of libsgg_b200: error string, version, TMA tensor-map encoding.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include <cudaTypedefs.h>
#include <cxxabi.h>

#include "common.cuh"
#include "../../include/sgg_b200.h"

namespace sgg {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

// demangled name of a kernel entry point ("void sgg::gemm_kernel<256, false, true, 2>(...)")
static std::string kernel_name(const void* func) {
  const char* name = nullptr;
  if (cudaFuncGetName(&name, func) != cudaSuccess || !name) return "?";
  int status = 0;
  char* dm = abi::__cxa_demangle(name, nullptr, nullptr, &status);
  std::string out = (status == 0 && dm) ? dm : name;
  free(dm);
  return out;
}

// launches per kernel entry point (always on; a map update per launch): lets a caller prove which kernels ran
static std::mutex g_kern_mu;
static std::map<const void*, long long> g_kern_counts;
void note_kernel(const void* func) {
  std::lock_guard<std::mutex> lk(g_kern_mu);
  g_kern_counts[func] += 1;
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("SGG_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

// SGG_L2_POLICY=1: L2 eviction-priority hints (annotation tiles evict_last; weight / optimiser streams evict_first)
bool l2_policy_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("SGG_L2_POLICY");
    on = (e && e[0] == '1') ? 1 : 0;
  }
  return on == 1;
}

// ---- per-launch event timing (debug / profiling aid)
struct TimedLaunch { cudaEvent_t e0, e1; const void* func; dim3 grid, block; };
static std::vector<TimedLaunch> g_timed;
static thread_local cudaEvent_t t_ev0 = nullptr;
bool timing_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("SGG_TIMING");
    on = (e && e[0] == '1') ? 1 : 0;
  }
  return on == 1;
}
void timing_begin(cudaStream_t stream) {
  t_ev0 = nullptr;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return;
  if (cudaEventCreate(&t_ev0) != cudaSuccess) { t_ev0 = nullptr; return; }
  cudaEventRecord(t_ev0, stream);
}
void timing_end(cudaStream_t stream, const void* func, dim3 grid, dim3 block) {
  if (!t_ev0) return;
  TimedLaunch t{t_ev0, nullptr, func, grid, block};
  if (cudaEventCreate(&t.e1) != cudaSuccess) return;
  cudaEventRecord(t.e1, stream);
  g_timed.push_back(t);
  t_ev0 = nullptr;
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* map, const void* gptr, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                      uint32_t box_cols, uint32_t box_rows) {
  auto fn = get_encode_fn();
  SGG_CHECK(fn != nullptr, "cuTensorMapEncodeTiled driver entry point not available (no CUDA driver?)");
  SGG_CHECK((reinterpret_cast<uintptr_t>(gptr) & 15) == 0, "TMA base pointer %p not 16-byte aligned", gptr);
  SGG_CHECK((ld_elems * 2) % 16 == 0, "TMA row pitch %llu elements not a multiple of 16 bytes",
            (unsigned long long)ld_elems);
  SGG_CHECK(box_cols * 2 <= 128 && box_rows <= 256, "TMA box %ux%u too large", box_cols, box_rows);
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(gptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SGG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): rows=%llu cols=%llu ld=%llu box=%ux%u", (int)r,
            (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems, box_cols, box_rows);
  return 0;
}

// 3-D bf16 tensor [d2][d1][d0] (d0 contiguous, dense), 128B swizzle, box {box0 (<=64), box1, 1}.  Out-of-bounds rows of
// a box (d1 beyond the extent) are zero-filled without touching memory: used for the per-sample annotation tiles.
int make_tmap_bf16_3d(CUtensorMap* map, const void* gptr, uint64_t d2, uint64_t d1, uint64_t d0, uint32_t box0,
                      uint32_t box1) {
  auto fn = get_encode_fn();
  SGG_CHECK(fn != nullptr, "cuTensorMapEncodeTiled driver entry point not available (no CUDA driver?)");
  SGG_CHECK((reinterpret_cast<uintptr_t>(gptr) & 15) == 0, "TMA base pointer %p not 16-byte aligned", gptr);
  SGG_CHECK((d0 * 2) % 16 == 0 && box0 * 2 <= 128 && box1 <= 256, "TMA 3-D box %ux%u / row %llu unsupported", box0, box1,
            (unsigned long long)d0);
  cuuint64_t gdim[3] = {d0, d1, d2};
  cuuint64_t gstride[2] = {d0 * 2, d0 * d1 * 2};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(gptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SGG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (3-D) failed (%d): dims=%llux%llux%llu box=%ux%u", (int)r,
            (unsigned long long)d2, (unsigned long long)d1, (unsigned long long)d0, box0, box1);
  return 0;
}

}  // namespace sgg

extern "C" const char* sgg_last_error(void) { return sgg::g_err; }
extern "C" int sgg_version(void) { return 100; }
extern "C" int64_t sgg_launch_count(void) { return (int64_t)sgg::launch_count(); }

// "kernel name;launches" lines for every kernel of this library launched (or captured) so far; reset != 0 clears the
// counters afterwards.  Returns the number of bytes needed.
extern "C" int64_t sgg_kernel_counts(char* buf, int64_t cap, int32_t reset) {
  using namespace sgg;
  std::string out;
  {
    std::lock_guard<std::mutex> lk(g_kern_mu);
    for (auto& kv : g_kern_counts) {
      out += kernel_name(kv.first) + ";" + std::to_string(kv.second) + "\n";
    }
    if (reset) g_kern_counts.clear();
  }
  if (buf && cap > 0) {
    const size_t n = out.size() < (size_t)cap - 1 ? out.size() : (size_t)cap - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return (int64_t)out.size() + 1;
}

// Synchronises the device and writes "kernel;grid;block;launches;total_us" lines for every launch timed since the
// last report (SGG_TIMING=1) into buf; returns the number of bytes needed.
extern "C" int64_t sgg_timing_report(char* buf, int64_t cap) {
  using namespace sgg;
  cudaDeviceSynchronize();
  struct Agg { long long n = 0; double us = 0; };
  std::map<std::string, Agg> agg;
  std::vector<std::string> order;
  for (auto& t : g_timed) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, t.e0, t.e1);
    char key[768];
    snprintf(key, sizeof(key), "%s;%ux%ux%u;%u", kernel_name(t.func).c_str(), t.grid.x, t.grid.y, t.grid.z, t.block.x);
    auto it = agg.find(key);
    if (it == agg.end()) { order.push_back(key); it = agg.emplace(key, Agg{}).first; }
    it->second.n += 1;
    it->second.us += ms * 1e3;
    cudaEventDestroy(t.e0);
    cudaEventDestroy(t.e1);
  }
  g_timed.clear();
  std::string out;
  for (auto& k : order) {
    char line[896];
    snprintf(line, sizeof(line), "%s;%lld;%.2f\n", k.c_str(), agg[k].n, agg[k].us);
    out += line;
  }
  if (buf && cap > 0) {
    const size_t n = out.size() < (size_t)cap - 1 ? out.size() : (size_t)cap - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return (int64_t)out.size() + 1;
}

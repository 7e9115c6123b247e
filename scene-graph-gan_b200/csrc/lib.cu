// Host-side plumbing of libsgg_b200: error string, version, TMA tensor-map encoding.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cudaTypedefs.h>

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

}  // namespace sgg

extern "C" const char* sgg_last_error(void) { return sgg::g_err; }
extern "C" int sgg_version(void) { return 100; }
extern "C" int64_t sgg_launch_count(void) { return (int64_t)sgg::launch_count(); }

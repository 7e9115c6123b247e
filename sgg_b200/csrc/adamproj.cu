// Optimiser-fused attention projection: one pass over the annotation rows W_a of attention_perceptron/kernel that
//   (1) applies the TF-form Adam update (train:258-259) to W_a, and
//   (2) feeds the freshly updated weights -- as a bf16 hi/lo pair written straight into the swizzled shared-memory
//       B operand -- to tcgen05 MMAs against the annotation tile, producing the NEXT pass's hoisted projection
//       P = flat(a) W_a (gen:14-15) by split-K reduction.
// The separate K1 GEMM would stream the 79 MB hi/lo shadow of W_a back from HBM right after Adam wrote it; here the
// weights never leave the SM between the update and the contraction, and the shadow is only written when a later
// pass needs it (new annotations).  Single GPU (or any case with M <= 256 rows and complete local gradients).
#include "common.cuh"
#include "../../include/sgg_b200.h"

namespace sgg {

constexpr int AP_THREADS = 448;                  // warp 0: TMA, warp 1: MMA, warps 2..13: Adam (2..9 also the epilogue)
constexpr int AP_WORKERS = 384;
constexpr int AP_A_STAGE = 2 * 128 * 64 * 2;     // two 128-row m-tiles x 64 k, bf16: 32 KB
constexpr int AP_B_PART = 256 * 64 * 2;          // 64 k x 256 n (4 chunks of 64 n), bf16: 32 KB
constexpr int AP_SMEM = 2 * AP_A_STAGE + 2 * 2 * AP_B_PART + 1024 + 256;

struct AdamProjParams {
  float* theta; const float* grad; float* m; float* v;   // the W_a block: [K, R] row-major fp32
  __nv_bfloat16* shadow; int pitch; long long lo_off; int write_shadow;
  float lr, b1, b2, eps;
  float lr_t;                                            // used when iter == nullptr (host-side step number)
  const long long* iter; long long step_mul, step_add;
  int R, total_kb, M;
  float* P; long long ldP;                               // [M, ldP] fp32, zero-filled by the caller
};

__global__ void __launch_bounds__(AP_THREADS, 1)
adam_proj_kernel(const __grid_constant__ CUtensorMap tmA, const AdamProjParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                               // [2][AP_A_STAGE]
  uint8_t* sB = smem + 2 * AP_A_STAGE;              // [2][hi | lo][AP_B_PART]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 4 * AP_B_PART);
  uint64_t* a_full = bars;          // [2]
  uint64_t* a_empty = bars + 2;     // [2]
  uint64_t* b_ready = bars + 4;     // [2]
  uint64_t* b_free = bars + 6;      // [2]
  uint64_t* tmem_full = bars + 8;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 9);
  float* s_lr = reinterpret_cast<float*>(bars + 10);

  const int warp = threadIdx.x >> 5;
  const int nsplit = gridDim.x;
  const int kb_per = (p.total_kb + nsplit - 1) / nsplit;
  const int kb_begin = blockIdx.x * kb_per;
  const int kb_end = min(p.total_kb, kb_begin + kb_per);
  const int nkb = max(0, kb_end - kb_begin);

  pdl_trigger();
  if (warp == 0 && elect_one()) tma_prefetch_desc(&tmA);
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < 2; ++s) {
        mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1);
        mbar_init(&b_ready[s], AP_WORKERS); mbar_init(&b_free[s], 1);
      }
      mbar_init(tmem_full, 1);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  if (warp >= 2) {   // clear both B buffers once: columns >= R of the MMA's N range are never written afterwards
    const int t = threadIdx.x - 64;
    uint4* z = reinterpret_cast<uint4*>(sB);
    for (int i = t; i < 4 * AP_B_PART / 16; i += AP_WORKERS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer: annotation tile [2 x 128 rows, 64 k] per k-block (an input of the step) =====
    if (elect_one()) {
      for (int i = 0; i < nkb; ++i) {
        const int st = i & 1;
        mbar_wait(&a_empty[st], ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(&a_full[st], AP_A_STAGE);
        const int kc = (kb_begin + i) * 64;
        tma_load_2d(sA + st * AP_A_STAGE, &tmA, &a_full[st], kc, 0);
        tma_load_2d(sA + st * AP_A_STAGE + 128 * 64 * 2, &tmA, &a_full[st], kc, 128);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc_bf16(128, (p.R + 15) & ~15, false, true);
    for (int i = 0; i < nkb; ++i) {
      const int st = i & 1;
      const uint32_t ph = (i >> 1) & 1;
      mbar_wait(&a_full[st], ph);
      mbar_wait(&b_ready[st], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a0 = smem_u32(sA + st * AP_A_STAGE);
        const uint32_t bh = smem_u32(sB + st * 2 * AP_B_PART), bl = bh + AP_B_PART;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t dbh = make_smem_desc(bh + k * 2048, 64 * 128, 1024);
          const uint64_t dbl = make_smem_desc(bl + k * 2048, 64 * 128, 1024);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            const uint64_t da = make_smem_desc(a0 + mt * (128 * 64 * 2) + k * 32, 0, 1024);
            umma_bf16(tmem_base + mt * 256, da, dbh, idesc, (i | k) ? 1u : 0u);
            umma_bf16(tmem_base + mt * 256, da, dbl, idesc, 1u);
          }
        }
        umma_commit(&a_empty[st]);
        umma_commit(&b_free[st]);
        if (i == nkb - 1) umma_commit(tmem_full);
      }
      __syncwarp();
    }
  } else {
    // ===================== Adam on 64 rows of W_a per k-block, result into the B operand =====================
    const int t = threadIdx.x - 64;
    pdl_wait();   // the gradient comes from the preceding kernels
    if (t == 0) {
      float lr_t = p.lr_t;
      if (p.iter) {
        const double step = (double)(p.iter[0] * p.step_mul + p.step_add);
        lr_t = (float)((double)p.lr * sqrt(1.0 - pow((double)p.b2, step)) / (1.0 - pow((double)p.b1, step)));
      }
      *s_lr = lr_t;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(AP_WORKERS) : "memory");
    const float lr_t = *s_lr;
    const unsigned R = (unsigned)p.R;
    const int n_f4 = 64 * p.R / 4;                 // float4 per k-block (R % 4 == 0: a float4 never straddles a row)
    const long long range_end = (long long)kb_end * 64 * p.R;
    for (int i = 0; i < nkb; ++i) {
      const int st = i & 1;
      mbar_wait(&b_free[st], ((i >> 1) & 1) ^ 1);
      uint8_t* bhi = sB + st * 2 * AP_B_PART;
      const long long row0 = (long long)(kb_begin + i) * 64;
      const long long base = row0 * p.R;
#pragma unroll 1
      for (int f0 = t; f0 < n_f4; f0 += 4 * AP_WORKERS) {
        float4 th[4], g[4], mm[4], vv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int f = f0 + u * AP_WORKERS;
          if (f < n_f4) {
            const long long gi = base + 4LL * f;
            // the CTA's rows are one contiguous run: pull the lines two passes ahead into L2 (more bytes in flight than
            // the registers of 256 threads can hold)
            const long long gp = gi + 2LL * 4 * 4 * AP_WORKERS;
            if (gp < range_end && (gi & 31) == 0) {   // one request per 128-byte line
              prefetch_l2(p.theta + gp); prefetch_l2(p.grad + gp); prefetch_l2(p.m + gp); prefetch_l2(p.v + gp);
            }
            th[u] = *reinterpret_cast<const float4*>(p.theta + gi);
            g[u] = __ldcs(reinterpret_cast<const float4*>(p.grad + gi));
            mm[u] = *reinterpret_cast<const float4*>(p.m + gi);
            vv[u] = *reinterpret_cast<const float4*>(p.v + gi);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int f = f0 + u * AP_WORKERS;
          if (f >= n_f4) continue;
          const long long gi = base + 4LL * f;
          float tv[4] = {th[u].x, th[u].y, th[u].z, th[u].w};
          const float gv[4] = {g[u].x, g[u].y, g[u].z, g[u].w};
          float mv[4] = {mm[u].x, mm[u].y, mm[u].z, mm[u].w};
          float v4[4] = {vv[u].x, vv[u].y, vv[u].z, vv[u].w};
          __nv_bfloat16 h[4], l[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            mv[e] = p.b1 * mv[e] + (1.0f - p.b1) * gv[e];
            v4[e] = p.b2 * v4[e] + (1.0f - p.b2) * gv[e] * gv[e];
            tv[e] -= lr_t * mv[e] / (sqrtf(v4[e]) + p.eps);
            split_bf16(tv[e], h[e], l[e]);
          }
          *reinterpret_cast<float4*>(p.theta + gi) = make_float4(tv[0], tv[1], tv[2], tv[3]);
          *reinterpret_cast<float4*>(p.m + gi) = make_float4(mv[0], mv[1], mv[2], mv[3]);
          *reinterpret_cast<float4*>(p.v + gi) = make_float4(v4[0], v4[1], v4[2], v4[3]);
          const unsigned e0 = 4u * (unsigned)f;
          const unsigned r = e0 / R, c = e0 - r * R;            // k row within the block, output column n
          // MN-major, 128B-swizzled operand: chunk of 64 n = 64 k-rows x 128 B; 16-byte unit index XOR (k & 7)
          const uint32_t off = (c >> 6) * 8192u + r * 128u + ((((c & 63u) >> 3) ^ (r & 7u)) << 4) + (c & 7u) * 2u;
          const uint2 hv = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
          const uint2 lv = make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
          *reinterpret_cast<uint2*>(bhi + off) = hv;
          *reinterpret_cast<uint2*>(bhi + AP_B_PART + off) = lv;
          if (p.write_shadow) {
            __nv_bfloat16* dst = p.shadow + (row0 + r) * p.pitch + c;
            *reinterpret_cast<uint2*>(dst) = hv;
            *reinterpret_cast<uint2*>(dst + p.lo_off) = lv;
          }
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&b_ready[st]);
    }
    // ===================== epilogue: split-K reduction of the two accumulators into P =====================
    if (nkb > 0 && warp < 10) {
      const int mt = (warp - 2) >> 2, q = warp & 3;
      const int row = mt * 128 + q * 32 + (int)lane_id();
      mbar_wait(tmem_full, 0);
      tc_fence_after();
      const int nch = (p.R + 31) / 32;
      const int rot = (int)(blockIdx.x % (unsigned)nch);
#pragma unroll 1
      for (int ci = 0; ci < nch; ++ci) {
        const int c = ci + rot < nch ? ci + rot : ci + rot - nch;
        uint32_t r32[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * 256 + c * 32), r32);
        tmem_ld_wait();
        if (row >= p.M) continue;
        float* prow = p.P + (long long)row * p.ldP + c * 32;
        const bool full = c * 32 + 32 <= p.R;
        if (full && ((p.ldP & 3) == 0)) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(prow + j), "f"(__uint_as_float(r32[j])),
                         "f"(__uint_as_float(r32[j + 1])), "f"(__uint_as_float(r32[j + 2])), "f"(__uint_as_float(r32[j + 3]))
                         : "memory");
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c * 32 + j < p.R) atomicAdd(prow + j, __uint_as_float(r32[j]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ann: [M, K] bf16 row-major (K = total_kb * 64).  P must be zero-filled (stream-ordered) before this launch.
int adam_proj(const AdamProjParams& p, const void* ann, cudaStream_t stream) {
  SGG_CHECK(p.M >= 1 && p.M <= 256 && p.R >= 4 && p.R <= 256 && (p.R & 3) == 0 && p.total_kb >= 1,
            "adam_proj: unsupported shape (M=%d R=%d)", p.M, p.R);
  static bool configured = false;
  if (!configured) {
    SGG_CUDA(cudaFuncSetAttribute(adam_proj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AP_SMEM));
    configured = true;
  }
  CUtensorMap tmA;
  const long long K = (long long)p.total_kb * 64;
  SGG_TRY(make_tmap_bf16_2d(&tmA, ann, (uint64_t)p.M, (uint64_t)K, (uint64_t)K, 64, 128));
  const int grid = p.total_kb < 148 ? p.total_kb : 148;
  SGG_LAUNCH(adam_proj_kernel, grid, AP_THREADS, AP_SMEM, stream, tmA, p);
  return 0;
}

}  // namespace sgg

// tcgen05 GEMM for sm_100a: TMA-staged bf16 operands (128B swizzle), fp32 accumulators in
// TMEM, warp-specialised (1 TMA warp, 1 MMA warp, 4 epilogue warps), split-K over grid.z.
//
// Replaces every tf MatMul of the reference hot path (gen:15,79/87,88; disc:87,90) and the
// MatMul gradients TF registers for them.  See include/sgg_b200.h (sgg_gemm_desc_t).
#include <cmath>

#include "common.cuh"
#include "../../include/sgg_b200.h"

namespace sgg {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 192;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB

// Zero-fill as a KERNEL (not a memset node): it takes part in the programmatic-dependent-launch chain, so the launch
// of the kernel behind it still overlaps, which a memset node in the middle of the chain prevents.  Up to 12 pitched
// regions per launch; widths / pitches / pointers in multiples of 16 bytes.
constexpr int ZERO_MAX_SEG = 12;
struct ZeroSegs { void* p[ZERO_MAX_SEG]; long long pitch16[ZERO_MAX_SEG], width16[ZERO_MAX_SEG], rows[ZERO_MAX_SEG]; long long start[ZERO_MAX_SEG + 1]; int n; };
__global__ void __launch_bounds__(256) zero_kernel(const ZeroSegs z) {
  pdl_trigger();
  pdl_wait();
  const long long total = z.start[z.n];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int s = 0;
    while (s + 1 < z.n && i >= z.start[s + 1]) ++s;
    const long long j = i - z.start[s];
    const long long r = j / z.width16[s], c = j - r * z.width16[s];
    reinterpret_cast<uint4*>(z.p[s])[r * z.pitch16[s] + c] = make_uint4(0u, 0u, 0u, 0u);
  }
}
struct ZeroList {
  ZeroSegs z{};
  // returns false when the region cannot be expressed in 16-byte units (caller falls back to cudaMemset*Async)
  bool add(void* p, long long pitch_bytes, long long width_bytes, long long rows) {
    if (z.n >= ZERO_MAX_SEG || (reinterpret_cast<uintptr_t>(p) & 15) || (pitch_bytes & 15) || (width_bytes & 15)) return false;
    z.p[z.n] = p; z.pitch16[z.n] = pitch_bytes / 16; z.width16[z.n] = width_bytes / 16; z.rows[z.n] = rows;
    z.start[z.n + 1] = z.start[z.n] + rows * (width_bytes / 16);
    ++z.n;
    return true;
  }
  bool add(void* p, long long bytes) { return add(p, bytes, bytes, 1); }
};
int zero_fill(const ZeroList& l, cudaStream_t stream) {
  const long long total = l.z.start[l.z.n];
  if (l.z.n == 0 || total <= 0) return 0;
  const long long want = (total + 255) / 256;
  const int grid = (int)(want < 148 * 8 ? want : 148 * 8);
  SGG_LAUNCH(zero_kernel, grid, 256, 0, stream, l.z);
  return 0;
}
// zero one pitched fp32 region, as a kernel when it can be (see zero_kernel), else as a memset node
int zero_2d(float* p, long long ld, long long cols, long long rows, cudaStream_t stream) {
  ZeroList l;
  if (l.add(p, ld * 4, cols * 4, rows)) return zero_fill(l, stream);
  SGG_CUDA(cudaMemset2DAsync(p, (size_t)ld * 4, 0, (size_t)cols * 4, (size_t)rows, stream));
  return 0;
}

struct GemmKParams {
  int M, N;
  // Operand parts.  x ~= hi + lo operands are given as two parts of the same tensor (different k / mn offsets);
  // a k-block loads every part ONCE and issues the products (A0,B0), (A1,B0), (A0,B1) into one accumulator.
  int nA, nB;
  int a_k[2], a_mn[2], b_k[2], b_mn[2];
  int total_kb;                  // k-blocks of 64
  int stages, stage_bytes;
  float* C; long long ldc; int atomic;
  __nv_bfloat16* Chl; long long ld_hl; long long lo_off;
  const float* bias;
  const float* addm; long long ld_addm; int add_mod;
  float alpha;
  int rm_d0, rm_d1; long long rm_s0, rm_s1;   // output row permutation (rm_d0 == 0: identity)
  // sampling epilogue (train:270 argmax / Gumbel-max): per output row the running maximum of v[n] (+ Gumbel noise) and
  // its column, packed as (orderable float bits << 32) | ~column and merged across n-tiles with a 64-bit atomicMax
  int b_stream;   // B operand is a read-once weight stream (W_a): L2 evict_first; fp32 output stores stream too
  unsigned long long* amax; long long amax_stride;
  int gumbel; unsigned long long gseed, goff;
  int dense_out;  // plain fp32 output whose rows are contiguous (ldc == N): staged through shared memory, coalesced stores
};

constexpr int GEMM_MAX_STAGES = 8;
constexpr int GEMM_SMEM_BUDGET = 200 * 1024;   // operand ring; + 1 KB alignment slack + barriers

// MT = m-tiles (of 128 rows) per CTA.  MT = 2 keeps two accumulators (2 x BN TMEM columns) so that a wide B tile is
// fetched once per 256 output rows: used for the W_a contractions, whose shared-memory fill is dominated by B.
template <int BN, bool A_MN, bool B_MN, int MT = 1>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const GemmKParams p) {
  constexpr int B_PART_BYTES = BN * BK * 2;
  constexpr int A_PART_BYTES = MT * A_STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes);
  uint64_t* empty_bar = full_bar + GEMM_MAX_STAGES;
  uint64_t* tmem_full_bar = empty_bar + GEMM_MAX_STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int m0 = blockIdx.x * (BM * MT);
  const int n0 = blockIdx.y * BN;
  // split-K range for this CTA
  const int splits = gridDim.z;
  const int kb_per = (p.total_kb + splits - 1) / splits;
  const int kb_begin = blockIdx.z * kb_per;
  const int kb_end = min(p.total_kb, kb_begin + kb_per);
  const int nkb = max(0, kb_end - kb_begin);
  const int b_off = p.nA * A_PART_BYTES;   // B parts follow the A parts inside a stage

  pdl_trigger();
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < p.stages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      mbar_init(tmem_full_bar, 1);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, MT * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();   // barriers, TMEM and descriptors are set up; operands / outputs belong to the preceding kernels

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint64_t pol_first = l2_policy_evict_first();
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = smem + stage * p.stage_bytes;
        mbar_expect_tx(&full_bar[stage], p.stage_bytes);
        const int kc = (kb_begin + i) * BK;
        for (int pa = 0; pa < p.nA; ++pa) {
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            uint8_t* sA = st + pa * A_PART_BYTES + mt * A_STAGE_BYTES;
            if (!A_MN) {
              tma_load_2d(sA, &tmA, &full_bar[stage], p.a_k[pa] + kc, m0 + mt * BM + p.a_mn[pa]);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j)
                tma_load_2d(sA + j * (BK * 128), &tmA, &full_bar[stage], m0 + mt * BM + p.a_mn[pa] + 64 * j, p.a_k[pa] + kc);
            }
          }
        }
        for (int pb = 0; pb < p.nB; ++pb) {
          uint8_t* sB = st + b_off + pb * B_PART_BYTES;
          if (!B_MN) {
            tma_load_2d(sB, &tmB, &full_bar[stage], p.b_k[pb] + kc, n0 + p.b_mn[pb]);
          } else if (MT == 2 && p.b_stream) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d_hint(sB + j * (BK * 128), &tmB, &full_bar[stage], n0 + p.b_mn[pb] + 64 * j, p.b_k[pb] + kc, pol_first);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(sB + j * (BK * 128), &tmB, &full_bar[stage], n0 + p.b_mn[pb] + 64 * j, p.b_k[pb] + kc);
          }
        }
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // the MMA spans only the valid columns of a ragged last n-tile (N rounded up to the UMMA granule of 16)
    const int n_valid = p.N - n0;
    const uint32_t idesc = make_idesc_bf16(BM, n_valid >= BN ? BN : ((n_valid + 15) & ~15), A_MN, B_MN);
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < nkb; ++i) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sA0 = smem_u32(smem + stage * p.stage_bytes);
        const uint32_t sB0 = sA0 + b_off;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint32_t ao = A_MN ? k * 2048 : k * 32, bo = B_MN ? k * 2048 : k * 32;
          const uint64_t db0 = B_MN ? make_smem_desc(sB0 + bo, BK * 128, 1024) : make_smem_desc(sB0 + bo, 0, 1024);
          const uint32_t sB1 = sB0 + B_PART_BYTES;
          const uint64_t db1 = B_MN ? make_smem_desc(sB1 + bo, BK * 128, 1024) : make_smem_desc(sB1 + bo, 0, 1024);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint32_t sAm = sA0 + mt * A_STAGE_BYTES;
            const uint32_t acc = tmem_base + mt * BN;
            const uint64_t da0 = A_MN ? make_smem_desc(sAm + ao, BK * 128, 1024) : make_smem_desc(sAm + ao, 0, 1024);
            umma_bf16(acc, da0, db0, idesc, (i | k) ? 1u : 0u);
            if (p.nA == 2) {
              const uint32_t sA1 = sAm + A_PART_BYTES;
              const uint64_t da1 = A_MN ? make_smem_desc(sA1 + ao, BK * 128, 1024) : make_smem_desc(sA1 + ao, 0, 1024);
              umma_bf16(acc, da1, db0, idesc, 1u);
            }
            if (p.nB == 2) umma_bf16(acc, da0, db1, idesc, 1u);
          }
        }
        umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
        if (i == nkb - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (nkb > 0) {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
#pragma unroll 1
   for (int mi = 0; mi < MT; ++mi) {
    const int mt = (MT == 2 && gridDim.z > 1) ? (mi ^ (int)((blockIdx.z >> 3) & 1u)) : mi;   // see `rot` below
    const int row = m0 + mt * BM + q * 32 + (int)lane_id();
    const bool row_ok = row < p.M;
    long long orow = row;
    if (p.rm_d0 > 0) {
      const int rem = row % p.rm_d0;
      orow = (long long)(row / p.rm_d0) * p.rm_s0 + (long long)(rem / p.rm_d1) * p.rm_s1 + (rem % p.rm_d1);
    }
    const float* addrow = p.addm ? p.addm + (long long)(row % p.add_mod) * p.ld_addm : nullptr;
    if (addrow && blockIdx.z == 0) {   // broadcast-add operand: into L1 while the main loop runs
      for (int c = 0; c < BN / 32; ++c)
        if (n0 + c * 32 < p.N) prefetch_l1(addrow + n0 + c * 32);
    }
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (p.dense_out) {
      // The m-tile's output rows are ONE contiguous block of global memory (ldc == N, e.g. dW_a [R*C, R]).  A thread owns
      // a row in TMEM, so direct stores would touch 32 partial sectors per instruction; instead the rows are staged in the
      // (measured: dW_a 72.9 -> 64.5 us.  The same staging for the split-K reductions / strided outputs of the other GEMMs
      // was slower than their direct row-per-thread red.add epilogue: 6.05 vs 5.35 ms per iteration, so they keep it.)
      // (now idle) operand ring with pitch N and the four epilogue warps stream the block out with coalesced float4 stores.
      float* stg = reinterpret_cast<float*>(smem);
      const int rloc = q * 32 + (int)lane_id();
      const int nch = (p.N - n0 + 31) / 32;
      for (int c = 0; c < nch; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * BN + c * 32), r);
        tmem_ld_wait();
        float* dst = stg + (long long)rloc * p.N + c * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          if (c * 32 + j < p.N)
            *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(r[j]) * p.alpha, __uint_as_float(r[j + 1]) * p.alpha,
                                                              __uint_as_float(r[j + 2]) * p.alpha, __uint_as_float(r[j + 3]) * p.alpha);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int rows_valid = min(BM, p.M - (m0 + mt * BM));
      if (rows_valid > 0) {
        const int n4 = rows_valid * p.N / 4;
        float4* out = reinterpret_cast<float4*>(p.C + (long long)(m0 + mt * BM) * p.N);
        const float4* src = reinterpret_cast<const float4*>(stg);
        for (int i = (int)threadIdx.x - 64; i < n4; i += 128) out[i] = src[i];
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");   // the staging block is rewritten by the next m-tile
      continue;
    }
    float best_v = -INFINITY;
    int best_c = 0;
    // split-K: the CTAs that share an output tile finish together and would all reduce into the same addresses in the
    // same order; each starts at its own column chunk so that same-address reductions are spread out in time
    const int nch = min(BN / 32, (p.N - n0 + 31) / 32);
    const int rot = gridDim.z > 1 ? (int)(blockIdx.z % (unsigned)nch) : 0;
#pragma unroll 1
    for (int ci = 0; ci < nch; ++ci) {
      const int c = ci + rot < nch ? ci + rot : ci + rot - nch;
      const int col0 = n0 + c * 32;
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * BN + c * 32), r);
      tmem_ld_wait();
      if (!row_ok) continue;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
      const bool full = (col0 + 32 <= p.N);
      // bias / broadcast-add operands: 16-byte loads where the row is aligned (a thread owns a row, so every scalar
      // load instruction would touch 32 different lines; the scores GEMM's epilogue was dominated by 64 of them)
      const bool vec4 = ((p.N & 3) == 0);
      if (p.bias && blockIdx.z == 0) {
        if (vec4 && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (col0 + j < p.N) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (full || col0 + j < p.N) v[j] += __ldg(p.bias + col0 + j);
        }
      }
      if (addrow && blockIdx.z == 0) {
        if (vec4 && (p.ld_addm & 3) == 0 && (reinterpret_cast<uintptr_t>(p.addm) & 15) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (col0 + j < p.N) {
              const float4 a4 = __ldg(reinterpret_cast<const float4*>(addrow + col0 + j));
              v[j] += a4.x; v[j + 1] += a4.y; v[j + 2] += a4.z; v[j + 3] += a4.w;
            }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (full || col0 + j < p.N) v[j] += __ldg(addrow + col0 + j);
        }
      }
      if (p.amax) {
        float g[32];
        if (p.gumbel) {   // g = -log(-log u), u = Philox(seed; row, column quad): the draw of element (row, n) is fixed
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const unsigned long long ctr = p.goff + (unsigned long long)orow;
            const uint4 rr = philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)((col0 + j) >> 2), 0x9E37u),
                                           make_uint2((uint32_t)p.gseed, (uint32_t)(p.gseed >> 32)));
            g[j] = -__logf(-__logf(u01(rr.x))); g[j + 1] = -__logf(-__logf(u01(rr.y)));
            g[j + 2] = -__logf(-__logf(u01(rr.z))); g[j + 3] = -__logf(-__logf(u01(rr.w)));
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (full || col0 + j < p.N) {
            const float s = p.gumbel ? v[j] + g[j] : v[j];
            if (s > best_v) { best_v = s; best_c = col0 + j; }   // strict: the lowest column wins a tie (tf.argmax)
          }
        }
      }
      if (p.C) {
        float* crow = p.C + orow * p.ldc + col0;
        if (p.atomic) {
          if (full && ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)   // 16-byte vector reduction (sm_90+): 4x fewer L2 atomic operations
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(crow + j), "f"(v[j]), "f"(v[j + 1]),
                           "f"(v[j + 2]), "f"(v[j + 3])
                           : "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (full || col0 + j < p.N) atomicAdd(crow + j, v[j]);
          }
        } else if (full && ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(crow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (full || col0 + j < p.N) crow[j] = v[j];
        }
      }
      if (p.Chl) {
        __nv_bfloat16* hrow = p.Chl + orow * p.ld_hl + col0;
        __nv_bfloat16* lrow = hrow + p.lo_off;
        if (full && ((p.ld_hl & 7) == 0) && ((p.lo_off & 7) == 0) &&
            ((reinterpret_cast<uintptr_t>(p.Chl) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint32_t h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              __nv_bfloat16 h0, l0, h1, l1;
              split_bf16(v[j + 2 * e], h0, l0);
              split_bf16(v[j + 2 * e + 1], h1, l1);
              h[e] = pack_bf16x2(h0, h1);
              l[e] = pack_bf16x2(l0, l1);
            }
            *reinterpret_cast<uint4*>(hrow + j) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(lrow + j) = make_uint4(l[0], l[1], l[2], l[3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (full || col0 + j < p.N) {
              __nv_bfloat16 h0, l0;
              split_bf16(v[j], h0, l0);
              hrow[j] = h0;
              lrow[j] = l0;
            }
        }
      }
    }
    if (p.amax && row_ok && n0 < p.N) {
      uint32_t b = __float_as_uint(best_v);
      b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
      atomicMax(p.amax + orow * p.amax_stride, ((unsigned long long)b << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)best_c));
    }
   }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, MT * BN);
}

static int gemm_smem_bytes() { return GEMM_SMEM_BUDGET + 1024 + 256; }

template <int BN, bool A_MN, bool B_MN, int MT = 1>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKParams& kp, int splits,
                       cudaStream_t stream) {
  auto kern = gemm_kernel<BN, A_MN, B_MN, MT>;
  static bool configured = false;
  if (!configured) {
    SGG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_bytes()));
    configured = true;
  }
  dim3 grid((kp.M + BM * MT - 1) / (BM * MT), (kp.N + BN - 1) / BN, splits);
  int smem = kp.stages * kp.stage_bytes + 1024 + 256;
  {  // A single-wave grid runs faster with one CTA per SM (measured: two co-resident small-tile CTAs share the SM's
     // operand feed while other SMs idle; scores GEMM 14.2 -> 9.7 us): pad the shared-memory request past half an SM.
     // SGG_GEMM_MIN_SMEM (bytes) overrides the padding (0 disables it).
    static int min_smem = -1;
    if (min_smem < 0) { const char* e = getenv("SGG_GEMM_MIN_SMEM"); min_smem = e ? atoi(e) : 118000; }
    const long long ctas = (long long)grid.x * grid.y * grid.z;
    if (ctas <= 148 && smem < min_smem) smem = min_smem < gemm_smem_bytes() ? min_smem : gemm_smem_bytes();
  }
  SGG_LAUNCH(kern, grid, GEMM_THREADS, smem, stream, tmA, tmB, kp);
  return 0;
}

template <int BN>
static int dispatch_major(bool a_mn, bool b_mn, const CUtensorMap& tmA, const CUtensorMap& tmB,
                          const GemmKParams& kp, int splits, cudaStream_t stream, int mt = 1) {
  if (BN == 256 && mt == 2) {   // the two W_a contractions: K1 (A K-major, B MN-major) and dW_a (both MN-major)
    if (!a_mn && b_mn) return launch_gemm<256, false, true, 2>(tmA, tmB, kp, splits, stream);
    if (a_mn && b_mn) return launch_gemm<256, true, true, 2>(tmA, tmB, kp, splits, stream);
  }
  if (!a_mn && !b_mn) return launch_gemm<BN, false, false>(tmA, tmB, kp, splits, stream);
  if (!a_mn && b_mn) return launch_gemm<BN, false, true>(tmA, tmB, kp, splits, stream);
  if (a_mn && b_mn) return launch_gemm<BN, true, true>(tmA, tmB, kp, splits, stream);
  return launch_gemm<BN, true, false>(tmA, tmB, kp, splits, stream);
}

constexpr int GEMM_MAX_CHAIN_KB = 128;   // longest accumulation chain (64-wide k-blocks) of an automatically split GEMM

// sgg_gemm_plan: when set, gemm_fused reports its (tile width, split-K, m-tiles per CTA) choice here and launches nothing
struct GemmPlan { int block_n, splits, m_tiles; };
static thread_local GemmPlan* t_plan_out = nullptr;

// One fused launch: parts (nA, nB) as described at GemmKParams.
static int gemm_fused(const sgg_gemm_desc_t& d, GemmKParams kp, cudaStream_t stream) {
  // ---- tile width and split-K.  These GEMMs are small (M = a few hundred rows) and long in K, so a plain tile grid
  // leaves most of the 148 SMs idle.  Unless the caller fixed them, pick (block_n, splits) minimising a simple
  // cost: waves x (k-blocks per CTA x max(stage bytes / L2 feed rate, MMA cycles)).  A split-K launch accumulates
  // with fp32 vector reductions into an output that is zeroed first (or, when the caller asked for `atomic`, into
  // whatever the output already holds).
  const int tm = (d.M + BM - 1) / BM;
  const int nprod = kp.nA + kp.nB - 1;
  const bool can_split = d.C && !d.Chl && !d.argmax_keys;
  int bn = d.block_n, splits = d.splits > 1 ? d.splits : 1;
  // B-heavy contractions (single A part, hi/lo B, N <= 256: the two W_a GEMMs): 256-row CTAs with two accumulators, so
  // the wide B tile enters shared memory once per 256 output rows.
  static int no_wide = -1;   // SGG_NO_WIDE_B=1: tuning aid
  if (no_wide < 0) { const char* e = getenv("SGG_NO_WIDE_B"); no_wide = (e && e[0] == '1') ? 1 : 0; }
  const bool wide_b = !no_wide && kp.nA == 1 && kp.nB == 2 && d.b_mn_major && d.M >= 2 * BM && d.N > 128 && d.N <= 256 &&
                      (d.block_n == 0 || d.block_n == 256);
  const int mt = wide_b ? 2 : 1;
  kp.b_stream = (wide_b && !d.a_mn_major && l2_policy_enabled()) ? 1 : 0;   // K1: W_a is read once per pass
  if (wide_b) {
    bn = 256;
    if (d.splits <= 0) {
      const int tiles = (d.M + 2 * BM - 1) / (2 * BM);
      splits = (can_split && tiles < 148) ? 148 / tiles : 1;
      if (splits > kp.total_kb / 2) splits = kp.total_kb / 2 > 0 ? kp.total_kb / 2 : 1;
    }
  } else if (d.splits <= 0 || d.block_n == 0) {
    // Cost model (cycles), constants fitted to graph-replayed timings of the step's GEMM shapes on B200
    // (tools/gemm_micro.py, profiles/): fixed launch + prologue + pipeline fill, per-k-block max(operand feed, MMA issue),
    // epilogue per output column, and a penalty for split-K (output memset node + fp32 reductions).
    double best = 1e30;
    int best_bn = 128, best_sp = 1;
    static double split_penalty_env = -1.0;   // SGG_SPLIT_PENALTY (cycles) overrides the fitted constant (tuning aid)
    if (split_penalty_env < 0.0) { const char* e = getenv("SGG_SPLIT_PENALTY"); split_penalty_env = e ? atof(e) : 0.0; }
    const double split_penalty = split_penalty_env > 0.0 ? split_penalty_env : 6370.0;
    const int bn_lo = d.block_n ? d.block_n : 64, bn_hi = d.block_n ? d.block_n : 256;
    for (int b = bn_lo; b <= bn_hi; b *= 2) {
      if (b > 64 && d.N <= b / 2 && !d.block_n) continue;         // do not pad N by more than 2x
      const int tiles = tm * ((d.N + b - 1) / b);
      const int sp_lo = d.splits > 0 ? splits : 1;
      const int sp_hi = d.splits > 0 ? splits : (can_split ? (tiles < 148 && 148 / tiles > 8 ? 148 / tiles : 8) : 1);
      for (int sp = sp_lo; sp <= sp_hi; ++sp) {
        if (sp > kp.total_kb) break;
        const int kb_per = (kp.total_kb + sp - 1) / sp;
        if (sp > 1 && kb_per < 2) break;
        const int ctas = tiles * sp;
        if (sp > sp_lo && ctas > 3 * 148) break;
        const int waves = (ctas + 147) / 148;
        const int conc = ctas < 148 ? ctas : 148;                  // CTAs competing for the L2 -> SM feed
        const double feed = 13440.0 / conc < 88.0 ? 13440.0 / conc : 88.0;   // bytes / cycle / SM
        const double stage_b = (double)kp.nA * A_STAGE_BYTES + (double)kp.nB * b * BK * 2;
        const double kb_cyc = fmax(stage_b / feed, nprod * 1.93 * b);   // TMA-bound vs MMA-bound k-block
        const double epi = 3140.0 + 69.0 * b * (sp > 1 ? 0.2 : 1.0);
        const double cost = waves * (8885.0 + kb_per * kb_cyc + epi) + (sp > 1 ? split_penalty : 0.0);
        if (cost < best) { best = cost; best_bn = b; best_sp = sp; }
      }
    }
    bn = best_bn;
    splits = best_sp;
  }
  // The tcgen05 fp32 accumulator truncates: over one accumulator's chain the error is a bias that grows linearly with its
  // length (measured, tools/accum_micro.py: 1.2e-9 per contraction index on signed data, 6e-9 on same-sign data -- 1.5e-4 /
  // 8e-4 at K = 128 k), while split-K partial sums are combined by IEEE fp32 reductions in L2.  Unless the caller fixed the
  // split, no accumulator runs over more than GEMM_MAX_CHAIN_KB k-blocks (8 k contraction indices: <= 1e-5 per product).
  if (d.splits <= 0 && can_split && (kp.total_kb + splits - 1) / splits > GEMM_MAX_CHAIN_KB)
    splits = (kp.total_kb + GEMM_MAX_CHAIN_KB - 1) / GEMM_MAX_CHAIN_KB;
  if (splits > kp.total_kb) splits = kp.total_kb;
  SGG_CHECK(splits == 1 || can_split, "sgg_gemm: split-K needs the fp32 output only");
  SGG_CHECK(bn == 64 || bn == 128 || bn == 256, "sgg_gemm: block_n=%d unsupported", bn);
  if (t_plan_out) {   // host-side query only
    *t_plan_out = GemmPlan{bn, splits, mt};
    return 0;
  }
  if (d.atomic == 2) {             // the caller guarantees a zero-filled output: accumulate or overwrite, whichever fits
    kp.atomic = splits > 1 ? 1 : 0;
  } else if (splits > 1 && !d.atomic) {   // overwrite semantics: clear the output, then accumulate
    SGG_CHECK(d.out_d0 == 0, "sgg_gemm: split-K with an output row permutation is not supported");
    SGG_TRY(zero_2d(d.C, d.ldc, d.N, d.M, stream));
    kp.atomic = 1;
  }
  kp.stage_bytes = kp.nA * mt * A_STAGE_BYTES + kp.nB * bn * BK * 2;
  kp.stages = GEMM_SMEM_BUDGET / kp.stage_bytes;
  if (kp.stages > GEMM_MAX_STAGES) kp.stages = GEMM_MAX_STAGES;
  {
    static int cap = -1;   // SGG_GEMM_MAX_STAGES: tuning aid
    if (cap < 0) { const char* e = getenv("SGG_GEMM_MAX_STAGES"); cap = e ? atoi(e) : 0; }
    if (cap > 0 && kp.stages > cap) kp.stages = cap;
  }
  const int kb_per = (kp.total_kb + splits - 1) / splits;
  if (kp.stages > kb_per) kp.stages = kb_per < 1 ? 1 : kb_per;
  // dense contiguous output (see the kernel's dense_out path): one n-tile, no split, plain overwrite, nothing fused, and
  // the staging block must fit inside the operand ring (the barriers behind it stay intact)
  kp.dense_out = (d.C && !d.Chl && !d.bias && !d.addm && !d.argmax_keys && d.out_d0 == 0 && splits == 1 && kp.atomic == 0 &&
                  d.ldc == d.N && d.N <= bn && (d.N & 3) == 0 && (reinterpret_cast<uintptr_t>(d.C) & 15) == 0 &&
                  (long long)BM * d.N * 4 <= (long long)kp.stages * kp.stage_bytes) ? 1 : 0;
  CUtensorMap tmA, tmB;
  // K-major: tensor [MN rows, K cols], box {64 k, tile rows}.  MN-major: tensor [K rows, MN cols], box {64 mn, 64 k}.
  SGG_TRY(make_tmap_bf16_2d(&tmA, d.A, d.a_rows, d.a_cols, d.a_ld, 64, d.a_mn_major ? BK : BM));
  SGG_TRY(make_tmap_bf16_2d(&tmB, d.B, d.b_rows, d.b_cols, d.b_ld, 64, d.b_mn_major ? BK : bn));
  const bool amn = d.a_mn_major != 0, bmn = d.b_mn_major != 0;
  switch (bn) {
    case 64: return dispatch_major<64>(amn, bmn, tmA, tmB, kp, splits, stream);
    case 128: return dispatch_major<128>(amn, bmn, tmA, tmB, kp, splits, stream);
    default: return dispatch_major<256>(amn, bmn, tmA, tmB, kp, splits, stream, mt);
  }
}

int gemm(const sgg_gemm_desc_t& d, cudaStream_t stream) {
  SGG_CHECK(d.A && d.B, "sgg_gemm: null operand");
  SGG_CHECK(d.M > 0 && d.N > 0, "sgg_gemm: bad M/N (%d, %d)", d.M, d.N);
  SGG_CHECK(d.nseg >= 1 && d.nseg <= SGG_GEMM_MAX_SEG, "sgg_gemm: nseg=%d out of range", d.nseg);
  SGG_CHECK(d.C || d.Chl || d.argmax_keys, "sgg_gemm: no output");
  GemmKParams kp{};
  kp.M = d.M;
  kp.N = d.N;
  kp.C = d.C; kp.ldc = d.ldc; kp.atomic = d.atomic;
  kp.Chl = reinterpret_cast<__nv_bfloat16*>(d.Chl); kp.ld_hl = d.ld_hl; kp.lo_off = d.lo_off;
  kp.bias = d.bias;
  kp.addm = d.addm; kp.ld_addm = d.ld_addm; kp.add_mod = d.add_mod > 0 ? d.add_mod : 1;
  kp.alpha = d.alpha;
  kp.rm_d0 = d.out_d0; kp.rm_d1 = d.out_d1 > 0 ? d.out_d1 : 1; kp.rm_s0 = d.out_s0; kp.rm_s1 = d.out_s1;
  kp.amax = reinterpret_cast<unsigned long long*>(d.argmax_keys); kp.amax_stride = d.argmax_stride > 0 ? d.argmax_stride : 1;
  kp.gumbel = d.gumbel; kp.gseed = d.gumbel_seed; kp.goff = d.gumbel_offset;
  SGG_CHECK(!d.argmax_keys || d.splits <= 1, "sgg_gemm: the argmax epilogue cannot be combined with split-K");
  for (int s = 0; s < d.nseg; ++s)
    SGG_CHECK(d.seg_klen[s] > 0, "sgg_gemm: segment %d has length %d", s, d.seg_klen[s]);
  // ---- recognise the hi/lo product pattern: every segment pairs A0 or B0 of segment 0 with at most one other part
  // and all segments have the same length.  (A length that is not a multiple of 64 is only valid when the tail of
  // the last k-block lies outside the tensors or in zero padding: TMA zero-fills out-of-bounds elements.)
  bool fusable = true;
  kp.nA = kp.nB = 1;
  kp.a_k[0] = d.seg_a_k[0]; kp.a_mn[0] = d.seg_a_mn[0]; kp.b_k[0] = d.seg_b_k[0]; kp.b_mn[0] = d.seg_b_mn[0];
  for (int s = 1; s < d.nseg && fusable; ++s) {
    const bool a_same = d.seg_a_k[s] == kp.a_k[0] && d.seg_a_mn[s] == kp.a_mn[0];
    const bool b_same = d.seg_b_k[s] == kp.b_k[0] && d.seg_b_mn[s] == kp.b_mn[0];
    if (d.seg_klen[s] != d.seg_klen[0] || a_same == b_same) { fusable = false; break; }
    if (b_same) {
      if (kp.nA == 2) { fusable = false; break; }
      kp.nA = 2; kp.a_k[1] = d.seg_a_k[s]; kp.a_mn[1] = d.seg_a_mn[s];
    } else {
      if (kp.nB == 2) { fusable = false; break; }
      kp.nB = 2; kp.b_k[1] = d.seg_b_k[s]; kp.b_mn[1] = d.seg_b_mn[s];
    }
  }
  if (fusable) {
    kp.total_kb = (d.seg_klen[0] + BK - 1) / BK;
    return gemm_fused(d, kp, stream);
  }
  // ---- general segment lists: one accumulate-launch per segment
  SGG_CHECK(t_plan_out == nullptr, "sgg_gemm_plan: this segment list runs as several launches, there is no single plan");
  SGG_CHECK(d.C && !d.Chl && d.out_d0 == 0 && !d.argmax_keys, "sgg_gemm: this segment pattern needs the plain fp32 output");
  if (!d.atomic) SGG_TRY(zero_2d(d.C, d.ldc, d.N, d.M, stream));
  for (int s = 0; s < d.nseg; ++s) {
    sgg_gemm_desc_t e = d;
    e.atomic = 1;
    if (s > 0) { e.bias = nullptr; e.addm = nullptr; }
    GemmKParams ks = kp;
    ks.atomic = 1;
    if (s > 0) { ks.bias = nullptr; ks.addm = nullptr; }
    ks.nA = ks.nB = 1;
    ks.a_k[0] = d.seg_a_k[s]; ks.a_mn[0] = d.seg_a_mn[s]; ks.b_k[0] = d.seg_b_k[s]; ks.b_mn[0] = d.seg_b_mn[s];
    ks.total_kb = (d.seg_klen[s] + BK - 1) / BK;
    SGG_TRY(gemm_fused(e, ks, stream));
  }
  return 0;
}

}  // namespace sgg

extern "C" int sgg_gemm_plan(const sgg_gemm_desc_t* d, int32_t* block_n, int32_t* splits, int32_t* k_blocks_per_split) {
  if (!d || !block_n || !splits || !k_blocks_per_split) {
    sgg::set_error("sgg_gemm_plan: null argument");
    return -1;
  }
  sgg::GemmPlan pl{0, 0, 0};
  sgg::t_plan_out = &pl;
  const int rc = sgg::gemm(*d, nullptr);
  sgg::t_plan_out = nullptr;
  if (rc != 0) return rc;
  *block_n = pl.block_n;
  *splits = pl.splits;
  const int total_kb = (d->seg_klen[0] + 63) / 64;
  *k_blocks_per_split = (total_kb + pl.splits - 1) / pl.splits;
  return 0;
}

extern "C" int sgg_gemm(const sgg_gemm_desc_t* d, sgg_stream_t stream) {
  if (!d) {
    sgg::set_error("sgg_gemm: null descriptor");
    return -1;
  }
  return sgg::gemm(*d, reinterpret_cast<cudaStream_t>(stream));
}

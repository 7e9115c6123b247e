// Per-timestep attention kernels (HBM-bound): one CTA per sample streams that sample's
// annotation tile a[b] (R x 512 bf16, 200 KB) through a ring of shared-memory stages with
// TMA bulk copies, and serves every stream (fake / real / interpolate / tangent) that
// shares the tile from the one read.
//
//   attn_fwd : alpha = softmax(e) over R (gen:16) ; z = sum_r alpha_r a_r (gen:17)
//   attn_tan : JVP of the same (interp stream of the WGAN-GP tangent pass)
//   attn_rev : alpha_bar_r = <z_bar, a_r>, softmax reverse (first order, or reverse over
//              primal+tangent for the interp stream), P_bar accumulation.
// e = P + c W_h itself (gen:14-15 in split form) is produced by the tcgen05 GEMM.
#include <cstdlib>

#include "common.cuh"
#include "../../include/sgg_b200.h"

namespace sgg {

constexpr int AT_C = 512;           // channels (gen:68)
constexpr int AT_THREADS = 256;
constexpr int AT_CHUNK_ROWS = 28;   // 196 = 7 * 28
constexpr int AT_STAGES = 3;
constexpr int AT_RMAX = 256;
constexpr int AT_MAXV = 8;          // streams served by one forward tile read (generator: n_critic+1 noise draws)
constexpr int AT_MAXV_REV = 4;      // streams served by one reverse tile read
constexpr int AT_STAGE_BYTES = AT_CHUNK_ROWS * AT_C * 2;  // 28 KB

struct AttnSmem {
  __align__(128) uint8_t stage[AT_STAGES][AT_STAGE_BYTES];
  __align__(16) float w[AT_RMAX][AT_MAXV];       // fwd: weights per region and stream; rev: alpha_bar (first 4)
  __align__(16) float eb[AT_MAXV_REV][AT_RMAX];  // rev: e_bar per stream
  __align__(8) uint64_t full[AT_STAGES];
};

__device__ __forceinline__ void attn_issue_chunk(AttnSmem& sm, const __nv_bfloat16* a_b, int chunk, int R) {
  const int r0 = chunk * AT_CHUNK_ROWS;
  const int rows = min(AT_CHUNK_ROWS, R - r0);
  const int st = chunk % AT_STAGES;
  const uint32_t bytes = rows * AT_C * 2;
  mbar_expect_tx(&sm.full[st], bytes);
  bulk_load_1d(sm.stage[st], a_b + (size_t)r0 * AT_C, bytes, &sm.full[st]);
}

// ------------------------------------------------------------------------------------ fwd
struct AttnFwdParams {
  const __nv_bfloat16* a;  // [B,R,C]
  int B, R, nv;
  int row_blk[AT_MAXV];    // stream v writes rows row_blk[v]*B + b of alpha_out / X
  int e_blk[AT_MAXV];      // ... and reads rows e_blk[v]*B + b of E
  int ain_blk;             // mode 1: saved alpha is read from rows ain_blk*B + b
  // mode 0 (softmax): E -> alpha ; mode 1 (tangent): alpha (saved) + edot -> adot
  const float* E; long long ldE;      // scores (mode 0) or edot (mode 1)
  const float* alpha_in;              // mode 1: saved alpha (same row indexing / ld as alpha_out)
  float* alpha_out; long long ldA;    // alpha (mode 0) or adot (mode 1)
  __nv_bfloat16* X; long long ldX; long long lo_off;  // z written at X[row, 0:C] (hi) and +lo_off (lo)
  int early_a;             // the annotations are not written by any kernel of this stream's PDL chain
  int l2_keep;             // load the tile with the L2 evict_last policy (it is re-read by the next attention launches)
  // optional: zero-fill the gate GEMM's output rows of this sample's streams (zero_cols floats at zero_p + row * zero_ld),
  // so that the split-K GEMM behind this kernel needs no zero-fill launch of its own
  float* zero_p; long long zero_ld; int zero_cols;
};

template <int MODE, int NV>
__global__ void __launch_bounds__(AT_THREADS) attn_fwd_kernel(const AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  AttnSmem& sm = *reinterpret_cast<AttnSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  const int b = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R = p.R;
  const int nchunks = (R + AT_CHUNK_ROWS - 1) / AT_CHUNK_ROWS;
  const __nv_bfloat16* a_b = p.a + (size_t)b * R * AT_C;
  pdl_trigger();
  if (tid == 0) {
    for (int s = 0; s < AT_STAGES; ++s) mbar_init(&sm.full[s], 1);
    mbar_fence_init();
    // annotations that no kernel of the current stream writes (a training-iteration input) may be fetched before
    // the preceding kernel has finished; everything else waits for it
    if (!p.early_a) pdl_wait();
    for (int c = 0; c < min(AT_STAGES, nchunks); ++c) attn_issue_chunk(sm, a_b, c, R);
  }
  pdl_wait();
  // ---- per-stream weights (softmax or its tangent), one warp per stream (8 warps >= AT_MAXV streams)
  if (warp >= p.nv && warp < NV) {
    for (int r = lane; r < R; r += 32) sm.w[r][warp] = 0.f;   // unused stream slots of the float4 reads
  }
  if (warp < p.nv) {
    const long long row = (long long)p.row_blk[warp] * p.B + b;
    const float* e = p.E + ((long long)p.e_blk[warp] * p.B + b) * p.ldE;
    float v[AT_RMAX / 32];
    if (MODE == 0) {
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < AT_RMAX / 32; ++i) {
        const int r = lane + 32 * i;
        v[i] = (r < R) ? e[r] : -INFINITY;
        mx = fmaxf(mx, v[i]);
      }
      mx = warp_max(mx);
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < AT_RMAX / 32; ++i) {
        v[i] = (lane + 32 * i < R) ? __expf(v[i] - mx) : 0.f;
        s += v[i];
      }
      s = warp_sum(s);
      const float inv = 1.0f / s;
#pragma unroll
      for (int i = 0; i < AT_RMAX / 32; ++i) v[i] *= inv;
    } else {
      const float* al = p.alpha_in + ((long long)p.ain_blk * p.B + b) * p.ldA;
      float m = 0.f;
      float ed[AT_RMAX / 32];
#pragma unroll
      for (int i = 0; i < AT_RMAX / 32; ++i) {
        const int r = lane + 32 * i;
        v[i] = (r < R) ? al[r] : 0.f;
        ed[i] = (r < R) ? e[r] : 0.f;
        m += v[i] * ed[i];
      }
      m = warp_sum(m);
#pragma unroll
      for (int i = 0; i < AT_RMAX / 32; ++i) v[i] = v[i] * (ed[i] - m);
    }
    float* ao = p.alpha_out + row * p.ldA;
#pragma unroll
    for (int i = 0; i < AT_RMAX / 32; ++i) {
      const int r = lane + 32 * i;
      if (r < R) {
        ao[r] = v[i];
        sm.w[r][warp] = v[i];
      }
    }
  }
  __syncthreads();
  // ---- z_v = sum_r w_v[r] a[r,:] ; thread owns channels 2*tid, 2*tid+1
  float acc[NV][2];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v][0] = acc[v][1] = 0.f;
  for (int c = 0; c < nchunks; ++c) {
    const int st = c % AT_STAGES;
    mbar_wait(&sm.full[st], (c / AT_STAGES) & 1);
    const int r0 = c * AT_CHUNK_ROWS;
    const int rows = min(AT_CHUNK_ROWS, R - r0);
    const uint32_t* tile = reinterpret_cast<const uint32_t*>(sm.stage[st]);
#pragma unroll 4
    for (int r = 0; r < rows; ++r) {
      const uint32_t pk = tile[r * (AT_C / 2) + tid];
      const float a0 = bf16lo_to_f32(pk), a1 = bf16hi_to_f32(pk);
#pragma unroll
      for (int v4 = 0; v4 < NV / 4; ++v4) {
        const float4 w = *reinterpret_cast<const float4*>(&sm.w[r0 + r][4 * v4]);
        acc[4 * v4 + 0][0] = fmaf(w.x, a0, acc[4 * v4 + 0][0]); acc[4 * v4 + 0][1] = fmaf(w.x, a1, acc[4 * v4 + 0][1]);
        acc[4 * v4 + 1][0] = fmaf(w.y, a0, acc[4 * v4 + 1][0]); acc[4 * v4 + 1][1] = fmaf(w.y, a1, acc[4 * v4 + 1][1]);
        acc[4 * v4 + 2][0] = fmaf(w.z, a0, acc[4 * v4 + 2][0]); acc[4 * v4 + 2][1] = fmaf(w.z, a1, acc[4 * v4 + 2][1]);
        acc[4 * v4 + 3][0] = fmaf(w.w, a0, acc[4 * v4 + 3][0]); acc[4 * v4 + 3][1] = fmaf(w.w, a1, acc[4 * v4 + 3][1]);
      }
    }
    __syncthreads();
    if (tid == 0 && c + AT_STAGES < nchunks) attn_issue_chunk(sm, a_b, c + AT_STAGES, R);
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    if (v < p.nv) {
      const long long row = (long long)p.row_blk[v] * p.B + b;
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(acc[v][0], h0, l0);
      split_bf16(acc[v][1], h1, l1);
      uint32_t* xh = reinterpret_cast<uint32_t*>(p.X + row * p.ldX);
      uint32_t* xl = reinterpret_cast<uint32_t*>(p.X + row * p.ldX + p.lo_off);
      xh[tid] = pack_bf16x2(h0, h1);
      xl[tid] = pack_bf16x2(l0, l1);
    }
  }
}

// ------------------------------------------------------------------------------------ fwd on tensor cores
// The context reduction z_v = sum_r w_v[r] a[r,:] is a small contraction per sample: D[channel, stream] =
// sum_r a[r, channel] * w[stream, r] with M = 512 channels (4 UMMA tiles of 128), N = streams padded to 16,
// K = R regions.  Done with CUDA-core FMAs it is issue-bound (2*NV FMAs per tile element); on tcgen05 the tile
// streams HBM -> shared memory (TMA, 128B-swizzled MN-major A operand) -> tensor core and no thread ever
// touches it.  The softmax weights are the B operand, written to shared memory as a bf16 hi/lo pair (two MMAs per
// k-step) so that alpha keeps 2^-17 relative precision; accumulation is fp32 in TMEM.
//   warp 0: TMA producer (ring of A2_STAGES stages of 16 regions x 512 channels = 16 KB)
//   warp 1: TMEM allocation + MMA issue
//   warps 2-5: softmax (or its tangent) -> weights, then epilogue TMEM -> z hi/lo
constexpr int A2_ROWS = 16;
constexpr int A2_STAGES = 5;
constexpr int A2_STAGE_BYTES = A2_ROWS * AT_C * 2;   // 16 KB = 8 boxes of {64 channels x 16 regions}
constexpr int A2_THREADS = 192;
constexpr int A2_NPAD = 16;
constexpr int A2_W_BYTES = A2_NPAD * AT_RMAX * 2;    // 8 KB per part: 4 k-blocks of {16 rows x 128 B}
constexpr int A2_TMEM_COLS = 64;                     // 4 channel tiles x 16 streams

struct Attn2Smem {
  uint8_t stage[A2_STAGES][A2_STAGE_BYTES];
  uint8_t whi[A2_W_BYTES];
  uint8_t wlo[A2_W_BYTES];
  uint64_t full[A2_STAGES];
  uint64_t empty[A2_STAGES];
  uint64_t wready, tmem_full;
  uint32_t tmem_ptr;
};

// byte offset of weight element (stream n, region k) inside a K-major, 128B-swizzled [16 x 256] bf16 operand
__device__ __forceinline__ uint32_t a2_w_off(int n, int k) {
  return (uint32_t)((k >> 6) * 2048 + (n >> 3) * 1024 + (n & 7) * 128 + ((((k & 63) >> 3) ^ (n & 7)) << 4) + (k & 7) * 2);
}

template <int MODE>
__global__ void __launch_bounds__(A2_THREADS) attn_fwd_mma_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                   const AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  Attn2Smem& sm = *reinterpret_cast<Attn2Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int R = p.R;
  const int nch = (R + A2_ROWS - 1) / A2_ROWS;
  pdl_trigger();
  if (warp == 0 && elect_one()) tma_prefetch_desc(&tmA);
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < A2_STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
      mbar_init(&sm.wready, 128);
      mbar_init(&sm.tmem_full, 1);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(&sm.tmem_ptr, A2_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      if (!p.early_a) pdl_wait();   // annotations that a preceding kernel may have written
      const uint64_t pol = l2_policy_evict_last();
      for (int c = 0; c < nch; ++c) {
        const int st = c % A2_STAGES;
        mbar_wait(&sm.empty[st], ((c / A2_STAGES) & 1) ^ 1);
        mbar_expect_tx(&sm.full[st], A2_STAGE_BYTES);
#pragma unroll
        for (int j = 0; j < AT_C / 64; ++j)
          if (p.l2_keep) tma_load_3d_hint(sm.stage[st] + j * (A2_ROWS * 128), &tmA, &sm.full[st], 64 * j, c * A2_ROWS, b, pol);
          else tma_load_3d(sm.stage[st] + j * (A2_ROWS * 128), &tmA, &sm.full[st], 64 * j, c * A2_ROWS, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc_bf16(128, A2_NPAD, true, false);
    mbar_wait(&sm.wready, 0);
    tc_fence_after();
    const uint32_t whi = smem_u32(sm.whi), wlo = smem_u32(sm.wlo);
    for (int c = 0; c < nch; ++c) {
      const int st = c % A2_STAGES;
      mbar_wait(&sm.full[st], (c / A2_STAGES) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sA = smem_u32(sm.stage[st]);
        const uint32_t boff = (uint32_t)((c >> 2) * 2048 + (c & 3) * 32);   // k-step c = regions [16c, 16c+16)
        const uint64_t dbh = make_smem_desc(whi + boff, 0, 1024);
        const uint64_t dbl = make_smem_desc(wlo + boff, 0, 1024);
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
          const uint64_t da = make_smem_desc(sA + mt * (2 * A2_ROWS * 128), A2_ROWS * 128, 1024);
          umma_bf16(tmem_base + mt * A2_NPAD, da, dbh, idesc, c ? 1u : 0u);
          umma_bf16(tmem_base + mt * A2_NPAD, da, dbl, idesc, 1u);
        }
        umma_commit(&sm.empty[st]);
        if (c == nch - 1) umma_commit(&sm.tmem_full);
      }
      __syncwarp();
    }
  } else {
    // ===================== softmax / tangent weights, then epilogue =====================
    pdl_wait();
    const int wi = warp - 2;
#pragma unroll 1
    for (int n = wi; n < p.nv; n += 4) {   // rows of unused streams stay unwritten: their D columns are never read
      float v[AT_RMAX / 32];
      {
        const long long row = (long long)p.row_blk[n] * p.B + b;
        const float* e = p.E + ((long long)p.e_blk[n] * p.B + b) * p.ldE;
        if (MODE == 0) {
          float mx = -INFINITY;
#pragma unroll
          for (int i = 0; i < AT_RMAX / 32; ++i) {
            const int r = lane + 32 * i;
            v[i] = (r < R) ? e[r] : -INFINITY;
            mx = fmaxf(mx, v[i]);
          }
          mx = warp_max(mx);
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < AT_RMAX / 32; ++i) {
            v[i] = (lane + 32 * i < R) ? __expf(v[i] - mx) : 0.f;
            s += v[i];
          }
          s = warp_sum(s);
          const float inv = 1.0f / s;
#pragma unroll
          for (int i = 0; i < AT_RMAX / 32; ++i) v[i] *= inv;
        } else {
          const float* al = p.alpha_in + ((long long)p.ain_blk * p.B + b) * p.ldA;
          float m = 0.f;
          float ed[AT_RMAX / 32];
#pragma unroll
          for (int i = 0; i < AT_RMAX / 32; ++i) {
            const int r = lane + 32 * i;
            v[i] = (r < R) ? al[r] : 0.f;
            ed[i] = (r < R) ? e[r] : 0.f;
            m += v[i] * ed[i];
          }
          m = warp_sum(m);
#pragma unroll
          for (int i = 0; i < AT_RMAX / 32; ++i) v[i] = v[i] * (ed[i] - m);
        }
        float* ao = p.alpha_out + row * p.ldA;
#pragma unroll
        for (int i = 0; i < AT_RMAX / 32; ++i)
          if (lane + 32 * i < R) ao[lane + 32 * i] = v[i];
      }
#pragma unroll
      for (int i = 0; i < AT_RMAX / 32; ++i) {
        __nv_bfloat16 h, l;
        split_bf16(v[i], h, l);
        const uint32_t off = a2_w_off(n, lane + 32 * i);
        *reinterpret_cast<__nv_bfloat16*>(sm.whi + off) = h;
        *reinterpret_cast<__nv_bfloat16*>(sm.wlo + off) = l;
      }
    }
    fence_proxy_async_smem();
    mbar_arrive(&sm.wready);
    if (p.zero_p) {   // while the contraction runs: clear the gate pre-activation rows of this sample's streams
      const int wt = threadIdx.x - 64;
      for (int v = 0; v < p.nv; ++v) {
        float4* d = reinterpret_cast<float4*>(p.zero_p + ((long long)p.row_blk[v] * p.B + b) * p.zero_ld);
        for (int i = wt; i < (p.zero_cols >> 2); i += 128) d[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    // ---- epilogue: TMEM lane = channel within the tile, column = stream
    const int q = warp & 3;
    mbar_wait(&sm.tmem_full, 0);
    tc_fence_after();
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      uint32_t r[16];
      tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * A2_NPAD), r);
      tmem_ld_wait();
      const int ch = mt * 128 + q * 32 + lane;
#pragma unroll
      for (int v = 0; v < AT_MAXV; ++v) {
        if (v < p.nv) {
          const long long row = (long long)p.row_blk[v] * p.B + b;
          __nv_bfloat16 h, l;
          split_bf16(__uint_as_float(r[v]), h, l);
          p.X[row * p.ldX + ch] = h;
          p.X[row * p.ldX + p.lo_off + ch] = l;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, A2_TMEM_COLS);
}

// ------------------------------------------------------------------------------------ rev
struct AttnRevParams {
  const __nv_bfloat16* a;
  int B, R;
  int nv;                   // number of z_bar vectors (primal streams [+ 1 tangent adjoint, last])
  int tan_stream;           // index v of the stream that carries the tangent, or -1
  int row_blk[AT_MAXV_REV]; // row block of each vector in XB / alpha / EB
  const float* XB; long long ldXB;        // z_bar = XB[row, 0:C]
  const float* alpha; long long ldA;      // saved alpha
  const float* edot;                      // [B, ldA] tangent of e for tan_stream (row b)
  __nv_bfloat16* EB; long long ldEB; long long lo_off;  // e_bar hi/lo out (pad columns untouched)
  float* Pbar; long long ldP;             // [B,R] += sum over primal streams of e_bar (may be null)
  int early_a;                            // see AttnFwdParams
  int l2_keep;
};

__global__ void __launch_bounds__(AT_THREADS) attn_rev_kernel(const AttnRevParams p) {
  extern __shared__ uint8_t smem_raw[];
  AttnSmem& sm = *reinterpret_cast<AttnSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  const int b = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R = p.R;
  const int nchunks = (R + AT_CHUNK_ROWS - 1) / AT_CHUNK_ROWS;
  const __nv_bfloat16* a_b = p.a + (size_t)b * R * AT_C;
  pdl_trigger();
  if (tid == 0) {
    for (int s = 0; s < AT_STAGES; ++s) mbar_init(&sm.full[s], 1);
    mbar_fence_init();
    if (!p.early_a) pdl_wait();
    for (int c = 0; c < min(AT_STAGES, nchunks); ++c) attn_issue_chunk(sm, a_b, c, R);
  }
  pdl_wait();
  // z_bar slices: lane owns channels [lane*8, +8) and [256 + lane*8, +8)
  float zb[AT_MAXV_REV][16];
#pragma unroll
  for (int v = 0; v < AT_MAXV_REV; ++v) {
    if (v < p.nv) {
      const float* z = p.XB + ((long long)p.row_blk[v] * p.B + b) * p.ldXB;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4 q0 = *reinterpret_cast<const float4*>(z + h * 256 + lane * 8);
        const float4 q1 = *reinterpret_cast<const float4*>(z + h * 256 + lane * 8 + 4);
        zb[v][h * 8 + 0] = q0.x; zb[v][h * 8 + 1] = q0.y; zb[v][h * 8 + 2] = q0.z; zb[v][h * 8 + 3] = q0.w;
        zb[v][h * 8 + 4] = q1.x; zb[v][h * 8 + 5] = q1.y; zb[v][h * 8 + 6] = q1.z; zb[v][h * 8 + 7] = q1.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) zb[v][j] = 0.f;
    }
  }
  __syncthreads();  // barrier init visible before anyone waits
  for (int c = 0; c < nchunks; ++c) {
    const int st = c % AT_STAGES;
    mbar_wait(&sm.full[st], (c / AT_STAGES) & 1);
    const int r0 = c * AT_CHUNK_ROWS;
    const int rows = min(AT_CHUNK_ROWS, R - r0);
    for (int r = warp; r < rows; r += AT_THREADS / 32) {
      const uint4* rowp = reinterpret_cast<const uint4*>(sm.stage[st] + (size_t)r * AT_C * 2);
      float d[AT_MAXV_REV] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint4 pk = rowp[h * 32 + lane];
        const uint32_t w4[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float a0 = bf16lo_to_f32(w4[e]), a1 = bf16hi_to_f32(w4[e]);
#pragma unroll
          for (int v = 0; v < AT_MAXV_REV; ++v) {
            d[v] = fmaf(a0, zb[v][h * 8 + 2 * e], d[v]);
            d[v] = fmaf(a1, zb[v][h * 8 + 2 * e + 1], d[v]);
          }
        }
      }
#pragma unroll
      for (int v = 0; v < AT_MAXV_REV; ++v) d[v] = warp_sum(d[v]);
      if (lane == 0) *reinterpret_cast<float4*>(sm.w[r0 + r]) = make_float4(d[0], d[1], d[2], d[3]);
    }
    __syncthreads();
    if (tid == 0 && c + AT_STAGES < nchunks) attn_issue_chunk(sm, a_b, c + AT_STAGES, R);
  }
  // ---- softmax reverse: one warp per primal stream
  const int ns = (p.tan_stream >= 0) ? p.nv - 1 : p.nv;
  if (warp < ns) {
    const int v = warp;
    const long long row = (long long)p.row_blk[v] * p.B + b;
    const float* al = p.alpha + row * p.ldA;
    const bool tan = (v == p.tan_stream);
    float alv[AT_RMAX / 32], ab[AT_RMAX / 32], adb[AT_RMAX / 32], ed[AT_RMAX / 32];
    float m_t = 0.f, m_e = 0.f;
#pragma unroll
    for (int i = 0; i < AT_RMAX / 32; ++i) {
      const int r = lane + 32 * i;
      const bool ok = r < R;
      alv[i] = ok ? al[r] : 0.f;
      ab[i] = ok ? sm.w[r][v] : 0.f;
      adb[i] = (ok && tan) ? sm.w[r][p.nv - 1] : 0.f;
      ed[i] = (ok && tan) ? p.edot[(long long)b * p.ldA + r] : 0.f;
      m_t += alv[i] * adb[i];
      m_e += alv[i] * ed[i];
    }
    float s = 0.f;
    if (tan) {
      m_t = warp_sum(m_t);
      m_e = warp_sum(m_e);
    }
#pragma unroll
    for (int i = 0; i < AT_RMAX / 32; ++i) {
      if (tan) ab[i] = ab[i] + adb[i] * (ed[i] - m_e) - ed[i] * m_t;  // w = abar + second-order terms
      s += alv[i] * ab[i];
    }
    s = warp_sum(s);
    __nv_bfloat16* ebh = p.EB + row * p.ldEB;
    __nv_bfloat16* tbh = tan ? p.EB + ((long long)p.row_blk[p.nv - 1] * p.B + b) * p.ldEB : nullptr;
#pragma unroll
    for (int i = 0; i < AT_RMAX / 32; ++i) {
      const int r = lane + 32 * i;
      if (r < R) {
        const float ebar = alv[i] * (ab[i] - s);
        sm.eb[v][r] = ebar;
        __nv_bfloat16 h, l;
        split_bf16(ebar, h, l);
        ebh[r] = h;
        ebh[p.lo_off + r] = l;
        if (tan) {
          const float edb = alv[i] * (adb[i] - m_t);
          split_bf16(edb, h, l);
          tbh[r] = h;
          tbh[p.lo_off + r] = l;
        }
      }
    }
  }
  if (p.Pbar) {
    __syncthreads();
    for (int r = tid; r < R; r += AT_THREADS) {
      float s = 0.f;
      for (int v = 0; v < ns; ++v) s += sm.eb[v][r];
      p.Pbar[(long long)b * p.ldP + r] += s;
    }
  }
}

// ------------------------------------------------------------------------------------ rev on tensor cores
// alpha_bar_v[r] = <z_bar_v, a[r,:]> is the contraction D[region, stream] = sum_c a[r, c] * zbar[stream, c]:
// M = regions (UMMA tiles of 128 rows; rows beyond R are zero-filled by TMA without memory traffic), N = streams
// padded to 16, K = 512 channels.  The annotation tile is the K-major A operand in its natural layout; z_bar is
// written to shared memory as a bf16 hi/lo pair (B operand).  The softmax reverse then runs on the 4 worker warps.
constexpr int R2_STAGES = 4;
constexpr int R2_STAGE_BYTES = 128 * 64 * 2;         // 16 KB: {64 channels x 128 regions}
constexpr int R2_ZB_BYTES = A2_NPAD * AT_C * 2;      // 16 KB per part: 8 k-blocks of {16 rows x 128 B}
constexpr int R2_TMEM_COLS = 32;                     // 2 region tiles x 16 streams

struct AttnRev2Smem {
  uint8_t stage[R2_STAGES][R2_STAGE_BYTES];
  uint8_t zhi[R2_ZB_BYTES];
  uint8_t zlo[R2_ZB_BYTES];
  __align__(16) float w[AT_RMAX][AT_MAXV_REV];       // alpha_bar per region and stream
  __align__(16) float eb[AT_MAXV_REV][AT_RMAX];      // e_bar per stream
  uint64_t full[R2_STAGES];
  uint64_t empty[R2_STAGES];
  uint64_t zready, tmem_full;
  uint32_t tmem_ptr;
};

__global__ void __launch_bounds__(A2_THREADS) attn_rev_mma_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                   const AttnRevParams p) {
  extern __shared__ uint8_t smem_raw[];
  AttnRev2Smem& sm = *reinterpret_cast<AttnRev2Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int R = p.R;
  const int nmt = (R + 127) / 128;            // region tiles
  const int nch = nmt * (AT_C / 64);          // ring chunks: (region tile, k-block of 64 channels)
  pdl_trigger();
  if (warp == 0 && elect_one()) tma_prefetch_desc(&tmA);
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < R2_STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
      mbar_init(&sm.zready, 128);
      mbar_init(&sm.tmem_full, 1);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(&sm.tmem_ptr, R2_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      if (!p.early_a) pdl_wait();
      const uint64_t pol = l2_policy_evict_last();
      for (int c = 0; c < nch; ++c) {
        const int st = c % R2_STAGES;
        mbar_wait(&sm.empty[st], ((c / R2_STAGES) & 1) ^ 1);
        mbar_expect_tx(&sm.full[st], R2_STAGE_BYTES);
        if (p.l2_keep) tma_load_3d_hint(sm.stage[st], &tmA, &sm.full[st], 64 * (c & 7), 128 * (c >> 3), b, pol);
        else tma_load_3d(sm.stage[st], &tmA, &sm.full[st], 64 * (c & 7), 128 * (c >> 3), b);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, A2_NPAD, false, false);
    mbar_wait(&sm.zready, 0);
    tc_fence_after();
    const uint32_t zhi = smem_u32(sm.zhi), zlo = smem_u32(sm.zlo);
    for (int c = 0; c < nch; ++c) {
      const int st = c % R2_STAGES;
      mbar_wait(&sm.full[st], (c / R2_STAGES) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sA = smem_u32(sm.stage[st]);
        const int kb = c & 7, mt = c >> 3;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = make_smem_desc(sA + k * 32, 0, 1024);
          const uint64_t dbh = make_smem_desc(zhi + kb * 2048 + k * 32, 0, 1024);
          const uint64_t dbl = make_smem_desc(zlo + kb * 2048 + k * 32, 0, 1024);
          umma_bf16(tmem_base + mt * A2_NPAD, da, dbh, idesc, (kb | k) ? 1u : 0u);
          umma_bf16(tmem_base + mt * A2_NPAD, da, dbl, idesc, 1u);
        }
        umma_commit(&sm.empty[st]);
        if (c == nch - 1) umma_commit(&sm.tmem_full);
      }
      __syncwarp();
    }
  } else {
    pdl_wait();
    const int wi = warp - 2;
    const int wtid = threadIdx.x - 64;         // 0..127 among the worker warps
    // ---- z_bar -> B operand (stream n, channel k), K-major 128B-swizzled, hi/lo.  Thread owns 4 channels of every
    // stream; rows of unused streams stay unwritten (their D columns are never read).
    {
      float4 zv[AT_MAXV_REV];
#pragma unroll
      for (int n = 0; n < AT_MAXV_REV; ++n)
        zv[n] = (n < p.nv) ? *reinterpret_cast<const float4*>(p.XB + ((long long)p.row_blk[n] * p.B + b) * p.ldXB + wtid * 4)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int n = 0; n < AT_MAXV_REV; ++n) {
        __nv_bfloat16 h[4], l[4];
        split_bf16(zv[n].x, h[0], l[0]); split_bf16(zv[n].y, h[1], l[1]);
        split_bf16(zv[n].z, h[2], l[2]); split_bf16(zv[n].w, h[3], l[3]);
        const uint32_t off = a2_w_off(n, wtid * 4);
        *reinterpret_cast<uint2*>(sm.zhi + off) = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
        *reinterpret_cast<uint2*>(sm.zlo + off) = make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
      }
    }
    fence_proxy_async_smem();
    mbar_arrive(&sm.zready);
    // ---- loads of the softmax reverse that do not depend on the contraction go out before waiting for it
    const int ns = (p.tan_stream >= 0) ? p.nv - 1 : p.nv;
    const bool tan = (wi == p.tan_stream);
    float alv[AT_RMAX / 32], ed[AT_RMAX / 32];
    if (wi < ns) {
      const float* al = p.alpha + ((long long)p.row_blk[wi] * p.B + b) * p.ldA;
#pragma unroll
      for (int i = 0; i < AT_RMAX / 32; ++i) {
        const int r = lane + 32 * i;
        alv[i] = (r < R) ? al[r] : 0.f;
        ed[i] = (r < R && tan) ? p.edot[(long long)b * p.ldA + r] : 0.f;
      }
    }
    // ---- alpha_bar from TMEM: lane = region within the tile, column = stream
    const int q = warp & 3;
    mbar_wait(&sm.tmem_full, 0);
    tc_fence_after();
    for (int mt = 0; mt < nmt; ++mt) {
      uint32_t r[16];
      tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * A2_NPAD), r);
      tmem_ld_wait();
      const int row = mt * 128 + q * 32 + lane;
      if (row < R)
        *reinterpret_cast<float4*>(sm.w[row]) = make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]),
                                                            __uint_as_float(r[2]), __uint_as_float(r[3]));
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    // ---- softmax reverse: one warp per primal stream
    if (wi < ns) {
      const int v = wi;
      const long long row = (long long)p.row_blk[v] * p.B + b;
      float ab[AT_RMAX / 32], adb[AT_RMAX / 32];
      float m_t = 0.f, m_e = 0.f;
#pragma unroll
      for (int i = 0; i < AT_RMAX / 32; ++i) {
        const int r = lane + 32 * i;
        const bool ok = r < R;
        ab[i] = ok ? sm.w[r][v] : 0.f;
        adb[i] = (ok && tan) ? sm.w[r][p.nv - 1] : 0.f;
        m_t += alv[i] * adb[i];
        m_e += alv[i] * ed[i];
      }
      float s = 0.f;
      if (tan) {
        m_t = warp_sum(m_t);
        m_e = warp_sum(m_e);
      }
#pragma unroll
      for (int i = 0; i < AT_RMAX / 32; ++i) {
        if (tan) ab[i] = ab[i] + adb[i] * (ed[i] - m_e) - ed[i] * m_t;  // w = abar + second-order terms
        s += alv[i] * ab[i];
      }
      s = warp_sum(s);
      __nv_bfloat16* ebh = p.EB + row * p.ldEB;
      __nv_bfloat16* tbh = tan ? p.EB + ((long long)p.row_blk[p.nv - 1] * p.B + b) * p.ldEB : nullptr;
#pragma unroll
      for (int i = 0; i < AT_RMAX / 32; ++i) {
        const int r = lane + 32 * i;
        if (r < R) {
          const float ebar = alv[i] * (ab[i] - s);
          sm.eb[v][r] = ebar;
          __nv_bfloat16 h, l;
          split_bf16(ebar, h, l);
          ebh[r] = h;
          ebh[p.lo_off + r] = l;
          if (tan) {
            const float edb = alv[i] * (adb[i] - m_t);
            split_bf16(edb, h, l);
            tbh[r] = h;
            tbh[p.lo_off + r] = l;
          }
        }
      }
    }
    if (p.Pbar) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int r = wtid; r < R; r += 128) {
        float s = 0.f;
        for (int v = 0; v < ns; ++v) s += sm.eb[v][r];
        atomicAdd(p.Pbar + (long long)b * p.ldP + r, s);   // fire-and-forget reduction (one writer per address and launch)
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, R2_TMEM_COLS);
}

// ------------------------------------------------------------------------------------ persistent variants
// One CTA per SM (grid = min(B, SM count)), each walking the samples b = blockIdx.x, blockIdx.x + grid, ...  The TMA ring
// runs ACROSS samples: while the workers compute the softmax of the next sample, read the previous accumulator out of
// TMEM and store z, the producer keeps up to AP_STAGES x 16 KB of annotation tiles in flight, so the HBM stream never
// drains at a sample boundary (the one-CTA-per-sample kernels above pay barrier init, TMEM allocation, the softmax
// before the first MMA and the epilogue once per 200 KB tile: 0.68 / 0.59 of HBM).  Weights and accumulators are
// double-buffered: the softmax of sample i+1 overlaps the contraction of sample i.
constexpr int AP_STAGES = 10;

template <int MODE>
__device__ __forceinline__ void attn_stream_weights(const AttnFwdParams& p, int b, int n, int lane, uint8_t* whi, uint8_t* wlo) {
  const int R = p.R;
  float v[AT_RMAX / 32];
  const long long row = (long long)p.row_blk[n] * p.B + b;
  const float* e = p.E + ((long long)p.e_blk[n] * p.B + b) * p.ldE;
  if (MODE == 0) {
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < AT_RMAX / 32; ++i) {
      const int r = lane + 32 * i;
      v[i] = (r < R) ? e[r] : -INFINITY;
      mx = fmaxf(mx, v[i]);
    }
    mx = warp_max(mx);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < AT_RMAX / 32; ++i) {
      v[i] = (lane + 32 * i < R) ? __expf(v[i] - mx) : 0.f;
      s += v[i];
    }
    s = warp_sum(s);
    const float inv = 1.0f / s;
#pragma unroll
    for (int i = 0; i < AT_RMAX / 32; ++i) v[i] *= inv;
  } else {
    const float* al = p.alpha_in + ((long long)p.ain_blk * p.B + b) * p.ldA;
    float m = 0.f;
    float ed[AT_RMAX / 32];
#pragma unroll
    for (int i = 0; i < AT_RMAX / 32; ++i) {
      const int r = lane + 32 * i;
      v[i] = (r < R) ? al[r] : 0.f;
      ed[i] = (r < R) ? e[r] : 0.f;
      m += v[i] * ed[i];
    }
    m = warp_sum(m);
#pragma unroll
    for (int i = 0; i < AT_RMAX / 32; ++i) v[i] = v[i] * (ed[i] - m);
  }
  float* ao = p.alpha_out + row * p.ldA;
#pragma unroll
  for (int i = 0; i < AT_RMAX / 32; ++i)
    if (lane + 32 * i < R) ao[lane + 32 * i] = v[i];
#pragma unroll
  for (int i = 0; i < AT_RMAX / 32; ++i) {
    __nv_bfloat16 h, l;
    split_bf16(v[i], h, l);
    const uint32_t off = a2_w_off(n, lane + 32 * i);
    *reinterpret_cast<__nv_bfloat16*>(whi + off) = h;
    *reinterpret_cast<__nv_bfloat16*>(wlo + off) = l;
  }
}

struct AttnPSmem {
  uint8_t stage[AP_STAGES][A2_STAGE_BYTES];
  uint8_t whi[2][A2_W_BYTES];
  uint8_t wlo[2][A2_W_BYTES];
  uint64_t full[AP_STAGES];
  uint64_t empty[AP_STAGES];
  uint64_t wready[2], tmem_full[2], tmem_empty[2];
  uint32_t tmem_ptr;
};

template <int MODE>
__global__ void __launch_bounds__(A2_THREADS, 1) attn_fwd_p_kernel(const __grid_constant__ CUtensorMap tmA, const AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  AttnPSmem& sm = *reinterpret_cast<AttnPSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int R = p.R;
  const int nch = (R + A2_ROWS - 1) / A2_ROWS;
  const int G = gridDim.x;
  const int n_items = ((int)blockIdx.x < p.B) ? (p.B - (int)blockIdx.x + G - 1) / G : 0;
  pdl_trigger();
  if (warp == 0 && elect_one()) tma_prefetch_desc(&tmA);
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < AP_STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
      for (int s = 0; s < 2; ++s) { mbar_init(&sm.wready[s], 128); mbar_init(&sm.tmem_full[s], 1); mbar_init(&sm.tmem_empty[s], 128); }
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(&sm.tmem_ptr, 2 * A2_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer: the ring runs across samples =====================
    if (elect_one()) {
      if (!p.early_a) pdl_wait();   // annotations that a preceding kernel may have written
      const uint64_t pol = l2_policy_evict_last();
      int cg = 0;
      for (int it = 0; it < n_items; ++it) {
        const int b = (int)blockIdx.x + it * G;
        for (int c = 0; c < nch; ++c, ++cg) {
          const int st = cg % AP_STAGES;
          mbar_wait(&sm.empty[st], ((cg / AP_STAGES) & 1) ^ 1);
          mbar_expect_tx(&sm.full[st], A2_STAGE_BYTES);
#pragma unroll
          for (int j = 0; j < AT_C / 64; ++j)
            if (p.l2_keep) tma_load_3d_hint(sm.stage[st] + j * (A2_ROWS * 128), &tmA, &sm.full[st], 64 * j, c * A2_ROWS, b, pol);
            else tma_load_3d(sm.stage[st] + j * (A2_ROWS * 128), &tmA, &sm.full[st], 64 * j, c * A2_ROWS, b);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc_bf16(128, A2_NPAD, true, false);
    int cg = 0;
    for (int it = 0; it < n_items; ++it) {
      const int slot = it & 1;
      mbar_wait(&sm.wready[slot], (it >> 1) & 1);              // weights of this sample are in shared memory
      mbar_wait(&sm.tmem_empty[slot], ((it >> 1) & 1) ^ 1);    // accumulator slot read out (sample it - 2)
      tc_fence_after();
      const uint32_t whi = smem_u32(sm.whi[slot]), wlo = smem_u32(sm.wlo[slot]);
      const uint32_t acc = tmem_base + slot * A2_TMEM_COLS;
      for (int c = 0; c < nch; ++c, ++cg) {
        const int st = cg % AP_STAGES;
        mbar_wait(&sm.full[st], (cg / AP_STAGES) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sA = smem_u32(sm.stage[st]);
          const uint32_t boff = (uint32_t)((c >> 2) * 2048 + (c & 3) * 32);   // k-step c = regions [16c, 16c+16)
          const uint64_t dbh = make_smem_desc(whi + boff, 0, 1024);
          const uint64_t dbl = make_smem_desc(wlo + boff, 0, 1024);
#pragma unroll
          for (int mt = 0; mt < 4; ++mt) {
            const uint64_t da = make_smem_desc(sA + mt * (2 * A2_ROWS * 128), A2_ROWS * 128, 1024);
            umma_bf16(acc + mt * A2_NPAD, da, dbh, idesc, c ? 1u : 0u);
            umma_bf16(acc + mt * A2_NPAD, da, dbl, idesc, 1u);
          }
          umma_commit(&sm.empty[st]);
          if (c == nch - 1) umma_commit(&sm.tmem_full[slot]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== workers: weights of sample it + 1, then read-out of sample it =====================
    pdl_wait();
    const int wi = warp - 2;
    const int q = warp & 3;
    auto weights = [&](int it) {
      const int b = (int)blockIdx.x + it * G, slot = it & 1;
#pragma unroll 1
      for (int n = wi; n < p.nv; n += 4) attn_stream_weights<MODE>(p, b, n, lane, sm.whi[slot], sm.wlo[slot]);
      fence_proxy_async_smem();
      mbar_arrive(&sm.wready[slot]);
    };
    if (n_items > 0) weights(0);
    for (int it = 0; it < n_items; ++it) {
      const int b = (int)blockIdx.x + it * G, slot = it & 1;
      if (it + 1 < n_items) weights(it + 1);   // its weight slot was last read by the MMAs of sample it - 1 (tmem_full seen)
      if (p.zero_p) {   // clear the gate pre-activation rows of this sample's streams
        const int wt = threadIdx.x - 64;
        for (int v = 0; v < p.nv; ++v) {
          float4* d = reinterpret_cast<float4*>(p.zero_p + ((long long)p.row_blk[v] * p.B + b) * p.zero_ld);
          for (int i = wt; i < (p.zero_cols >> 2); i += 128) d[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      mbar_wait(&sm.tmem_full[slot], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        uint32_t r[16];
        tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * A2_TMEM_COLS + mt * A2_NPAD), r);
        tmem_ld_wait();
        const int ch = mt * 128 + q * 32 + lane;
#pragma unroll
        for (int v = 0; v < AT_MAXV; ++v) {
          if (v < p.nv) {
            const long long row = (long long)p.row_blk[v] * p.B + b;
            __nv_bfloat16 h, l;
            split_bf16(__uint_as_float(r[v]), h, l);
            p.X[row * p.ldX + ch] = h;
            p.X[row * p.ldX + p.lo_off + ch] = l;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&sm.tmem_empty[slot]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * A2_TMEM_COLS);
}

constexpr int RP_STAGES = 8;
struct AttnRevPSmem {
  uint8_t stage[RP_STAGES][R2_STAGE_BYTES];
  uint8_t zhi[2][R2_ZB_BYTES];
  uint8_t zlo[2][R2_ZB_BYTES];
  __align__(16) float w[AT_RMAX][AT_MAXV_REV];       // alpha_bar per region and stream
  __align__(16) float eb[AT_MAXV_REV][AT_RMAX];      // e_bar per stream
  uint64_t full[RP_STAGES];
  uint64_t empty[RP_STAGES];
  uint64_t zready[2], tmem_full[2], tmem_empty[2];
  uint32_t tmem_ptr;
};

__global__ void __launch_bounds__(A2_THREADS, 1) attn_rev_p_kernel(const __grid_constant__ CUtensorMap tmA, const AttnRevParams p) {
  extern __shared__ uint8_t smem_raw[];
  AttnRevPSmem& sm = *reinterpret_cast<AttnRevPSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int R = p.R;
  const int nmt = (R + 127) / 128;            // region tiles
  const int nch = nmt * (AT_C / 64);          // ring chunks per sample: (region tile, k-block of 64 channels)
  const int G = gridDim.x;
  const int n_items = ((int)blockIdx.x < p.B) ? (p.B - (int)blockIdx.x + G - 1) / G : 0;
  pdl_trigger();
  if (warp == 0 && elect_one()) tma_prefetch_desc(&tmA);
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < RP_STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
      for (int s = 0; s < 2; ++s) { mbar_init(&sm.zready[s], 128); mbar_init(&sm.tmem_full[s], 1); mbar_init(&sm.tmem_empty[s], 128); }
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(&sm.tmem_ptr, 2 * R2_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      if (!p.early_a) pdl_wait();
      const uint64_t pol = l2_policy_evict_last();
      int cg = 0;
      for (int it = 0; it < n_items; ++it) {
        const int b = (int)blockIdx.x + it * G;
        for (int c = 0; c < nch; ++c, ++cg) {
          const int st = cg % RP_STAGES;
          mbar_wait(&sm.empty[st], ((cg / RP_STAGES) & 1) ^ 1);
          mbar_expect_tx(&sm.full[st], R2_STAGE_BYTES);
          if (p.l2_keep) tma_load_3d_hint(sm.stage[st], &tmA, &sm.full[st], 64 * (c & 7), 128 * (c >> 3), b, pol);
          else tma_load_3d(sm.stage[st], &tmA, &sm.full[st], 64 * (c & 7), 128 * (c >> 3), b);
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, A2_NPAD, false, false);
    int cg = 0;
    for (int it = 0; it < n_items; ++it) {
      const int slot = it & 1;
      mbar_wait(&sm.zready[slot], (it >> 1) & 1);
      mbar_wait(&sm.tmem_empty[slot], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t zhi = smem_u32(sm.zhi[slot]), zlo = smem_u32(sm.zlo[slot]);
      const uint32_t acc = tmem_base + slot * R2_TMEM_COLS;
      for (int c = 0; c < nch; ++c, ++cg) {
        const int st = cg % RP_STAGES;
        mbar_wait(&sm.full[st], (cg / RP_STAGES) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sA = smem_u32(sm.stage[st]);
          const int kb = c & 7, mt = c >> 3;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = make_smem_desc(sA + k * 32, 0, 1024);
            const uint64_t dbh = make_smem_desc(zhi + kb * 2048 + k * 32, 0, 1024);
            const uint64_t dbl = make_smem_desc(zlo + kb * 2048 + k * 32, 0, 1024);
            umma_bf16(acc + mt * A2_NPAD, da, dbh, idesc, (kb | k) ? 1u : 0u);
            umma_bf16(acc + mt * A2_NPAD, da, dbl, idesc, 1u);
          }
          umma_commit(&sm.empty[st]);
          if (c == nch - 1) umma_commit(&sm.tmem_full[slot]);
        }
        __syncwarp();
      }
    }
  } else {
    pdl_wait();
    const int wi = warp - 2;
    const int wtid = threadIdx.x - 64;         // 0..127 among the worker warps
    const int q = warp & 3;
    const int ns = (p.tan_stream >= 0) ? p.nv - 1 : p.nv;
    const bool tan = (wi == p.tan_stream);
    // z_bar of sample `it` -> B operand (stream n, channel k), K-major 128B-swizzled, hi/lo.  Thread owns 4 channels of
    // every stream; rows of unused streams stay unwritten (their D columns are never read).
    auto zprep = [&](int it) {
      const int b = (int)blockIdx.x + it * G, slot = it & 1;
      float4 zv[AT_MAXV_REV];
#pragma unroll
      for (int n = 0; n < AT_MAXV_REV; ++n)
        zv[n] = (n < p.nv) ? *reinterpret_cast<const float4*>(p.XB + ((long long)p.row_blk[n] * p.B + b) * p.ldXB + wtid * 4)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int n = 0; n < AT_MAXV_REV; ++n) {
        __nv_bfloat16 h[4], l[4];
        split_bf16(zv[n].x, h[0], l[0]); split_bf16(zv[n].y, h[1], l[1]);
        split_bf16(zv[n].z, h[2], l[2]); split_bf16(zv[n].w, h[3], l[3]);
        const uint32_t off = a2_w_off(n, wtid * 4);
        *reinterpret_cast<uint2*>(sm.zhi[slot] + off) = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
        *reinterpret_cast<uint2*>(sm.zlo[slot] + off) = make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
      }
      fence_proxy_async_smem();
      mbar_arrive(&sm.zready[slot]);
    };
    if (n_items > 0) zprep(0);
    for (int it = 0; it < n_items; ++it) {
      const int b = (int)blockIdx.x + it * G, slot = it & 1;
      if (it + 1 < n_items) zprep(it + 1);
      // ---- loads of the softmax reverse that do not depend on the contraction go out before waiting for it
      float alv[AT_RMAX / 32], ed[AT_RMAX / 32];
      if (wi < ns) {
        const float* al = p.alpha + ((long long)p.row_blk[wi] * p.B + b) * p.ldA;
#pragma unroll
        for (int i = 0; i < AT_RMAX / 32; ++i) {
          const int r = lane + 32 * i;
          alv[i] = (r < R) ? al[r] : 0.f;
          ed[i] = (r < R && tan) ? p.edot[(long long)b * p.ldA + r] : 0.f;
        }
      }
      // ---- alpha_bar from TMEM: lane = region within the tile, column = stream
      mbar_wait(&sm.tmem_full[slot], (it >> 1) & 1);
      tc_fence_after();
      for (int mt = 0; mt < nmt; ++mt) {
        uint32_t r[16];
        tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * R2_TMEM_COLS + mt * A2_NPAD), r);
        tmem_ld_wait();
        const int row = mt * 128 + q * 32 + lane;
        if (row < R)
          *reinterpret_cast<float4*>(sm.w[row]) = make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]),
                                                              __uint_as_float(r[2]), __uint_as_float(r[3]));
      }
      tc_fence_before();
      mbar_arrive(&sm.tmem_empty[slot]);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // ---- softmax reverse: one warp per primal stream
      if (wi < ns) {
        const int v = wi;
        const long long row = (long long)p.row_blk[v] * p.B + b;
        float ab[AT_RMAX / 32], adb[AT_RMAX / 32];
        float m_t = 0.f, m_e = 0.f;
#pragma unroll
        for (int i = 0; i < AT_RMAX / 32; ++i) {
          const int r = lane + 32 * i;
          const bool ok = r < R;
          ab[i] = ok ? sm.w[r][v] : 0.f;
          adb[i] = (ok && tan) ? sm.w[r][p.nv - 1] : 0.f;
          m_t += alv[i] * adb[i];
          m_e += alv[i] * ed[i];
        }
        float s = 0.f;
        if (tan) {
          m_t = warp_sum(m_t);
          m_e = warp_sum(m_e);
        }
#pragma unroll
        for (int i = 0; i < AT_RMAX / 32; ++i) {
          if (tan) ab[i] = ab[i] + adb[i] * (ed[i] - m_e) - ed[i] * m_t;  // w = abar + second-order terms
          s += alv[i] * ab[i];
        }
        s = warp_sum(s);
        __nv_bfloat16* ebh = p.EB + row * p.ldEB;
        __nv_bfloat16* tbh = tan ? p.EB + ((long long)p.row_blk[p.nv - 1] * p.B + b) * p.ldEB : nullptr;
#pragma unroll
        for (int i = 0; i < AT_RMAX / 32; ++i) {
          const int r = lane + 32 * i;
          if (r < R) {
            const float ebar = alv[i] * (ab[i] - s);
            sm.eb[v][r] = ebar;
            __nv_bfloat16 h, l;
            split_bf16(ebar, h, l);
            ebh[r] = h;
            ebh[p.lo_off + r] = l;
            if (tan) {
              const float edb = alv[i] * (adb[i] - m_t);
              split_bf16(edb, h, l);
              tbh[r] = h;
              tbh[p.lo_off + r] = l;
            }
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");   // eb complete; also: nobody still reads w when the next sample overwrites it
      if (p.Pbar) {
        for (int r = wtid; r < R; r += 128) {
          float s = 0.f;
          for (int v = 0; v < ns; ++v) s += sm.eb[v][r];
          atomicAdd(p.Pbar + (long long)b * p.ldP + r, s);   // fire-and-forget reduction (one writer per address and launch)
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");   // eb is rewritten by the next sample
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * R2_TMEM_COLS);
}

static size_t attn_smem_bytes() { return sizeof(AttnSmem) + 128; }

static size_t attn2_smem_bytes() { return sizeof(Attn2Smem) + 1024; }
static int attn_sm_count() {
  static int n = -1;
  if (n < 0) {
    int dev = 0;
    cudaDeviceProp pr;
    n = (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&pr, dev) == cudaSuccess) ? pr.multiProcessorCount : 148;
  }
  return n;
}
// Which family serves a batch of B samples (tools/attn_micro.py, B200): up to two tiles per SM the one-CTA-per-sample
// kernels win (all tiles stream concurrently at two CTAs per SM: 12.1 us vs 15.9 us at B = 256); from three tiles per SM
// on the persistent kernels win (B = 444: 18.3 vs 22.0 us; B = 2368: 79.8 vs 87.2 us = 0.92 of measured HBM).
// SGG_ATTN_PERSIST=0 / 1 forces a family.
static int g_attn_persist = -2;   // -1 = by batch size, 0 / 1 = forced (SGG_ATTN_PERSIST or sgg_set_option("attn_persistent", v))
void attn_set_persistent(int v) { g_attn_persist = v < 0 ? -1 : (v ? 1 : 0); }
static bool attn_persistent(int B) {
  if (g_attn_persist == -2) { const char* e = getenv("SGG_ATTN_PERSIST"); g_attn_persist = e ? (e[0] == '0' ? 0 : 1) : -1; }
  if (g_attn_persist >= 0) return g_attn_persist == 1;
  return B > 2 * attn_sm_count();
}
static bool attn_use_simt() {   // SGG_ATTN_SIMT=1 selects the CUDA-core forward kernel (A/B measurements)
  static int v = -1;
  if (v < 0) { const char* e = getenv("SGG_ATTN_SIMT"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

int attn_fwd(const AttnFwdParams& p_in, int mode, cudaStream_t stream) {
  AttnFwdParams p = p_in;
  p.l2_keep = l2_policy_enabled() ? 1 : 0;
  SGG_CHECK(p.R > 0 && p.R <= AT_RMAX, "attn_fwd: R=%d out of range (1..%d)", p.R, AT_RMAX);
  SGG_CHECK(p.nv >= 1 && p.nv <= (mode == 0 ? AT_MAXV : 4), "attn_fwd: nv=%d out of range", p.nv);
  static bool configured = false;
  if (!configured) {
    SGG_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn_smem_bytes()));
    SGG_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<0, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn_smem_bytes()));
    SGG_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn_smem_bytes()));
    SGG_CUDA(cudaFuncSetAttribute(attn_fwd_mma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn2_smem_bytes()));
    SGG_CUDA(cudaFuncSetAttribute(attn_fwd_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn2_smem_bytes()));
    SGG_CUDA(cudaFuncSetAttribute(attn_fwd_p_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(AttnPSmem) + 1024)));
    SGG_CUDA(cudaFuncSetAttribute(attn_fwd_p_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(AttnPSmem) + 1024)));
    configured = true;
  }
  if (!attn_use_simt()) {
    CUtensorMap tm;
    SGG_TRY(make_tmap_bf16_3d(&tm, p.a, (uint64_t)p.B, (uint64_t)p.R, AT_C, 64, A2_ROWS));
    if (attn_persistent(p.B)) {
      const int grid = p.B < attn_sm_count() ? p.B : attn_sm_count();
      if (mode == 0) SGG_LAUNCH(attn_fwd_p_kernel<0>, grid, A2_THREADS, sizeof(AttnPSmem) + 1024, stream, tm, p);
      else SGG_LAUNCH(attn_fwd_p_kernel<1>, grid, A2_THREADS, sizeof(AttnPSmem) + 1024, stream, tm, p);
      return 0;
    }
    if (mode == 0)
      SGG_LAUNCH(attn_fwd_mma_kernel<0>, p.B, A2_THREADS, attn2_smem_bytes(), stream, tm, p);
    else
      SGG_LAUNCH(attn_fwd_mma_kernel<1>, p.B, A2_THREADS, attn2_smem_bytes(), stream, tm, p);
    return 0;
  }
  if (p.zero_p) {   // the CUDA-core kernels do not carry the fused zero-fill
    for (int v = 0; v < p.nv; ++v)
      SGG_TRY(zero_2d(p.zero_p + (long long)p.row_blk[v] * p.B * p.zero_ld, p.zero_ld, p.zero_cols, p.B, stream));
  }
  if (mode == 0 && p.nv > 4)
    SGG_LAUNCH((attn_fwd_kernel<0, 8>), p.B, AT_THREADS, attn_smem_bytes(), stream, p);
  else if (mode == 0)
    SGG_LAUNCH((attn_fwd_kernel<0, 4>), p.B, AT_THREADS, attn_smem_bytes(), stream, p);
  else
    SGG_LAUNCH((attn_fwd_kernel<1, 4>), p.B, AT_THREADS, attn_smem_bytes(), stream, p);
  return 0;
}

int attn_rev(const AttnRevParams& p_in, cudaStream_t stream) {
  AttnRevParams p = p_in;
  p.l2_keep = l2_policy_enabled() ? 1 : 0;
  SGG_CHECK(p.R > 0 && p.R <= AT_RMAX, "attn_rev: R=%d out of range (1..%d)", p.R, AT_RMAX);
  SGG_CHECK(p.nv >= 1 && p.nv <= AT_MAXV_REV, "attn_rev: nv=%d out of range", p.nv);
  static bool configured = false;
  if (!configured) {
    SGG_CUDA(cudaFuncSetAttribute(attn_rev_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn_smem_bytes()));
    SGG_CUDA(cudaFuncSetAttribute(attn_rev_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(sizeof(AttnRev2Smem) + 1024)));
    SGG_CUDA(cudaFuncSetAttribute(attn_rev_p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(AttnRevPSmem) + 1024)));
    configured = true;
  }
  if (!attn_use_simt()) {
    CUtensorMap tm;
    SGG_TRY(make_tmap_bf16_3d(&tm, p.a, (uint64_t)p.B, (uint64_t)p.R, AT_C, 64, 128));
    if (attn_persistent(p.B)) {
      const int grid = p.B < attn_sm_count() ? p.B : attn_sm_count();
      SGG_LAUNCH(attn_rev_p_kernel, grid, A2_THREADS, sizeof(AttnRevPSmem) + 1024, stream, tm, p);
      return 0;
    }
    SGG_LAUNCH(attn_rev_mma_kernel, p.B, A2_THREADS, sizeof(AttnRev2Smem) + 1024, stream, tm, p);
    return 0;
  }
  SGG_LAUNCH(attn_rev_kernel, p.B, AT_THREADS, attn_smem_bytes(), stream, p);
  return 0;
}

}  // namespace sgg

// Op-level entry point (include/sgg_b200.h): softmax + context reduction for nv streams sharing a tile.
extern "C" int sgg_attn_forward(const void* a, int32_t B, int32_t R, int32_t nv, const float* e, float* alpha,
                                int64_t ld_e, void* z_hl, int64_t ld_z, int64_t lo_off, sgg_stream_t stream) {
  using namespace sgg;
  SGG_CHECK(a && e && alpha && z_hl, "sgg_attn_forward: null argument");
  SGG_CHECK(B >= 1 && ld_e >= R && ld_z >= lo_off + AT_C && lo_off >= AT_C, "sgg_attn_forward: bad shape/pitch");
  SGG_CHECK((ld_z % 2) == 0 && (lo_off % 2) == 0, "sgg_attn_forward: z pitch / lo offset must be even");
  AttnFwdParams p{};
  p.a = reinterpret_cast<const __nv_bfloat16*>(a);
  p.B = B; p.R = R; p.nv = nv;
  for (int v = 0; v < AT_MAXV; ++v) { p.row_blk[v] = v; p.e_blk[v] = v; }
  p.E = e; p.ldE = ld_e;
  p.alpha_out = alpha; p.ldA = ld_e;
  p.X = reinterpret_cast<__nv_bfloat16*>(z_hl); p.ldX = ld_z; p.lo_off = lo_off;
  return attn_fwd(p, 0, reinterpret_cast<cudaStream_t>(stream));
}

// Op-level entry point: reverse of the attention step for nv first-order streams sharing a tile.
extern "C" int sgg_attn_reverse(const void* a, int32_t B, int32_t R, int32_t nv, const float* z_bar, int64_t ld_zb,
                                const float* alpha, int64_t ld_alpha, void* e_bar_hl, int64_t ld_eb, int64_t lo_off,
                                float* p_bar, int64_t ld_p, sgg_stream_t stream) {
  using namespace sgg;
  SGG_CHECK(a && z_bar && alpha && e_bar_hl, "sgg_attn_reverse: null argument");
  SGG_CHECK(B >= 1 && nv >= 1 && nv <= AT_MAXV_REV && ld_zb >= AT_C && ld_alpha >= R && ld_eb >= lo_off + R && lo_off >= R,
            "sgg_attn_reverse: bad shape/pitch");
  SGG_CHECK((ld_zb % 4) == 0 && (reinterpret_cast<uintptr_t>(z_bar) % 16) == 0, "sgg_attn_reverse: z_bar must be 16-byte aligned rows");
  AttnRevParams p{};
  p.a = reinterpret_cast<const __nv_bfloat16*>(a);
  p.B = B; p.R = R; p.nv = nv; p.tan_stream = -1;
  for (int v = 0; v < AT_MAXV_REV; ++v) p.row_blk[v] = v;
  p.XB = z_bar; p.ldXB = ld_zb;
  p.alpha = alpha; p.ldA = ld_alpha;
  p.EB = reinterpret_cast<__nv_bfloat16*>(e_bar_hl); p.ldEB = ld_eb; p.lo_off = lo_off;
  p.Pbar = p_bar; p.ldP = ld_p;
  return attn_rev(p, reinterpret_cast<cudaStream_t>(stream));
}

// LSTM gate GEMM with the LayerNormBasicLSTMCell epilogue fused in (gen:79,87 / disc:81,89):
//   q = [z, u, h] K  on tcgen05 (bf16 hi/lo operands, three products, fp32 accumulators in TMEM), then -- without the
//   pre-activations ever being read back from memory --
//   i, j, f, o = LN(split(q)); c' = c * sigmoid(f + 1) + sigmoid(i) * tanh(j); c_new = LN(c'); h = tanh(c_new) * sigmoid(o)
//   (+ the discriminator's head y = h . w_dec + b, disc:90).
//
// Work decomposition.  A CTA owns a 128-row m-tile and a slice of NPC hidden units, for which it accumulates all four
// gates: its N = 4 * NPC accumulator columns are [i | j | f | o] of those units (the weight shadow `Kp` stores the
// kernel's columns in that interleaved order, see perm_gate_col).  The 512 / NPC CTAs that share an m-tile form one
// thread-block cluster: the cluster is co-scheduled by the hardware and synchronises with barrier.cluster, which is
// what the two layer-norm exchanges need -- every gate's LayerNorm wants the mean / variance of a row over all 512
// units, i.e. over all CTAs of the cluster, and so does the LayerNorm of the new cell state.  Each CTA contributes
// (mean, M2) of its own units per row and gate; the partials are combined with Chan's parallel-variance formula (no
// E[x^2] - mean^2 cancellation; eps = 1e-12 stays meaningful).  The partials travel through a small L2-resident
// scratch: an all-to-all of 64 KB per CTA is served faster by L2 (~60 B/cycle/SM) than by SM-to-SM shared-memory
// reads (~20 B/cycle/SM, B300_MICROARCH.md), and barrier.cluster's release / acquire orders global memory as well.
//   NPC = 32 : 16-CTA clusters (non-portable size), N = 128: m-tiles <= 8 (one cluster per GPC)
//   NPC = 64 :  8-CTA clusters, N = 256                    : larger row counts
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 epilogue (two warps per TMEM lane quadrant, each
// owning half of the CTA's units of a row).
//
// The pre-activations q are still WRITTEN (fp32, TF column order) because the reverse pass recomputes the cell from
// (q, c_in); what disappears is the read-back, the separate cell kernel and its launch on the serial time loop.
#include "common.cuh"
#include "../../include/sgg_b200.h"

namespace sgg {

constexpr int GF_THREADS = 320;
constexpr int GF_EPI = 256;
constexpr int GF_SMEM_BUDGET = 196 * 1024;

// column of gate g, hidden unit u in the interleaved weight shadow Kp (groups of 32 units: [i | j | f | o] x 32)
__host__ __device__ __forceinline__ int perm_gate_col(int g, int u) { return (u >> 5) * 128 + g * 32 + (u & 31); }

__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

struct GatesParams {
  int nrows;                  // rows of this launch (row r of every per-step buffer below is relative to its base pointer)
  int kb;                     // k-blocks of 64 (KXP / 64)
  int a_lo, b_lo;             // k offset of the lo part of X (columns) and of Kp (rows)
  const float* Cin;           // [nrows, 512] fp32
  LstmLN ln;
  float* Q; long long ldQ;    // [nrows, 4*512] fp32 pre-activations (TF order: gate * 512 + unit), for the reverse pass
  float* Cout;                // [nrows, 512]
  __nv_bfloat16* CH; long long ldCH; long long ch_lo;
  __nv_bfloat16* Xn; long long ldX; long long x_lo; int hoff;
  const float* wdec; const float* bdec; float* Y; long long ldY;   // optional D head: Y[row * ldY] += partial dots (zero-filled)
  ZeroRow zero;               // optional: next step's scores rows
  float* scratch;             // [m-tiles][cluster][128][10] fp32 exchange buffer
};

template <int NPC>
__global__ void __launch_bounds__(GF_THREADS, 1)
gates_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GatesParams p) {
  constexpr int CS = 512 / NPC;              // cluster size
  constexpr int N = 4 * NPC;                 // accumulator columns
  constexpr int UPT = NPC / 2;               // units per epilogue thread
  constexpr int A_PART = 128 * 64 * 2;       // 16 KB
  constexpr int B_PART = N * 64 * 2;
  constexpr int STAGE = 2 * A_PART + 2 * B_PART;
  constexpr int STAGES = GF_SMEM_BUDGET / STAGE;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* s_part = reinterpret_cast<float*>(smem + STAGES * STAGE);          // [128][8]  half-1 partials (gates)
  float* s_part2 = s_part + 128 * 8;                                        // [128][2]  half-1 partials (state)
  float* s_ln = s_part2 + 128 * 2;                                          // [10][NPC] gamma[5], beta[5] of this CTA's units
  float* s_wd = s_ln + 10 * NPC;                                            // [NPC] D head weights
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_wd + NPC);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x;                  // unit slice = rank in the cluster (cluster dims = (CS, 1, 1))
  const int m0 = blockIdx.y * 128;
  const int u0 = j * NPC;

  pdl_trigger();
  if (warp == 0 && elect_one()) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
      mbar_init(tmem_full_bar, 1);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, N);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();

  // per-row statistics produced by the epilogue (registers of the epilogue threads)
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int i = 0; i < p.kb; ++i) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = smem + stage * STAGE;
        mbar_expect_tx(&full_bar[stage], STAGE);
        const int kc = i * 64;
        tma_load_2d(st, &tmA, &full_bar[stage], kc, m0);
        tma_load_2d(st + A_PART, &tmA, &full_bar[stage], p.a_lo + kc, m0);
#pragma unroll
        for (int pb = 0; pb < 2; ++pb)
#pragma unroll
          for (int c = 0; c < N / 64; ++c)
            tma_load_2d(st + 2 * A_PART + pb * B_PART + c * (64 * 128), &tmB, &full_bar[stage], j * N + 64 * c, pb * p.b_lo + kc);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc_bf16(128, N, false, true);
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < p.kb; ++i) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sA0 = smem_u32(smem + stage * STAGE), sA1 = sA0 + A_PART;
        const uint32_t sB0 = sA0 + 2 * A_PART, sB1 = sB0 + B_PART;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t da0 = make_smem_desc(sA0 + k * 32, 0, 1024), da1 = make_smem_desc(sA1 + k * 32, 0, 1024);
          const uint64_t db0 = make_smem_desc(sB0 + k * 2048, 64 * 128, 1024), db1 = make_smem_desc(sB1 + k * 2048, 64 * 128, 1024);
          umma_bf16(tmem_base, da0, db0, idesc, (i | k) ? 1u : 0u);
          umma_bf16(tmem_base, da1, db0, idesc, 1u);
          umma_bf16(tmem_base, da0, db1, idesc, 1u);
        }
        umma_commit(&empty_bar[stage]);
        if (i == p.kb - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  }

  // ===================== epilogue (warps 2..9); warps 0 / 1 only take part in the cluster barriers =====================
  const bool epi = warp >= 2;
  const int q = warp & 3;                        // TMEM lane quadrant of this warp
  const int half = epi ? ((warp - 2) >> 2) : 0;  // which half of the CTA's units of the row
  const int rt = q * 32 + lane;                  // row within the m-tile
  const long long row = (long long)m0 + rt;
  const bool row_ok = epi && row < p.nrows;
  const int et = threadIdx.x - 64;               // 0..255 among the epilogue threads
  // TMEM column of (gate g, this thread's unit k): NPC = 32: g*32 + half*16 + k ; NPC = 64: half*128 + g*32 + k
  const uint32_t tcol0 = (NPC == 32) ? (uint32_t)(half * 16) : (uint32_t)(half * 128);
  const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + tcol0;
  const int ut0 = u0 + half * UPT;               // first hidden unit of this thread
  float* scr = p.scratch + ((long long)blockIdx.y * CS) * 128 * 10;   // this m-tile's exchange block: [CS][128][10]

  float gmean[4], grstd[4];
  if (epi) {
    // stage the LN parameters / head weights of this CTA's units while the contraction runs
    for (int i = et; i < 10 * NPC; i += GF_EPI) {
      const int v = i / NPC, u = i - v * NPC;
      s_ln[i] = v < 5 ? p.ln.gamma[v][u0 + u] : p.ln.beta[v - 5][u0 + u];
    }
    if (p.Y) for (int i = et; i < NPC; i += GF_EPI) s_wd[i] = p.wdec[u0 + i];
    if (p.zero.p) {   // this CTA's share of the next split-K output rows of its m-tile
      const int c4 = p.zero.cols >> 2, per = (c4 + CS - 1) / CS;
      const int lo4 = j * per, hi4 = min(c4, lo4 + per);
      for (int r = et >> 3; r < 128; r += GF_EPI >> 3) {
        if (m0 + r >= p.nrows) break;
        float4* d = reinterpret_cast<float4*>(p.zero.p + (long long)(m0 + r) * p.zero.ld);
        for (int c = lo4 + (et & 7); c < hi4; c += 8) d[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    if (row_ok) prefetch_l1(p.Cin + row * 512 + ut0);
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    // ---- pass 1: pre-activations out (for the reverse pass), local (mean, M2) per gate over this thread's units
    float part[8];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float x[UPT];
      if (UPT == 16) {
        uint32_t r[16];
        tmem_ld_32x16(taddr + g * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = __uint_as_float(r[k]);
      } else {
        uint32_t r[32];
        tmem_ld_32x32(taddr + g * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < UPT; ++k) x[k] = __uint_as_float(r[k]);
      }
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < UPT; ++k) s += x[k];
      const float mu = s * (1.0f / UPT);
      float m2 = 0.f;
#pragma unroll
      for (int k = 0; k < UPT; ++k) { const float d = x[k] - mu; m2 = fmaf(d, d, m2); }
      part[2 * g] = mu; part[2 * g + 1] = m2;
      if (row_ok) {
        float4* dst = reinterpret_cast<float4*>(p.Q + row * p.ldQ + g * 512 + ut0);
#pragma unroll
        for (int k = 0; k < UPT; k += 4) dst[k >> 2] = make_float4(x[k], x[k + 1], x[k + 2], x[k + 3]);
      }
    }
    // combine the two halves of the row (Chan), publish this CTA's partial for the row
    if (half == 1) {
      *reinterpret_cast<float4*>(s_part + rt * 8) = make_float4(part[0], part[1], part[2], part[3]);
      *reinterpret_cast<float4*>(s_part + rt * 8 + 4) = make_float4(part[4], part[5], part[6], part[7]);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (half == 0) {
      const float4 o0 = *reinterpret_cast<const float4*>(s_part + rt * 8), o1 = *reinterpret_cast<const float4*>(s_part + rt * 8 + 4);
      const float om[4] = {o0.x, o0.z, o1.x, o1.z}, o2[4] = {o0.y, o0.w, o1.y, o1.w};
      float c[8];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float d = om[g] - part[2 * g];
        c[2 * g] = 0.5f * (om[g] + part[2 * g]);
        c[2 * g + 1] = part[2 * g + 1] + o2[g] + d * d * (0.5f * UPT);
      }
      float* dst = scr + ((long long)j * 128 + rt) * 10;
      *reinterpret_cast<float2*>(dst) = make_float2(c[0], c[1]);
      *reinterpret_cast<float2*>(dst + 2) = make_float2(c[2], c[3]);
      *reinterpret_cast<float2*>(dst + 4) = make_float2(c[4], c[5]);
      *reinterpret_cast<float2*>(dst + 6) = make_float2(c[6], c[7]);
    }
  }
  cluster_arrive_release();
  cluster_wait_acquire();
  float cp[UPT], so[UPT];
  if (epi) {
    // ---- pass 2: full-row statistics of the four gates from the CS partials (NPC units each)
    {
      float mu[4] = {0.f, 0.f, 0.f, 0.f}, pm[4][CS], m2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int s = 0; s < CS; ++s) {
        const float* src = scr + ((long long)s * 128 + rt) * 10;
        const float2 a0 = __ldcg(reinterpret_cast<const float2*>(src)), a1 = __ldcg(reinterpret_cast<const float2*>(src + 2));
        const float2 a2 = __ldcg(reinterpret_cast<const float2*>(src + 4)), a3 = __ldcg(reinterpret_cast<const float2*>(src + 6));
        pm[0][s] = a0.x; pm[1][s] = a1.x; pm[2][s] = a2.x; pm[3][s] = a3.x;
        m2[0] += a0.y; m2[1] += a1.y; m2[2] += a2.y; m2[3] += a3.y;
        mu[0] += a0.x; mu[1] += a1.x; mu[2] += a2.x; mu[3] += a3.x;
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        mu[g] *= (1.0f / CS);
        float dd = 0.f;
#pragma unroll
        for (int s = 0; s < CS; ++s) { const float d = pm[g][s] - mu[g]; dd = fmaf(d, d, dd); }
        const float var = (m2[g] + dd * NPC) * (1.0f / 512.0f);
        gmean[g] = mu[g];
        grstd[g] = 1.0f / sqrtf(var + 1e-12f);
      }
    }
    // ---- pass 3: gate activations and the cell update for this thread's units
    const float* lg = s_ln + half * UPT;          // gamma[v][u] at lg[v * NPC + u], beta at lg[(5 + v) * NPC + u]
    float cin[UPT];
    if (row_ok) {
      const float4* src = reinterpret_cast<const float4*>(p.Cin + row * 512 + ut0);
#pragma unroll
      for (int k = 0; k < UPT; k += 4) { const float4 t = src[k >> 2]; cin[k] = t.x; cin[k + 1] = t.y; cin[k + 2] = t.z; cin[k + 3] = t.w; }
    } else {
#pragma unroll
      for (int k = 0; k < UPT; ++k) cin[k] = 0.f;
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float x[UPT];
      if (UPT == 16) {
        uint32_t r[16];
        tmem_ld_32x16(taddr + g * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = __uint_as_float(r[k]);
      } else {
        uint32_t r[32];
        tmem_ld_32x32(taddr + g * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < UPT; ++k) x[k] = __uint_as_float(r[k]);
      }
      const float fb = (g == 2) ? 1.0f : 0.f;
      const float kk = (g == 1) ? -2.0f : -1.0f;
#pragma unroll
      for (int k = 0; k < UPT; ++k) {
        const float a = fmaf((x[k] - gmean[g]) * grstd[g], lg[g * NPC + k], lg[(5 + g) * NPC + k]) + fb;
        const float e = __expf(kk * fabsf(a));
        const float rr = 1.0f / (1.0f + e);
        const float y = (g == 1) ? copysignf((1.0f - e) * rr, a) : (a >= 0.f ? rr : e * rr);
        // g = 0: cp holds sigmoid(i); g = 1: cp = si * tanh(j); g = 2: cp += c * sigmoid(f + 1); g = 3: so = sigmoid(o)
        if (g == 0) cp[k] = y;
        else if (g == 1) cp[k] *= y;
        else if (g == 2) cp[k] = fmaf(cin[k], y, cp[k]);
        else so[k] = y;
      }
    }
    // local (mean, M2) of c' and the second exchange
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < UPT; ++k) s += cp[k];
    const float mu = s * (1.0f / UPT);
    float m2 = 0.f;
#pragma unroll
    for (int k = 0; k < UPT; ++k) { const float d = cp[k] - mu; m2 = fmaf(d, d, m2); }
    if (half == 1) *reinterpret_cast<float2*>(s_part2 + rt * 2) = make_float2(mu, m2);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (half == 0) {
      const float2 o = *reinterpret_cast<const float2*>(s_part2 + rt * 2);
      const float d = o.x - mu;
      *reinterpret_cast<float2*>(scr + ((long long)j * 128 + rt) * 10 + 8) = make_float2(0.5f * (o.x + mu), m2 + o.y + d * d * (0.5f * UPT));
    }
  }
  cluster_arrive_release();
  cluster_wait_acquire();
  if (epi) {
    float mu = 0.f, m2 = 0.f, pm[CS];
#pragma unroll
    for (int s = 0; s < CS; ++s) {
      const float2 a = __ldcg(reinterpret_cast<const float2*>(scr + ((long long)s * 128 + rt) * 10 + 8));
      pm[s] = a.x; mu += a.x; m2 += a.y;
    }
    mu *= (1.0f / CS);
    float dd = 0.f;
#pragma unroll
    for (int s = 0; s < CS; ++s) { const float d = pm[s] - mu; dd = fmaf(d, d, dd); }
    const float rc = 1.0f / sqrtf((m2 + dd * NPC) * (1.0f / 512.0f) + 1e-12f);
    if (row_ok) {
      const float* lg = s_ln + half * UPT;
      float cn[UPT], h[UPT];
      float sy = 0.f;
#pragma unroll
      for (int k = 0; k < UPT; ++k) {
        cn[k] = fmaf((cp[k] - mu) * rc, lg[4 * NPC + k], lg[9 * NPC + k]);
        h[k] = tanhf_(cn[k]) * so[k];
        if (p.Y) sy = fmaf(h[k], s_wd[half * UPT + k], sy);
      }
      float4* co = reinterpret_cast<float4*>(p.Cout + row * 512 + ut0);
#pragma unroll
      for (int k = 0; k < UPT; k += 4) co[k >> 2] = make_float4(cn[k], cn[k + 1], cn[k + 2], cn[k + 3]);
#pragma unroll
      for (int k = 0; k < UPT; k += 8) {
        uint32_t ch[4], cl[4], hh[4], hl[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          __nv_bfloat16 a0, b0, a1, b1;
          split_bf16(cn[k + 2 * e], a0, b0); split_bf16(cn[k + 2 * e + 1], a1, b1);
          ch[e] = pack_bf16x2(a0, a1); cl[e] = pack_bf16x2(b0, b1);
          split_bf16(h[k + 2 * e], a0, b0); split_bf16(h[k + 2 * e + 1], a1, b1);
          hh[e] = pack_bf16x2(a0, a1); hl[e] = pack_bf16x2(b0, b1);
        }
        if (p.CH) {
          __nv_bfloat16* d = p.CH + row * p.ldCH + ut0 + k;
          *reinterpret_cast<uint4*>(d) = make_uint4(ch[0], ch[1], ch[2], ch[3]);
          *reinterpret_cast<uint4*>(d + p.ch_lo) = make_uint4(cl[0], cl[1], cl[2], cl[3]);
        }
        if (p.Xn) {
          __nv_bfloat16* d = p.Xn + row * p.ldX + p.hoff + ut0 + k;
          *reinterpret_cast<uint4*>(d) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
          *reinterpret_cast<uint4*>(d + p.x_lo) = make_uint4(hl[0], hl[1], hl[2], hl[3]);
        }
      }
      if (p.Y) atomicAdd(p.Y + row * p.ldY, sy + ((j == 0 && half == 0) ? p.bdec[0] : 0.f));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, N);
}

template <int NPC>
static int gates_smem_bytes() {
  constexpr int N = 4 * NPC;
  constexpr int STAGE = 2 * 128 * 64 * 2 + 2 * N * 64 * 2;
  constexpr int STAGES = GF_SMEM_BUDGET / STAGE;
  return STAGES * STAGE + (128 * 8 + 128 * 2 + 10 * NPC + NPC) * 4 + (2 * STAGES + 1) * 8 + 16 + 1024;
}

template <int NPC>
static int launch_gates(const CUtensorMap& tmA, const CUtensorMap& tmB, const GatesParams& p, cudaStream_t stream) {
  auto kern = gates_fused_kernel<NPC>;
  constexpr int CS = 512 / NPC;
  static bool configured = false;
  if (!configured) {
    SGG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gates_smem_bytes<NPC>()));
    if (CS > 8) SGG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(CS, (p.nrows + 127) / 128, 1); cfg.blockDim = dim3(GF_THREADS); cfg.dynamicSmemBytes = gates_smem_bytes<NPC>();
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  note_kernel(reinterpret_cast<const void*>(kern));
  if (timing_enabled()) timing_begin(stream);
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, p);
  if (timing_enabled()) timing_end(stream, reinterpret_cast<const void*>(kern), cfg.gridDim, cfg.blockDim);
  SGG_CUDA(e);
  note_launch();
  return 0;
}

// Which cluster shape can run here: 16-CTA clusters need the opt-in non-portable size and one GPC with 16 free SMs per
// cluster (queried once); 0 = the fused kernel is unavailable (callers fall back to gate GEMM + cell kernel).
static int gates_max_clusters(int npc) {
  static int cached[2] = {-1, -1};
  const int idx = npc == 32 ? 0 : 1;
  if (cached[idx] >= 0) return cached[idx];
  int n = 0;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 512 / npc; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1; cfg.blockDim = dim3(GF_THREADS);
  cudaError_t e;
  if (npc == 32) {
    cudaFuncSetAttribute(gates_fused_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, gates_smem_bytes<32>());
    cudaFuncSetAttribute(gates_fused_kernel<32>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cfg.gridDim = dim3(16, 8, 1); cfg.dynamicSmemBytes = gates_smem_bytes<32>();
    e = cudaOccupancyMaxActiveClusters(&n, gates_fused_kernel<32>, &cfg);
  } else {
    cudaFuncSetAttribute(gates_fused_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, gates_smem_bytes<64>());
    cfg.gridDim = dim3(8, 16, 1); cfg.dynamicSmemBytes = gates_smem_bytes<64>();
    e = cudaOccupancyMaxActiveClusters(&n, gates_fused_kernel<64>, &cfg);
  }
  if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
  cached[idx] = n;
  return n;
}

constexpr int GATES_SCRATCH_FLOATS_PER_MTILE = 16 * 128 * 10;
long long gates_scratch_floats(long long max_rows) { return (max_rows + 127) / 128 * GATES_SCRATCH_FLOATS_PER_MTILE; }

// SGG_FUSED_GATES=0 keeps the gate GEMM and the cell as two kernels (A/B measurements); =16 / =8 force a cluster shape.
static int gates_mode() {
  static int v = -2;
  if (v == -2) { const char* e = getenv("SGG_FUSED_GATES"); v = e ? atoi(e) : -1; }
  return v;
}
bool gates_fused_available() {
  if (gates_mode() == 0) return false;
  return gates_max_clusters(64) > 0 || gates_max_clusters(32) > 0;
}

// X: [nrows, 2*KXP] hi/lo rows of this step; Kp: interleaved weight shadow [2*rK, 2048].
int gates_fused(const __nv_bfloat16* X, long long ldx, int KXP, const __nv_bfloat16* Kp, int rK, GatesParams p, cudaStream_t stream) {
  if (p.nrows <= 0) return 0;
  SGG_CHECK(KXP % 64 == 0, "gates_fused: KXP=%d must be a multiple of 64", KXP);
  p.kb = KXP / 64; p.a_lo = KXP; p.b_lo = rK;
  const int mt = (p.nrows + 127) / 128;
  const int c16 = gates_max_clusters(32), c8 = gates_max_clusters(64);
  int npc;
  if (gates_mode() == 16) npc = 32;
  else if (gates_mode() == 8) npc = 64;
  else if (c16 > 0 && (mt <= c16 || c8 == 0)) npc = 32;            // one wave of 16-CTA clusters: half the MMA time per CTA
  else if (c16 > 0 && c8 > 0 && (mt + c16 - 1) / c16 * 1 <= (mt * 8 + 147) / 148 * 2 - 1) npc = 32;
  else npc = 64;
  SGG_CHECK((npc == 32 ? c16 : c8) > 0, "gates_fused: no cluster configuration can be scheduled on this device");
  CUtensorMap tmA, tmB;
  SGG_TRY(make_tmap_bf16_2d(&tmA, X, (uint64_t)p.nrows, (uint64_t)(2 * KXP), (uint64_t)ldx, 64, 128));
  SGG_TRY(make_tmap_bf16_2d(&tmB, Kp, (uint64_t)(2 * rK), 2048, 2048, 64, 64));
  return npc == 32 ? launch_gates<32>(tmA, tmB, p, stream) : launch_gates<64>(tmA, tmB, p, stream);
}

}  // namespace sgg

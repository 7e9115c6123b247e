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
constexpr int GF_NSTAT = 12;                    // exchange rows per CTA: 8 gate partials, 2 state partials, head partial, pad

// column of gate g, hidden unit u in the interleaved weight shadow Kp (groups of 32 units: [i | j | f | o] x 32)
__host__ __device__ __forceinline__ int perm_gate_col(int g, int u) { return (u >> 5) * 128 + g * 32 + (u & 31); }

// .aligned: every thread of the warp executes the instruction together -> reconverge first (the elected producer /
// issuer lanes and the row-guarded epilogue code leave the warps diverged)
__device__ __forceinline__ void cluster_arrive_release() {
  __syncwarp();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
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
  float* scratch;             // [m-tiles][cluster][GF_NSTAT][128] fp32 exchange buffer
};

template <int NPC>
__global__ void __launch_bounds__(GF_THREADS, 1)
gates_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GatesParams p) {
  // tmA: box {64 k, 128 / CS rows}: every CTA of the cluster fetches 1/CS of the m-tile's A rows per k-block and
  // MULTICASTS it to all CS CTAs (they share the A tile; only the weight columns differ).  Without it the 16 CTAs of a
  // cluster -- one GPC -- pull the same 32 KB sixteen times through that GPC's L2 port, and the port, not the tensor
  // pipe, paces the main loop (measured 0.69 us per k-block against 0.40 us of MMA time, independent of the cluster count).
  constexpr int CS = 512 / NPC;              // cluster size
  constexpr int N = 4 * NPC;                 // accumulator columns
  constexpr int UPT = NPC / 2;               // units per epilogue thread
  constexpr int A_PART = 128 * 64 * 2;       // 16 KB
  constexpr int B_PART = N * 64 * 2;
  constexpr int STAGE = 2 * A_PART + 2 * B_PART;
  constexpr int STAGES = GF_SMEM_BUDGET / STAGE;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared state space
  float* s_part = reinterpret_cast<float*>(smem + STAGES * STAGE);          // [NPC/16][128][8] chunk partials (gates)
  float* s_part2 = s_part + (NPC / 16) * 128 * 8;                           // [NPC/16][128][2] chunk partials (state)
  float* s_ln = s_part2 + (NPC / 16) * 128 * 2;                             // [10][NPC] gamma[5], beta[5] of this CTA's units
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
      // a ring slot is free once ALL CTAs of the cluster have consumed it: peers multicast into it
      for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CS); }
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
  cluster_arrive_release();     // every CTA's barriers exist before a peer signals them
  cluster_wait_acquire();
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
        constexpr int RPS = 128 / CS;                       // A rows this CTA contributes (one or two swizzle atoms)
        constexpr uint16_t ALL = (uint16_t)((1u << CS) - 1u);
        tma_load_2d_mc(st + j * RPS * 128, &tmA, &full_bar[stage], kc, m0 + j * RPS, ALL);
        tma_load_2d_mc(st + A_PART + j * RPS * 128, &tmA, &full_bar[stage], p.a_lo + kc, m0 + j * RPS, ALL);
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
        // descriptors differ between k-steps only in the start-address field (bits 0-13, units of 16 bytes): build one
        // per operand part and k-block, then add the step offset (A, K-major: 32 B per step; B, MN-major: 16 k-rows = 2 KB)
        const uint32_t sA0 = smem_u32(smem + stage * STAGE);
        const uint64_t da0 = make_smem_desc(sA0, 0, 1024), da1 = make_smem_desc(sA0 + A_PART, 0, 1024);
        const uint64_t db0 = make_smem_desc(sA0 + 2 * A_PART, 64 * 128, 1024), db1 = make_smem_desc(sA0 + 2 * A_PART + B_PART, 64 * 128, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          umma_bf16(tmem_base, da0 + 2 * k, db0 + 128 * k, idesc, (i | k) ? 1u : 0u);
          umma_bf16(tmem_base, da1 + 2 * k, db0 + 128 * k, idesc, 1u);
          umma_bf16(tmem_base, da0 + 2 * k, db1 + 128 * k, idesc, 1u);
        }
        umma_commit_mc(&empty_bar[stage], (uint16_t)((1u << CS) - 1u));   // frees the slot in every CTA of the cluster
        if (i == p.kb - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  }

  // ===================== epilogue (warps 2..9); warps 0 / 1 only take part in the cluster barriers =====================
  // Thread = (row of the m-tile = TMEM lane, half of the CTA's units); a thread walks its units in chunks of 16.
  //  * Everything a thread produces for its row is staged in shared memory (the operand ring is free once the last MMA
  //    has retired) and leaves the SM through cooperative stores in which a warp writes whole 128-byte lines: a
  //    row-per-thread store pattern touches 32 partial sectors per instruction (measured: 20 us of epilogue).
  //  * The gate loop and the chunk loop are run-time loops over ONE copy of the cell code: fully unrolled, the epilogue
  //    was instruction-fetch bound (measured: 28 us for 32 units per thread against 7 us for 16).
  //  * The cross-CTA partials are pulled from the L2 scratch by cooperative, fully coalesced float4 loads (one L2
  //    round trip) into shared memory; per-thread scalar pulls serialised ~16 round trips.
  const bool epi = warp >= 2;
  const int q = warp & 3;                        // TMEM lane quadrant of this warp
  const int half = epi ? ((warp - 2) >> 2) : 0;  // which half of the CTA's units of the row
  const int rt = q * 32 + lane;                  // row within the m-tile
  const long long row = (long long)m0 + rt;
  const bool row_ok = epi && row < p.nrows;
  const int et = threadIdx.x - 64;               // 0..255 among the epilogue threads
  const int ew = warp - 2;                       // 0..7
  constexpr int NCH = UPT / 16;                  // 16-unit chunks per thread
  constexpr int NSLOT = NPC / 16;                // 16-unit chunks per row in this CTA
  // local unit index of (half, chunk c, k): lu = half * UPT + c * 16 + k ; its TMEM column for gate g:
  //   NPC = 32: g * 32 + lu ;  NPC = 64: (lu >> 5) * 128 + g * 32 + (lu & 31)
  const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
  // exchange block of this m-tile: [CS][12][128] fp32 (row fastest: coalesced), stat rows 0-7 gates, 8-9 state, 10 head
  float* scr = p.scratch + ((long long)blockIdx.y * CS) * GF_NSTAT * 128;
  constexpr int QS = N + 4;                      // staging pitches (floats): (pitch mod 32) = 4 keeps float4 accesses of
  constexpr int CSP = NPC + 4;                   // a quarter-warp of rows on distinct banks
  float* sQ = reinterpret_cast<float*>(smem);    // [128][QS]   pre-activations (staging for the Q store)
  float* sST = reinterpret_cast<float*>(smem + ((128 * QS * 4 + 1023) & ~1023));   // [CS][8][128] pulled partials
  // NPC = 32: the ring has room for separate c' / sigmoid(o) tiles, so the 64 KB Q store can be issued late (between the
  // arrive and the wait of the SECOND cluster barrier): arrive.release waits for the thread's outstanding global stores,
  // and only the few scratch stores should be in front of it.  NPC = 64: the tiles alias sQ, the Q store stays early.
  constexpr bool kLateQ = (NPC == 32);
  float* sCN = kLateQ ? sST + CS * 8 * 128 : sQ; // [128][CSP]  c' then c_new
  float* sH = sCN + 128 * CSP;                   // [128][CSP]  sigmoid(o) then h
  auto store_q = [&]() {
    // warp = rows ew, ew + 8, ...; a lane covers one float4 of the row's N staged values; NPC consecutive floats belong
    // to one gate (global column gate * 512 + u0 + u)
    for (int r = ew; r < 128; r += 8) {
      if (m0 + r >= p.nrows) break;
      float* qrow = p.Q + (long long)(m0 + r) * p.ldQ + u0;
#pragma unroll
      for (int c = lane * 4; c < N; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(sQ + r * QS + c);
        const int g = c / NPC, u = c - g * NPC;
        *reinterpret_cast<float4*>(qrow + g * 512 + u) = v;
      }
    }
  };

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
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    // ---- pass 1: (mean, M2) per gate and 16-unit chunk; pre-activations into the staging tile
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
      const int lu = half * UPT + c * 16;
      const uint32_t tcol = (NPC == 32) ? (uint32_t)lu : (uint32_t)((lu >> 5) * 128 + (lu & 31));
      float* slot = s_part + ((half * NCH + c) * 128 + rt) * 8;
#pragma unroll 1
      for (int g = 0; g < 4; ++g) {
        uint32_t r[16];
        tmem_ld_32x16(tlane + tcol + g * 32, r);
        tmem_ld_wait();
        float x[16];
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) { x[k] = __uint_as_float(r[k]); sum += x[k]; }
        const float mu = sum * (1.0f / 16.0f);
        float m2 = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) { const float d = x[k] - mu; m2 = fmaf(d, d, m2); }
        *reinterpret_cast<float2*>(slot + 2 * g) = make_float2(mu, m2);
        float4* dst = reinterpret_cast<float4*>(sQ + rt * QS + g * NPC + lu);   // staging column = gate * NPC + unit
#pragma unroll
        for (int k = 0; k < 16; k += 4) dst[k >> 2] = make_float4(x[k], x[k + 1], x[k + 2], x[k + 3]);
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    // combine the row's NSLOT chunk partials (Chan, equal group sizes) and publish them: thread = (row, gate pair)
    {
      float* dst = scr + (long long)j * GF_NSTAT * 128 + rt;
#pragma unroll
      for (int gg = 0; gg < 2; ++gg) {
        const int g = half * 2 + gg;
        float pm[NSLOT], mu = 0.f, m2 = 0.f;
#pragma unroll
        for (int sl = 0; sl < NSLOT; ++sl) {
          const float2 v = *reinterpret_cast<const float2*>(s_part + (sl * 128 + rt) * 8 + 2 * g);
          pm[sl] = v.x; mu += v.x; m2 += v.y;
        }
        mu *= (1.0f / NSLOT);
#pragma unroll
        for (int sl = 0; sl < NSLOT; ++sl) { const float d = pm[sl] - mu; m2 = fmaf(16.0f * d, d, m2); }
        dst[(2 * g) * 128] = mu;
        dst[(2 * g + 1) * 128] = m2;
      }
    }
  }
  cluster_arrive_release();
  float cin_pref[4] = {0.f, 0.f, 0.f, 0.f};
  if (epi) {
    if (!kLateQ) store_q();   // while the cluster gathers at the barrier
    if (row_ok) {   // first quad of c_in: its L2 round trip overlaps the barrier (the rest of the row's line follows it into L1)
      const float4 t = *reinterpret_cast<const float4*>(p.Cin + row * 512 + u0 + half * UPT);
      cin_pref[0] = t.x; cin_pref[1] = t.y; cin_pref[2] = t.z; cin_pref[3] = t.w;
    }
  }
  cluster_wait_acquire();
  if (epi) {
    // ---- pass 2: pull the CS x 8 gate partials of the m-tile (coalesced float4, one round trip), then per-row statistics
    {
      const float4* src = reinterpret_cast<const float4*>(scr);
      float4* dst = reinterpret_cast<float4*>(sST);
#pragma unroll
      for (int i = et; i < CS * 8 * 32; i += GF_EPI) {
        const int sblk = i >> 8, rem = i & 255;                       // 8 stat rows x 32 float4 per CTA block
        dst[i] = __ldcg(src + sblk * (GF_NSTAT * 32) + rem);
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");                    // also: every warp has finished its Q-store rows
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float pm[CS], mu = 0.f, m2 = 0.f;
#pragma unroll
      for (int s2 = 0; s2 < CS; ++s2) {
        pm[s2] = sST[(s2 * 8 + 2 * g) * 128 + rt];
        mu += pm[s2];
        m2 += sST[(s2 * 8 + 2 * g + 1) * 128 + rt];
      }
      mu *= (1.0f / CS);
      float dd = 0.f;
#pragma unroll
      for (int s2 = 0; s2 < CS; ++s2) { const float d = pm[s2] - mu; dd = fmaf(d, d, dd); }
      gmean[g] = mu;
      grstd[g] = fast_rsqrt((m2 + dd * NPC) * (1.0f / 512.0f) + 1e-12f);
    }
    // ---- pass 3: gate activations and the cell update, chunk by chunk; c' and sigmoid(o) go to the staging tiles
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
      const int lu = half * UPT + c * 16;
      const uint32_t tcol = (NPC == 32) ? (uint32_t)lu : (uint32_t)((lu >> 5) * 128 + (lu & 31));
      float cin[16];
      if (row_ok) {
        const float4* src = reinterpret_cast<const float4*>(p.Cin + row * 512 + u0 + lu);
#pragma unroll
        for (int k = 0; k < 16; k += 4) { const float4 t = src[k >> 2]; cin[k] = t.x; cin[k + 1] = t.y; cin[k + 2] = t.z; cin[k + 3] = t.w; }
        if (c == 0) { cin[0] = cin_pref[0]; cin[1] = cin_pref[1]; cin[2] = cin_pref[2]; cin[3] = cin_pref[3]; }
      } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) cin[k] = 0.f;
      }
      float cp[16], so[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) cp[k] = 0.f;
#pragma unroll 1
      for (int g = 0; g < 4; ++g) {
        uint32_t r[16];
        tmem_ld_32x16(tlane + tcol + g * 32, r);
        tmem_ld_wait();
        const float fb = (g == 2) ? 1.0f : 0.f;
        const float kk = (g == 1) ? -2.0f : -1.0f;
        const float mean = g == 0 ? gmean[0] : (g == 1 ? gmean[1] : (g == 2 ? gmean[2] : gmean[3]));
        const float rstd = g == 0 ? grstd[0] : (g == 1 ? grstd[1] : (g == 2 ? grstd[2] : grstd[3]));
        const float* gam = s_ln + g * NPC + lu;
        const float* bet = s_ln + (5 + g) * NPC + lu;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float a = fmaf((__uint_as_float(r[k]) - mean) * rstd, gam[k], bet[k]) + fb;
          const float e = fast_exp(kk * fabsf(a));
          const float rr = fast_rcp(1.0f + e);
          const float y = (g == 1) ? copysignf((1.0f - e) * rr, a) : (a >= 0.f ? rr : e * rr);
          // g = 0: cp = sigmoid(i); g = 1: cp *= tanh(j); g = 2: cp += c * sigmoid(f + 1); g = 3: so = sigmoid(o)
          cp[k] = g == 0 ? y : (g == 1 ? cp[k] * y : (g == 2 ? fmaf(cin[k], y, cp[k]) : cp[k]));
          so[k] = y;
        }
      }
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) sum += cp[k];
      const float mu = sum * (1.0f / 16.0f);
      float m2 = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) { const float d = cp[k] - mu; m2 = fmaf(d, d, m2); }
      *reinterpret_cast<float2*>(s_part2 + ((half * NCH + c) * 128 + rt) * 2) = make_float2(mu, m2);
#pragma unroll
      for (int k = 0; k < 16; k += 4) {
        *reinterpret_cast<float4*>(sCN + rt * CSP + lu + k) = make_float4(cp[k], cp[k + 1], cp[k + 2], cp[k + 3]);
        *reinterpret_cast<float4*>(sH + rt * CSP + lu + k) = make_float4(so[k], so[k + 1], so[k + 2], so[k + 3]);
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (half == 0) {   // row partial of the state LayerNorm
      float pm[NSLOT], mu = 0.f, m2 = 0.f;
#pragma unroll
      for (int sl = 0; sl < NSLOT; ++sl) {
        const float2 v = *reinterpret_cast<const float2*>(s_part2 + (sl * 128 + rt) * 2);
        pm[sl] = v.x; mu += v.x; m2 += v.y;
      }
      mu *= (1.0f / NSLOT);
#pragma unroll
      for (int sl = 0; sl < NSLOT; ++sl) { const float d = pm[sl] - mu; m2 = fmaf(16.0f * d, d, m2); }
      float* dst = scr + (long long)j * GF_NSTAT * 128 + rt;
      dst[8 * 128] = mu;
      dst[9 * 128] = m2;
    }
  }
  cluster_arrive_release();
  if (epi && kLateQ) store_q();
  cluster_wait_acquire();
  if (epi) {
    float rc, cmean;
    {
      float pm[CS], mu = 0.f, m2 = 0.f;
#pragma unroll
      for (int s2 = 0; s2 < CS; ++s2) {
        const float* src = scr + (long long)s2 * GF_NSTAT * 128 + rt;
        pm[s2] = __ldcg(src + 8 * 128);
        m2 += __ldcg(src + 9 * 128);
        mu += pm[s2];
      }
      mu *= (1.0f / CS);
      float dd = 0.f;
#pragma unroll
      for (int s2 = 0; s2 < CS; ++s2) { const float d = pm[s2] - mu; dd = fmaf(d, d, dd); }
      cmean = mu;
      rc = fast_rsqrt((m2 + dd * NPC) * (1.0f / 512.0f) + 1e-12f);
    }
    // new cell state and hidden state of this thread's units, in place in the staging tiles
    float sy = 0.f;
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
      const int lu = half * UPT + c * 16;
      const float* gam = s_ln + 4 * NPC + lu;
      const float* bet = s_ln + 9 * NPC + lu;
#pragma unroll
      for (int k = 0; k < 16; k += 4) {
        const float4 cpv = *reinterpret_cast<const float4*>(sCN + rt * CSP + lu + k), sov = *reinterpret_cast<const float4*>(sH + rt * CSP + lu + k);
        const float cpa[4] = {cpv.x, cpv.y, cpv.z, cpv.w}, soa[4] = {sov.x, sov.y, sov.z, sov.w};
        float cn[4], h[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          cn[e] = fmaf((cpa[e] - cmean) * rc, gam[k + e], bet[k + e]);
          h[e] = tanhf_(cn[e]) * soa[e];
          if (p.Y) sy = fmaf(h[e], s_wd[lu + k + e], sy);
        }
        *reinterpret_cast<float4*>(sCN + rt * CSP + lu + k) = make_float4(cn[0], cn[1], cn[2], cn[3]);
        *reinterpret_cast<float4*>(sH + rt * CSP + lu + k) = make_float4(h[0], h[1], h[2], h[3]);
      }
    }
    if (p.Y && half == 1) s_part2[rt] = sy;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (p.Y && half == 0) scr[((long long)j * GF_NSTAT + 10) * 128 + rt] = sy + s_part2[rt];   // head partial of this CTA
    // ---- cooperative stores: 8 lanes cover one row's NPC floats (4 floats each for NPC = 32, 8 for NPC = 64)
    constexpr int FPL = NPC / 8;                   // floats per lane: 4 or 8
    for (int r = ew * 4 + (lane >> 3); r < 128; r += 32) {
      const long long grow = (long long)m0 + r;
      if (grow >= p.nrows) continue;
      const int c0 = (lane & 7) * FPL;
#pragma unroll
      for (int k = 0; k < FPL; k += 4) {
        const float4 a4 = *reinterpret_cast<const float4*>(sCN + r * CSP + c0 + k), b4 = *reinterpret_cast<const float4*>(sH + r * CSP + c0 + k);
        const float cn[4] = {a4.x, a4.y, a4.z, a4.w}, h[4] = {b4.x, b4.y, b4.z, b4.w};
        *reinterpret_cast<float4*>(p.Cout + grow * 512 + u0 + c0 + k) = a4;
        __nv_bfloat16 a[4], b[4], cc[4], d[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { split_bf16(cn[e], a[e], b[e]); split_bf16(h[e], cc[e], d[e]); }
        if (p.CH) {
          __nv_bfloat16* dst = p.CH + grow * p.ldCH + u0 + c0 + k;
          *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]));
          *reinterpret_cast<uint2*>(dst + p.ch_lo) = make_uint2(pack_bf16x2(b[0], b[1]), pack_bf16x2(b[2], b[3]));
        }
        if (p.Xn) {   // the h columns start at C + U (812 for the discriminator): 8-byte aligned only
          __nv_bfloat16* dst = p.Xn + grow * p.ldX + p.hoff + u0 + c0 + k;
          *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(cc[0], cc[1]), pack_bf16x2(cc[2], cc[3]));
          *reinterpret_cast<uint2*>(dst + p.x_lo) = make_uint2(pack_bf16x2(d[0], d[1]), pack_bf16x2(d[2], d[3]));
        }
      }
    }
  }
  if (p.Y) {   // the discriminator's head (disc:90): CTA 0 of the cluster sums the per-CTA partial dot products (no atomics)
    cluster_arrive_release();
    cluster_wait_acquire();
    if (j == 0 && epi && half == 0 && row_ok) {
      float y = p.bdec[0];
#pragma unroll
      for (int s2 = 0; s2 < CS; ++s2) y += __ldcg(scr + ((long long)s2 * GF_NSTAT + 10) * 128 + rt);
      p.Y[row * p.ldY] = y;
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
  return STAGES * STAGE + ((NPC / 16) * 128 * 10 + 10 * NPC + NPC) * 4 + (2 * STAGES + 1) * 8 + 16 + 1024;
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

constexpr int GATES_SCRATCH_FLOATS_PER_MTILE = 16 * GF_NSTAT * 128;
long long gates_scratch_floats(long long max_rows) { return (max_rows + 127) / 128 * GATES_SCRATCH_FLOATS_PER_MTILE; }

// Selection.  Measured on B200 at config 2 (profiles/README.md, round 2): the fused kernel is on par with, not faster
// than, the split-K gate GEMM + cell kernel pair (5.45 vs 5.32 ms per iteration): full-K accumulation per CTA limits the
// grid to 96 CTAs whose operand feed (64 KB per k-block and SM) runs at the per-SM L2 -> shared-memory rate, and the
// LayerNorm exchange adds three cluster barriers in series.  So the pair stays the default and the fused kernel is
// selected with SGG_FUSED_GATES=1 (=16 / =8 force a cluster shape) or sgg_set_option("fused_gates", v).
static int g_gates_mode = -2;
static int gates_mode() {
  if (g_gates_mode == -2) { const char* e = getenv("SGG_FUSED_GATES"); g_gates_mode = e ? atoi(e) : 0; }
  return g_gates_mode;
}
void gates_set_mode(int v) { g_gates_mode = v; }
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
  SGG_TRY(make_tmap_bf16_2d(&tmA, X, (uint64_t)p.nrows, (uint64_t)(2 * KXP), (uint64_t)ldx, 64, npc == 32 ? 8 : 16));
  SGG_TRY(make_tmap_bf16_2d(&tmB, Kp, (uint64_t)(2 * rK), 2048, 2048, 64, 64));
  return npc == 32 ? launch_gates<32>(tmA, tmB, p, stream) : launch_gates<64>(tmA, tmB, p, stream);
}

}  // namespace sgg

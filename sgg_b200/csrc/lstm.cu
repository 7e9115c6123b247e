// LayerNormBasicLSTMCell(512) pointwise stages (gen:79,87 / disc:81,89):
//   lstm_fwd : i,j,f,o = LN(split(q)); c' = c*sigmoid(f+1) + sigmoid(i)*tanh(j); c_new = LN(c');
//              h = tanh(c_new)*sigmoid(o); optional D head y = h.w_dec + b (disc:90)
//   lstm_tan : forward tangent (JVP) of the same, for the WGAN-GP interpolate stream
//   lstm_rev : reverse pass, first order or reverse-over-(primal+tangent); also LN gamma/beta
//              gradients and the D head's w_dec / b_dec gradients.
// The gate pre-activations q = [z,u,h] K come from the tcgen05 GEMM; all statistics stay fp32
// (LN eps = 1e-12).  Everything is recomputed from (q, c_in) in the reverse pass.
//
// Work decomposition: one CTA of 4 warps per row.  The rows are few (a few hundred) and each row is a long
// dependent chain of 512-wide reductions, so the row is split two ways:
//   gate phases  (A: LN + nonlinearity of one gate, C: its reverse)  -> warp G owns gate G, lane owns 16 columns
//   state phase  (B: cell update, LN(state), h and their reverse)    -> thread owns 4 columns, block reductions
// with the gate activations / their adjoints exchanged through shared memory.
#include "common.cuh"
#include "../../include/sgg_b200.h"

namespace sgg {

constexpr int LH = 512;             // LSTM units
constexpr int LS_THREADS = 128;     // 4 warps: one per gate
constexpr float LN_EPS_F = 1e-12f;
constexpr float FORGET_BIAS_F = 1.0f;
constexpr float INV_LH = 1.0f / LH;

typedef float V16[16];

// ---- warp-per-gate helpers: lane owns columns lane*4 + 128*i + e
__device__ __forceinline__ void ld16(const float* base, V16& v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = *reinterpret_cast<const float4*>(base + lane * 4 + 128 * i);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
__device__ __forceinline__ void st16(float* base, const V16& v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4*>(base + lane * 4 + 128 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
// hi at base[...], lo at base[lo_off + ...]
__device__ __forceinline__ void st16_hl(__nv_bfloat16* base, long long lo_off, const V16& v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split_bf16(v[4 * i + e], h[e], l[e]);
    *reinterpret_cast<uint2*>(base + lane * 4 + 128 * i) = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
    *reinterpret_cast<uint2*>(base + lo_off + lane * 4 + 128 * i) =
        make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
  }
}
__device__ __forceinline__ void acc16(float* base, const V16& v) {   // base[col] += v (lane-owned columns)
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float4* p = reinterpret_cast<float4*>(base + lane * 4 + 128 * i);
    float4 t = *p;
    t.x += v[4 * i]; t.y += v[4 * i + 1]; t.z += v[4 * i + 2]; t.w += v[4 * i + 3];
    *p = t;
  }
}
__device__ __forceinline__ float sum16(const V16& a) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) s += a[j];
  return warp_sum(s);
}
// n = (x - mean) * r, r = rsqrt(var + eps)   (tf.contrib.layers.layer_norm, biased variance)
__device__ __forceinline__ void ln_norm(const V16& x, V16& n, float& r) {
  const float mean = sum16(x) * INV_LH;
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    n[j] = x[j] - mean;
    s = fmaf(n[j], n[j], s);
  }
  const float var = warp_sum(s) * INV_LH;
  r = fast_rsqrt(var + LN_EPS_F);
#pragma unroll
  for (int j = 0; j < 16; ++j) n[j] *= r;
}
// out = r * (w - mean(w) - n * mean(n*w)) : LN JVP and VJP (the Jacobian is symmetric)
__device__ __forceinline__ void ln_proj(const V16& n, float r, const V16& w, V16& out) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) { s0 += w[j]; s1 = fmaf(n[j], w[j], s1); }
  s0 = warp_sum(s0) * INV_LH;
  s1 = warp_sum(s1) * INV_LH;
#pragma unroll
  for (int j = 0; j < 16; ++j) out[j] = r * (w[j] - s0 - n[j] * s1);
}

// ---- state-phase helpers: thread owns columns tid*4 .. tid*4+3
struct F4 { float v[4]; };
__device__ __forceinline__ F4 ld4(const float* base) {
  const float4 t = *reinterpret_cast<const float4*>(base + threadIdx.x * 4);
  return F4{{t.x, t.y, t.z, t.w}};
}
__device__ __forceinline__ F4 ld4_or_zero(const float* base) {
  if (base) return ld4(base);
  return F4{{0.f, 0.f, 0.f, 0.f}};
}
__device__ __forceinline__ void st4(float* base, const F4& a) {
  *reinterpret_cast<float4*>(base + threadIdx.x * 4) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
}
__device__ __forceinline__ void st4_hl(__nv_bfloat16* base, long long lo_off, const F4& a) {
  __nv_bfloat16 h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) split_bf16(a.v[e], h[e], l[e]);
  *reinterpret_cast<uint2*>(base + threadIdx.x * 4) = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
  *reinterpret_cast<uint2*>(base + lo_off + threadIdx.x * 4) = make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
}
// Sum of NV per-thread values over the 128 threads of the CTA.  `red` holds 2 x 4 x 8 floats; consecutive calls
// alternate between its halves (tracked by `tog`), so one barrier per call suffices.
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* red, int& tog) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* buf = red + tog * 32;
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) buf[warp * 8 + k] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = buf[k] + buf[8 + k] + buf[16 + k] + buf[24 + k];
  tog ^= 1;
}

struct LstmLN {
  const float* gamma[5];  // input, transform, forget, output, state
  const float* beta[5];
};

// Gate phase A for gate G (= this warp): LN + nonlinearity, optionally with the tangent.
//   n, r      : normalised pre-activation and 1/std (kept for the reverse gate phase)
//   xd, nd    : tangent pre-activation and its LN tangent (TAN only)
//   act, actd : gate activation and its tangent
// G (= warp index) is a run-time value: one copy of the gate code serves the four gates (the kernels were instruction-
// fetch bound with one specialised copy per gate).  sigmoid and tanh share one exponential and one reciprocal:
//   e = exp(-k |a|), r = 1/(1+e);  tanh(a) = sign(a) (1-e) r  (k = 2);  sigmoid(a) = a >= 0 ? r : e r  (k = 1).
template <bool TAN>
__device__ __forceinline__ void gate_fwd(int G, const float* q, const float* qd, const LstmLN& ln, V16& n, float& r,
                                         V16& xd, V16& nd, V16& act, V16& actd) {
  V16 x, g, bt;
  ld16(q + G * LH, x);
  if (TAN) ld16(qd + G * LH, xd);
  ld16(ln.gamma[G], g);
  ld16(ln.beta[G], bt);
  ln_norm(x, n, r);
  if (TAN) ln_proj(n, r, xd, nd);
  const bool is_tanh = (G == 1);
  const float fb = (G == 2) ? FORGET_BIAS_F : 0.f;
  const float k = is_tanh ? -2.0f : -1.0f;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float a = fmaf(n[j], g[j], bt[j]) + fb;
    const float e = fast_exp(k * fabsf(a));
    const float rr = fast_rcp(1.0f + e);
    const float y = is_tanh ? copysignf((1.0f - e) * rr, a) : (a >= 0.f ? rr : e * rr);
    act[j] = y;
    if (TAN) actd[j] = (is_tanh ? (1.f - y * y) : y * (1.f - y)) * nd[j] * g[j];
  }
}

// shared memory: gate activations [4][512] (+ tangents [4][512]) + reduction scratch
// Zero-fill of the NEXT split-K GEMM's output row that belongs to the row this CTA processes (the GEMM then accumulates
// with reductions and needs no zero-fill launch of its own): `cols` floats (multiple of 4) at zp + row * ld.
struct ZeroRow { float* p; long long ld; int cols; };
__device__ __forceinline__ void zero_row(const ZeroRow& z, long long row) {
  if (z.p == nullptr) return;
  float4* d = reinterpret_cast<float4*>(z.p + row * z.ld);
  for (int i = threadIdx.x; i < (z.cols >> 2); i += LS_THREADS) d[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

struct LstmSmemFwd {
  float xa[4][LH];
  float xad[4][LH];
  float red[64];
};

// ------------------------------------------------------------------------------------ fwd
struct LstmFwdParams {
  int nrows;
  const float* Q; long long ldQ;          // [rows, 4H] gate pre-activations
  const float* Cin;                       // [rows, H] fp32
  LstmLN ln;
  float* Cout;                            // [rows, H] fp32 new c
  __nv_bfloat16* CH; long long ldCH; long long ch_lo;       // new c hi/lo (A operand of e = P + c W_h)
  __nv_bfloat16* Xn; long long ldX; long long x_lo; int hoff;  // h hi/lo into next step's x buffer
  const float* wdec; const float* bdec;   // optional D head (fp32 master weights)
  float* Y; long long ldY;                // Y[row*ldY]
  ZeroRow zero;                           // optional: next step's scores rows
};

__global__ void __launch_bounds__(LS_THREADS) lstm_fwd_kernel(const LstmFwdParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) LstmSmemFwd sm;
  const int G = threadIdx.x >> 5;
  int tog = 0;
  for (int row = blockIdx.x; row < p.nrows; row += gridDim.x) {
    prefetch_l1(p.Cin + (long long)row * LH + threadIdx.x * 4);   // state-phase operand: fetched under the gate phase
    zero_row(p.zero, row);
    {
      V16 n, xd, nd, act, actd;
      float r;
      const float* q = p.Q + (long long)row * p.ldQ;
      gate_fwd<false>(G, q, q, p.ln, n, r, xd, nd, act, actd);
      st16(sm.xa[G], act);
    }
    __syncthreads();
    const F4 si = ld4(sm.xa[0]), tj = ld4(sm.xa[1]), sf = ld4(sm.xa[2]), so = ld4(sm.xa[3]);
    const F4 c = ld4(p.Cin + (long long)row * LH);
    const F4 g4 = ld4(p.ln.gamma[4]), b4 = ld4(p.ln.beta[4]);
    F4 cp;
    float s1[1] = {0.f};
#pragma unroll
    for (int e = 0; e < 4; ++e) { cp.v[e] = fmaf(c.v[e], sf.v[e], si.v[e] * tj.v[e]); s1[0] += cp.v[e]; }
    block_sum(s1, sm.red, tog);
    const float mean = s1[0] * INV_LH;
    float s2[1] = {0.f};
#pragma unroll
    for (int e = 0; e < 4; ++e) { cp.v[e] -= mean; s2[0] = fmaf(cp.v[e], cp.v[e], s2[0]); }
    block_sum(s2, sm.red, tog);
    const float rc = fast_rsqrt(s2[0] * INV_LH + LN_EPS_F);
    F4 cn, h;
    float sy[1] = {0.f};
    const F4 wd = p.Y ? ld4(p.wdec) : F4{{0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      cn.v[e] = fmaf(cp.v[e] * rc, g4.v[e], b4.v[e]);
      h.v[e] = tanhf_(cn.v[e]) * so.v[e];
      sy[0] = fmaf(h.v[e], wd.v[e], sy[0]);
    }
    st4(p.Cout + (long long)row * LH, cn);
    if (p.CH) st4_hl(p.CH + (long long)row * p.ldCH, p.ch_lo, cn);
    if (p.Xn) st4_hl(p.Xn + (long long)row * p.ldX + p.hoff, p.x_lo, h);
    if (p.Y) {
      block_sum(sy, sm.red, tog);
      if (threadIdx.x == 0) p.Y[(long long)row * p.ldY] = sy[0] + p.bdec[0];
    }
    __syncthreads();   // xa is rewritten by the next row
  }
}

// ------------------------------------------------------------------------------------ tangent
struct LstmTanParams {
  int nrows;                 // tangent rows; primal row = prow0 + i, tangent row = trow0 + i
  int prow0, trow0;
  const float* Q; long long ldQ;
  const float* C;            // [rows,H] fp32: primal c_in at prow, tangent c_in at trow
  LstmLN ln;
  float* Cout;               // tangent new c written at trow
  __nv_bfloat16* CH; long long ldCH; long long ch_lo;
  __nv_bfloat16* Xn; long long ldX; long long x_lo; int hoff;
  ZeroRow zero;              // optional: next step's tangent-score rows (indexed by i, not by trow)
};

// State phase shared by the tangent and reverse kernels: from the gate activations in shared memory and c_in,
// recompute c', LN(state) and (TAN) their tangents.
struct StateFwd {
  F4 si, tj, sf, so, sid, tjd, sfd, sod;   // gate activations and tangents
  F4 c, cd;                                // c_in and its tangent
  F4 nc, ncd, cpd;                         // normalised state, its tangent, tangent of c'
  float rc;
  float qq;                                // sum(nc * cpd)
};
template <bool TAN>
__device__ __forceinline__ void state_fwd(const LstmSmemFwd& sm, const float* c_in, const float* cd_in, StateFwd& s,
                                          float* red, int& tog) {
  s.si = ld4(sm.xa[0]); s.tj = ld4(sm.xa[1]); s.sf = ld4(sm.xa[2]); s.so = ld4(sm.xa[3]);
  s.c = ld4(c_in);
  if (TAN) {
    s.sid = ld4(sm.xad[0]); s.tjd = ld4(sm.xad[1]); s.sfd = ld4(sm.xad[2]); s.sod = ld4(sm.xad[3]);
    s.cd = ld4(cd_in);
  }
  F4 cp;
  float s1[2] = {0.f, 0.f};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    cp.v[e] = fmaf(s.c.v[e], s.sf.v[e], s.si.v[e] * s.tj.v[e]);
    s1[0] += cp.v[e];
    if (TAN) {
      s.cpd.v[e] = s.cd.v[e] * s.sf.v[e] + s.c.v[e] * s.sfd.v[e] + s.sid.v[e] * s.tj.v[e] + s.si.v[e] * s.tjd.v[e];
      s1[1] += s.cpd.v[e];
    }
  }
  block_sum(s1, red, tog);
  const float mean = s1[0] * INV_LH;
  float s2[1] = {0.f};
#pragma unroll
  for (int e = 0; e < 4; ++e) { cp.v[e] -= mean; s2[0] = fmaf(cp.v[e], cp.v[e], s2[0]); }
  block_sum(s2, red, tog);
  s.rc = fast_rsqrt(s2[0] * INV_LH + LN_EPS_F);
#pragma unroll
  for (int e = 0; e < 4; ++e) s.nc.v[e] = cp.v[e] * s.rc;
  if (TAN) {
    float s3[1] = {0.f};
#pragma unroll
    for (int e = 0; e < 4; ++e) s3[0] = fmaf(s.nc.v[e], s.cpd.v[e], s3[0]);
    block_sum(s3, red, tog);
    s.qq = s3[0];
    const float m0 = s1[1] * INV_LH, m1 = s3[0] * INV_LH;
#pragma unroll
    for (int e = 0; e < 4; ++e) s.ncd.v[e] = s.rc * (s.cpd.v[e] - m0 - s.nc.v[e] * m1);
  }
}

__global__ void __launch_bounds__(LS_THREADS) lstm_tan_kernel(const LstmTanParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) LstmSmemFwd sm;
  const int G = threadIdx.x >> 5;
  int tog = 0;
  for (int i = blockIdx.x; i < p.nrows; i += gridDim.x) {
    const long long prow = p.prow0 + i, trow = p.trow0 + i;
    prefetch_l1(p.C + prow * LH + threadIdx.x * 4);
    prefetch_l1(p.C + trow * LH + threadIdx.x * 4);
    zero_row(p.zero, i);
    {
      V16 n, xd, nd, act, actd;
      float r;
      const float* q = p.Q + prow * p.ldQ;
      const float* qd = p.Q + trow * p.ldQ;
      gate_fwd<true>(G, q, qd, p.ln, n, r, xd, nd, act, actd);
      st16(sm.xa[G], act);
      st16(sm.xad[G], actd);
    }
    __syncthreads();
    StateFwd s;
    state_fwd<true>(sm, p.C + prow * LH, p.C + trow * LH, s, sm.red, tog);
    const F4 g4 = ld4(p.ln.gamma[4]), b4 = ld4(p.ln.beta[4]);
    F4 cnd, hd;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float cn = fmaf(s.nc.v[e], g4.v[e], b4.v[e]);
      const float tc = tanhf_(cn);
      cnd.v[e] = s.ncd.v[e] * g4.v[e];
      hd.v[e] = (1.f - tc * tc) * cnd.v[e] * s.so.v[e] + tc * s.sod.v[e];
    }
    st4(p.Cout + trow * LH, cnd);
    if (p.CH) st4_hl(p.CH + trow * p.ldCH, p.ch_lo, cnd);
    if (p.Xn) st4_hl(p.Xn + trow * p.ldX + p.hoff, p.x_lo, hd);
    __syncthreads();   // xa / xad are rewritten by the next row
  }
}

// ------------------------------------------------------------------------------------ reverse
// One launch handles `n_tan` rows that carry a tangent (primal row tan_prow0 + j, tangent row trow0 + j) and
// `n_plain` first-order rows (row = prow0 + i).  Parameter gradients (5 LN gammas/betas, D head) accumulate in
// shared memory per CTA and are flushed to a per-CTA slice of `partials` ([grid][LR_NPART] fp32, overwritten when
// init_partials, else accumulated: the launches of one reverse pass use the same grid); lngrad_reduce() sums
// the slices into the gradient bucket once per pass.  No atomics anywhere.
struct LstmRevParams {
  int n_plain, prow0;
  int n_tan, tan_prow0, trow0;
  int B;                     // rows per stream block (for ybar_blk lookup)
  const float* Q; long long ldQ;
  const float* C;            // [rows,H] c_in (primal and tangent rows)
  LstmLN ln;
  // upstream adjoints
  const float* XBn; long long ldXB; int hoff;   // next step's x_bar (h columns), nullable
  const float* HB; long long ldHB;              // extra h_bar [rows,H] (G: dfake W_dec^T), nullable
  const float* CBn;                             // next step's total c_bar [rows,H], nullable
  float ybar_blk[8];                            // D head: y_bar per stream block (0 if unused)
  float ydot_bar;                               // D head tangent adjoint (lambda) for the tangent rows
  const float* wdec;                            // D head weights (nullable)
  // outputs
  __nv_bfloat16* QB; long long ldQB; long long qb_lo;   // q_bar hi/lo [rows, 2*4H]
  float* CB;                                    // c_bar [rows,H]
  // parameter gradients (partials == nullptr => data path only)
  float* partials; int init_partials;
  ZeroRow zero;                                 // optional: this step's x_bar rows (primal and tangent rows)
};

constexpr int LR_NVEC = 11;                 // dgamma[5], dbeta[5], dwdec
constexpr int LR_NPART = LR_NVEC * LH + 4;  // + dbdec (padded)
constexpr int LR_MAX_GRID = 444;            // 3 CTAs per SM

struct LstmSmemRev {
  LstmSmemFwd f;
  float yb[4][LH];     // adjoints of the gate activations
  float ydb[4][LH];    // ... and of their tangents
};
// Per-thread accumulators of the parameter gradients over the rows of a CTA: the gate phases always give a
// lane the same 16 columns of gate G, the state phase always gives a thread the same 4 columns.
struct RevAcc {
  V16 dg, db;          // d gamma[G], d beta[G] (lane-owned columns)
  F4 sg, sb, dw;       // d gamma[state], d beta[state], d w_dec (thread-owned columns)
  float dbd;           // d b_dec (thread 0)
};

// Gate phase C for gate G (= this warp): reverse through the nonlinearity and its LayerNorm.
template <bool TAN>
__device__ __forceinline__ void gate_rev(int G, const LstmRevParams& p, LstmSmemRev& sm, long long prow, long long trow,
                                         const V16& n, float r, const V16& xd, const V16& nd, RevAcc& acc) {
  V16 g, act, ybv, ydbv, nb, ndb, dgs;
  ld16(p.ln.gamma[G], g);
  ld16(sm.f.xa[G], act);
  ld16(sm.yb[G], ybv);
  if (TAN) ld16(sm.ydb[G], ydbv);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float y = act[j];
    const float d1 = (G == 1) ? (1.f - y * y) : y * (1.f - y);
    float abar = ybv[j] * d1;
    if (TAN) {
      const float d2 = (G == 1) ? (-2.f * y * d1) : d1 * (1.f - 2.f * y);
      abar += ydbv[j] * d2 * (nd[j] * g[j]);
      const float adb = ydbv[j] * d1;
      ndb[j] = adb * g[j];
      dgs[j] = abar * n[j] + adb * nd[j];
    } else {
      dgs[j] = abar * n[j];
    }
    ybv[j] = abar;           // = d beta contribution
    nb[j] = abar * g[j];
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) { acc.dg[j] += dgs[j]; acc.db[j] += ybv[j]; }
  V16 xb;
  if (!TAN) {
    ln_proj(n, r, nb, xb);
  } else {
    // six row sums at once: sum(nb), sum(n nb), sum(ndb), pp = sum(n ndb), ss = sum(ndb nd), qq = sum(n xd)
    float s[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      s[0] += nb[j]; s[1] = fmaf(n[j], nb[j], s[1]);
      s[2] += ndb[j]; s[3] = fmaf(n[j], ndb[j], s[3]);
      s[4] = fmaf(ndb[j], nd[j], s[4]); s[5] = fmaf(n[j], xd[j], s[5]);
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) s[k] = warp_sum(s[k]);
    const float kk = r * INV_LH;
    V16 xdb;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      xdb[j] = r * (ndb[j] - s[2] * INV_LH - n[j] * s[3] * INV_LH);
      xb[j] = r * (nb[j] - s[0] * INV_LH - n[j] * s[1] * INV_LH) - kk * (n[j] * s[4] + s[5] * xdb[j] + s[3] * nd[j]);
    }
    st16_hl(p.QB + trow * p.ldQB + G * LH, p.qb_lo, xdb);
  }
  st16_hl(p.QB + prow * p.ldQB + G * LH, p.qb_lo, xb);
}

template <bool TAN>
__device__ __forceinline__ void lstm_rev_row(const LstmRevParams& p, LstmSmemRev& sm, long long prow, long long trow,
                                             RevAcc& acc, int& tog) {
  const int G = threadIdx.x >> 5;
  const float* q = p.Q + prow * p.ldQ;
  const float* qd = p.Q + trow * p.ldQ;
  zero_row(p.zero, prow);
  if (TAN) zero_row(p.zero, trow);
  {  // operands of the state phase: fetched into L1 under the gate phase (each is one 2 KB row, 16 B per thread)
    const int c4 = threadIdx.x * 4;
    prefetch_l1(p.C + prow * LH + c4);
    if (p.XBn) prefetch_l1(p.XBn + prow * p.ldXB + p.hoff + c4);
    if (p.HB) prefetch_l1(p.HB + prow * p.ldHB + c4);
    if (p.CBn) prefetch_l1(p.CBn + prow * LH + c4);
    if (TAN) {
      prefetch_l1(p.C + trow * LH + c4);
      if (p.XBn) prefetch_l1(p.XBn + trow * p.ldXB + p.hoff + c4);
      if (p.CBn) prefetch_l1(p.CBn + trow * LH + c4);
    }
  }
  // ---- phase A (warp = gate): recompute the gate activations (and tangents); n, r, xd, nd stay in registers
  V16 n, xd, nd;
  float r;
  {
    V16 act, actd;
    gate_fwd<TAN>(G, q, qd, p.ln, n, r, xd, nd, act, actd);
    st16(sm.f.xa[G], act);
    if (TAN) st16(sm.f.xad[G], actd);
  }
  __syncthreads();
  // ---- phase B (thread = 4 columns): cell-level reverse
  {
    StateFwd s;
    state_fwd<TAN>(sm.f, p.C + prow * LH, p.C + trow * LH, s, sm.f.red, tog);
    const F4 g4 = ld4(p.ln.gamma[4]), b4 = ld4(p.ln.beta[4]);
    F4 hb = ld4_or_zero(p.XBn ? p.XBn + prow * p.ldXB + p.hoff : nullptr);
    if (p.HB) {
      const F4 x = ld4(p.HB + prow * p.ldHB);
#pragma unroll
      for (int e = 0; e < 4; ++e) hb.v[e] += x.v[e];
    }
    const float yb = p.ybar_blk[(int)(prow / p.B) & 7];
    const F4 wd = p.wdec ? ld4(p.wdec) : F4{{0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int e = 0; e < 4; ++e) hb.v[e] = fmaf(yb, wd.v[e], hb.v[e]);
    const F4 cnb = ld4_or_zero(p.CBn ? p.CBn + prow * LH : nullptr);
    F4 hdb = F4{{0.f, 0.f, 0.f, 0.f}}, cndb = F4{{0.f, 0.f, 0.f, 0.f}};
    if (TAN) {
      hdb = ld4_or_zero(p.XBn ? p.XBn + trow * p.ldXB + p.hoff : nullptr);
#pragma unroll
      for (int e = 0; e < 4; ++e) hdb.v[e] = fmaf(p.ydot_bar, wd.v[e], hdb.v[e]);
      cndb = ld4_or_zero(p.CBn ? p.CBn + trow * LH : nullptr);
    }
    F4 ncb, ncdb, ybo, ydbo;
    float rs[5] = {0.f, 0.f, 0.f, 0.f, 0.f};   // sum(ncb), sum(nc ncb), sum(ncdb), pp = sum(nc ncdb), ss = sum(ncdb ncd)
    {
      F4 dgs, dbs, dwv;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float cn = fmaf(s.nc.v[e], g4.v[e], b4.v[e]);
        const float tc = tanhf_(cn);
        const float dt = 1.f - tc * tc;
        const float so = s.so.v[e];
        float tcb = hb.v[e] * so;
        float sob = hb.v[e] * tc;
        float cnbar = cnb.v[e];
        float cndbar = 0.f;
        const float h = tc * so;
        float hd = 0.f;
        ydbo.v[e] = 0.f;
        ncdb.v[e] = 0.f;
        if (TAN) {
          const float cnd = s.ncd.v[e] * g4.v[e];
          const float tcd = dt * cnd;
          hd = tcd * so + tc * s.sod.v[e];
          tcb += hdb.v[e] * s.sod.v[e];
          sob += hdb.v[e] * tcd;
          const float tcdb = hdb.v[e] * so;
          ydbo.v[e] = hdb.v[e] * tc;
          cnbar += tcdb * (-2.f * tc * dt) * cnd;
          cndbar = cndb.v[e] + tcdb * dt;
        }
        cnbar += tcb * dt;
        ybo.v[e] = sob;
        ncb.v[e] = cnbar * g4.v[e];
        dgs.v[e] = cnbar * s.nc.v[e];
        dbs.v[e] = cnbar;
        dwv.v[e] = h * yb;
        rs[0] += ncb.v[e];
        rs[1] = fmaf(s.nc.v[e], ncb.v[e], rs[1]);
        if (TAN) {
          ncdb.v[e] = cndbar * g4.v[e];
          dgs.v[e] += cndbar * s.ncd.v[e];
          dwv.v[e] += hd * p.ydot_bar;
          rs[2] += ncdb.v[e];
          rs[3] = fmaf(s.nc.v[e], ncdb.v[e], rs[3]);
          rs[4] = fmaf(ncdb.v[e], s.ncd.v[e], rs[4]);
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) { acc.sg.v[e] += dgs.v[e]; acc.sb.v[e] += dbs.v[e]; acc.dw.v[e] += dwv.v[e]; }
      acc.dbd += yb;
    }
    block_sum(rs, sm.f.red, tog);
    // LN(state) reverse -> adjoint of c' (and of its tangent); then c' = c*sf + si*tj
    F4 cbar, cdbar, ybi, ybj, ybf, ydbi, ydbj, ydbf;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float cpb = s.rc * (ncb.v[e] - rs[0] * INV_LH - s.nc.v[e] * rs[1] * INV_LH);
      float cpdb = 0.f;
      if (TAN) {
        cpdb = s.rc * (ncdb.v[e] - rs[2] * INV_LH - s.nc.v[e] * rs[3] * INV_LH);
        cpb -= s.rc * INV_LH * (s.nc.v[e] * rs[4] + s.qq * cpdb + rs[3] * s.ncd.v[e]);
      }
      cbar.v[e] = cpb * s.sf.v[e];
      ybf.v[e] = cpb * s.c.v[e];
      ybi.v[e] = cpb * s.tj.v[e];
      ybj.v[e] = cpb * s.si.v[e];
      if (TAN) {
        cbar.v[e] += cpdb * s.sfd.v[e];
        ybf.v[e] += cpdb * s.cd.v[e];
        ybi.v[e] += cpdb * s.tjd.v[e];
        ybj.v[e] += cpdb * s.sid.v[e];
        cdbar.v[e] = cpdb * s.sf.v[e];
        ydbf.v[e] = cpdb * s.c.v[e];
        ydbi.v[e] = cpdb * s.tj.v[e];
        ydbj.v[e] = cpdb * s.si.v[e];
      }
    }
    st4(p.CB + prow * LH, cbar);
    st4(sm.yb[0], ybi); st4(sm.yb[1], ybj); st4(sm.yb[2], ybf); st4(sm.yb[3], ybo);
    if (TAN) {
      st4(p.CB + trow * LH, cdbar);
      st4(sm.ydb[0], ydbi); st4(sm.ydb[1], ydbj); st4(sm.ydb[2], ydbf); st4(sm.ydb[3], ydbo);
    }
  }
  __syncthreads();
  // ---- phase C (warp = gate): reverse through the nonlinearity and its LayerNorm
  gate_rev<TAN>(G, p, sm, prow, trow, n, r, xd, nd, acc);
  // No barrier needed here: the next row's phase A writes xa[G] / xad[G], which only warp G reads in phase C,
  // and every other buffer is rewritten only after the next row's post-phase-A barrier.
}

__global__ void __launch_bounds__(LS_THREADS, 3) lstm_rev_kernel(const LstmRevParams p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t lr_smem_raw[];
  LstmSmemRev& sm = *reinterpret_cast<LstmSmemRev*>(lr_smem_raw);
  RevAcc acc;
#pragma unroll
  for (int j = 0; j < 16; ++j) acc.dg[j] = acc.db[j] = 0.f;
#pragma unroll
  for (int e = 0; e < 4; ++e) acc.sg.v[e] = acc.sb.v[e] = acc.dw.v[e] = 0.f;
  acc.dbd = 0.f;
  const int nrows = p.n_plain + p.n_tan;
  int tog = 0;
  // tangent rows first: they are the longest
  for (int i = blockIdx.x; i < nrows; i += gridDim.x) {
    if (i < p.n_tan) lstm_rev_row<true>(p, sm, p.tan_prow0 + i, p.trow0 + i, acc, tog);
    else lstm_rev_row<false>(p, sm, p.prow0 + (i - p.n_tan), 0, acc, tog);
  }
  if (p.partials == nullptr) return;
  // flush this CTA's totals to its slice: vectors 0-3 / 5-8 by gate warps, 4 / 9 / 10 by column threads
  float* dst = p.partials + (long long)blockIdx.x * LR_NPART;
  const int G = threadIdx.x >> 5;
  if (!p.init_partials) {
    V16 o;
    ld16(dst + G * LH, o);
#pragma unroll
    for (int j = 0; j < 16; ++j) acc.dg[j] += o[j];
    ld16(dst + (5 + G) * LH, o);
#pragma unroll
    for (int j = 0; j < 16; ++j) acc.db[j] += o[j];
    const F4 a = ld4(dst + 4 * LH), b = ld4(dst + 9 * LH), c = ld4(dst + 10 * LH);
#pragma unroll
    for (int e = 0; e < 4; ++e) { acc.sg.v[e] += a.v[e]; acc.sb.v[e] += b.v[e]; acc.dw.v[e] += c.v[e]; }
    if (threadIdx.x == 0) acc.dbd += dst[LR_NVEC * LH];
  }
  st16(dst + G * LH, acc.dg);
  st16(dst + (5 + G) * LH, acc.db);
  st4(dst + 4 * LH, acc.sg);
  st4(dst + 9 * LH, acc.sb);
  st4(dst + 10 * LH, acc.dw);
  if (threadIdx.x == 0) dst[LR_NVEC * LH] = acc.dbd;
}

// Sums the per-CTA partial parameter gradients of one reverse pass into the gradient bucket.
struct LnGradParams {
  const float* partials; int nslices;
  float* dgamma[5]; float* dbeta[5]; float* dwdec; float* dbdec;
};
constexpr int LNG_SPLIT = 16;
__global__ void __launch_bounds__(256) lngrad_reduce_kernel(const LnGradParams p) {
  pdl_trigger();
  pdl_wait();
  const int k = blockIdx.x * 256 + threadIdx.x;
  if (k > LR_NVEC * LH) return;
  const int per = (p.nslices + LNG_SPLIT - 1) / LNG_SPLIT;
  const int b0 = blockIdx.y * per, b1 = min(p.nslices, b0 + per);
  if (b0 >= b1) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int b = b0;
  for (; b + 4 <= b1; b += 4) {
    s0 += p.partials[(long long)b * LR_NPART + k];
    s1 += p.partials[(long long)(b + 1) * LR_NPART + k];
    s2 += p.partials[(long long)(b + 2) * LR_NPART + k];
    s3 += p.partials[(long long)(b + 3) * LR_NPART + k];
  }
  for (; b < b1; ++b) s0 += p.partials[(long long)b * LR_NPART + k];
  const float s = (s0 + s1) + (s2 + s3);
  if (k == LR_NVEC * LH) {
    if (p.dbdec) atomicAdd(p.dbdec, s);
    return;
  }
  const int vec = k / LH, col = k % LH;
  float* dst = vec < 5 ? p.dgamma[vec] : (vec < 10 ? p.dbeta[vec - 5] : p.dwdec);
  if (dst) atomicAdd(dst + col, s);
}

int lstm_fwd(const LstmFwdParams& p, cudaStream_t stream) {
  if (p.nrows <= 0) return 0;
  const int grid = p.nrows < 148 * 8 ? p.nrows : 148 * 8;
  SGG_LAUNCH(lstm_fwd_kernel, grid, LS_THREADS, 0, stream, p);
  return 0;
}
int lstm_tan(const LstmTanParams& p, cudaStream_t stream) {
  if (p.nrows <= 0) return 0;
  const int grid = p.nrows < 148 * 8 ? p.nrows : 148 * 8;
  SGG_LAUNCH(lstm_tan_kernel, grid, LS_THREADS, 0, stream, p);
  return 0;
}
int lstm_rev_grid(int nrows) { return nrows < LR_MAX_GRID ? nrows : LR_MAX_GRID; }
long long lstm_rev_partials_floats() { return (long long)LR_MAX_GRID * LR_NPART; }
int lstm_rev(const LstmRevParams& p, cudaStream_t stream) {
  const int nrows = p.n_plain + p.n_tan;
  if (nrows <= 0) return 0;
  static bool configured = false;
  if (!configured) {
    SGG_CUDA(cudaFuncSetAttribute(lstm_rev_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LstmSmemRev)));
    configured = true;
  }
  SGG_LAUNCH(lstm_rev_kernel, lstm_rev_grid(nrows), LS_THREADS, sizeof(LstmSmemRev), stream, p);
  return 0;
}
int lngrad_reduce(const LnGradParams& p, cudaStream_t stream) {
  SGG_LAUNCH(lngrad_reduce_kernel, dim3((LR_NVEC * LH + 1 + 255) / 256, LNG_SPLIT), 256, 0, stream, p);
  return 0;
}

}  // namespace sgg

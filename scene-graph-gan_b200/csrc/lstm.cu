// LayerNormBasicLSTMCell(512) pointwise stages (gen:79,87 / disc:81,89), one warp per row:
//   lstm_fwd : i,j,f,o = LN(split(q)); c' = c*sigmoid(f+1) + sigmoid(i)*tanh(j); c_new = LN(c');
//              h = tanh(c_new)*sigmoid(o); optional D head y = h.w_dec + b (disc:90)
//   lstm_tan : forward tangent (JVP) of the same, for the WGAN-GP interpolate stream
//   lstm_rev : reverse pass, first order or reverse-over-(primal+tangent); also LN gamma/beta
//              gradients and the D head's w_dec / b_dec gradients.
// The gate pre-activations q = [z,u,h] K come from the tcgen05 GEMM; all statistics stay fp32
// (LN eps = 1e-12).  Everything is recomputed from (q, c_in) in the reverse pass.
#include "common.cuh"
#include "../../include/sgg_b200.h"

namespace sgg {

constexpr int LH = 512;             // LSTM units
constexpr int LS_WARPS = 4;
constexpr int LS_THREADS = LS_WARPS * 32;
constexpr float LN_EPS_F = 1e-12f;
constexpr float FORGET_BIAS_F = 1.0f;

typedef float V16[16];

__device__ __forceinline__ void ld16(const float* base, V16& v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = *reinterpret_cast<const float4*>(base + lane * 4 + 128 * i);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
__device__ __forceinline__ void ld16_or_zero(const float* base, V16& v) {
  if (base) ld16(base, v);
  else {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = 0.f;
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
__device__ __forceinline__ float sum16(const V16& a) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) s += a[j];
  return warp_sum(s);
}
__device__ __forceinline__ float dot16(const V16& a, const V16& b) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) s = fmaf(a[j], b[j], s);
  return warp_sum(s);
}
// n = (x - mean) * r, r = rsqrt(var + eps)   (tf.contrib.layers.layer_norm, biased variance)
__device__ __forceinline__ void ln_norm(const V16& x, V16& n, float& r) {
  const float mean = sum16(x) * (1.0f / LH);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    n[j] = x[j] - mean;
    s = fmaf(n[j], n[j], s);
  }
  const float var = warp_sum(s) * (1.0f / LH);
  r = 1.0f / sqrtf(var + LN_EPS_F);
#pragma unroll
  for (int j = 0; j < 16; ++j) n[j] *= r;
}
// out = r * (w - mean(w) - n * mean(n*w)) : LN JVP and VJP (the Jacobian is symmetric)
__device__ __forceinline__ void ln_proj(const V16& n, float r, const V16& w, V16& out) {
  const float mw = sum16(w) * (1.0f / LH);
  const float mnw = dot16(n, w) * (1.0f / LH);
#pragma unroll
  for (int j = 0; j < 16; ++j) out[j] = r * (w[j] - mw - n[j] * mnw);
}

struct LstmLN {
  const float* gamma[5];  // input, transform, forget, output, state
  const float* beta[5];
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
};

__global__ void __launch_bounds__(LS_THREADS) lstm_fwd_kernel(const LstmFwdParams p) {
  const int row = blockIdx.x * LS_WARPS + (threadIdx.x >> 5);
  if (row >= p.nrows) return;
  const float* q = p.Q + (long long)row * p.ldQ;
  V16 x, n, g, act[4];
  float r;
#pragma unroll
  for (int G = 0; G < 4; ++G) {
    ld16(q + G * LH, x);
    ln_norm(x, n, r);
    ld16(p.ln.gamma[G], g);
    ld16(p.ln.beta[G], x);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float a = fmaf(n[j], g[j], x[j]);
      act[G][j] = (G == 1) ? tanhf_(a) : sigmoidf_(G == 2 ? a + FORGET_BIAS_F : a);
    }
  }
  ld16(p.Cin + (long long)row * LH, x);
#pragma unroll
  for (int j = 0; j < 16; ++j) x[j] = fmaf(x[j], act[2][j], act[0][j] * act[1][j]);
  ln_norm(x, n, r);
  ld16(p.ln.gamma[4], g);
  ld16(p.ln.beta[4], x);
  V16 cn, h;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    cn[j] = fmaf(n[j], g[j], x[j]);
    h[j] = tanhf_(cn[j]) * act[3][j];
  }
  st16(p.Cout + (long long)row * LH, cn);
  if (p.CH) st16_hl(p.CH + (long long)row * p.ldCH, p.ch_lo, cn);
  if (p.Xn) st16_hl(p.Xn + (long long)row * p.ldX + p.hoff, p.x_lo, h);
  if (p.Y) {
    ld16(p.wdec, g);
    const float y = dot16(h, g) + p.bdec[0];
    if ((threadIdx.x & 31) == 0) p.Y[(long long)row * p.ldY] = y;
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
};

__global__ void __launch_bounds__(LS_THREADS) lstm_tan_kernel(const LstmTanParams p) {
  const int i = blockIdx.x * LS_WARPS + (threadIdx.x >> 5);
  if (i >= p.nrows) return;
  const long long prow = p.prow0 + i, trow = p.trow0 + i;
  const float* q = p.Q + prow * p.ldQ;
  const float* qd = p.Q + trow * p.ldQ;
  V16 x, n, g, act[4], actd[4];
  float r;
#pragma unroll
  for (int G = 0; G < 4; ++G) {
    ld16(q + G * LH, x);
    ln_norm(x, n, r);
    ld16(p.ln.gamma[G], g);
    V16 bt, xd, nd;
    ld16(p.ln.beta[G], bt);
    ld16(qd + G * LH, xd);
    ln_proj(n, r, xd, nd);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float a = fmaf(n[j], g[j], bt[j]);
      const float ad = nd[j] * g[j];
      if (G == 1) {
        const float t = tanhf_(a);
        act[G][j] = t;
        actd[G][j] = (1.f - t * t) * ad;
      } else {
        const float s = sigmoidf_(G == 2 ? a + FORGET_BIAS_F : a);
        act[G][j] = s;
        actd[G][j] = s * (1.f - s) * ad;
      }
    }
  }
  V16 c, cd, cp, cpd;
  ld16(p.C + prow * LH, c);
  ld16(p.C + trow * LH, cd);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    cp[j] = fmaf(c[j], act[2][j], act[0][j] * act[1][j]);
    cpd[j] = cd[j] * act[2][j] + c[j] * actd[2][j] + actd[0][j] * act[1][j] + act[0][j] * actd[1][j];
  }
  ln_norm(cp, n, r);
  V16 ncd;
  ln_proj(n, r, cpd, ncd);
  ld16(p.ln.gamma[4], g);
  ld16(p.ln.beta[4], x);
  V16 cnd, hd;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float cn = fmaf(n[j], g[j], x[j]);
    const float tc = tanhf_(cn);
    cnd[j] = ncd[j] * g[j];
    hd[j] = (1.f - tc * tc) * cnd[j] * act[3][j] + tc * actd[3][j];
  }
  st16(p.Cout + trow * LH, cnd);
  if (p.CH) st16_hl(p.CH + trow * p.ldCH, p.ch_lo, cnd);
  if (p.Xn) st16_hl(p.Xn + trow * p.ldX + p.hoff, p.x_lo, hd);
}

// ------------------------------------------------------------------------------------ reverse
// One launch handles `n_plain` first-order rows (row = prow0 + i) and `n_tan` rows that carry a tangent
// (primal row tan_prow0 + j, tangent row trow0 + j).  One warp per row, rows strided over the grid.
// Parameter gradients (5 LN gammas/betas, D head) are reduced without shared-memory atomics: every warp
// owns a [11][512] fp32 slab in shared memory, the CTA sums its slabs, and only the CTA totals go to
// global memory as fp32 reductions.
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
  // parameter gradients (nullable => data path only)
  float* dgamma[5]; float* dbeta[5];
  float* dwdec; float* dbdec;
};

constexpr int LR_WARPS = 8;
constexpr int LR_THREADS = LR_WARPS * 32;
constexpr int LR_SLAB = 11 * LH;   // dgamma[5], dbeta[5], dwdec

// slab accumulate: lane owns columns lane*4 + 128*i (conflict-free float4 accesses)
__device__ __forceinline__ void slab_acc(float* slab_vec, const V16& v, bool first) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float4* p = reinterpret_cast<float4*>(slab_vec + lane * 4 + 128 * i);
    float4 t = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    if (!first) {
      const float4 o = *p;
      t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
    }
    *p = t;
  }
}

template <bool TAN>
__device__ __forceinline__ void lstm_rev_row(const LstmRevParams& p, long long prow, long long trow, float* slab,
                                             bool first, float& dbd_acc) {
  const bool wgrad = slab != nullptr;
  const float* q = p.Q + prow * p.ldQ;
  const float* qd = p.Q + trow * p.ldQ;
  V16 x, n, g, bt, act[4], actd[4], ad[4];
  float r;
  // ---- phase A: recompute forward (and tangent) gate activations
#pragma unroll
  for (int G = 0; G < 4; ++G) {
    ld16(q + G * LH, x);
    ln_norm(x, n, r);
    ld16(p.ln.gamma[G], g);
    ld16(p.ln.beta[G], bt);
    V16 nd;
    if (TAN) {
      V16 xd;
      ld16(qd + G * LH, xd);
      ln_proj(n, r, xd, nd);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float a = fmaf(n[j], g[j], bt[j]);
      if (TAN) ad[G][j] = nd[j] * g[j];
      if (G == 1) {
        const float t = tanhf_(a);
        act[G][j] = t;
        if (TAN) actd[G][j] = (1.f - t * t) * ad[G][j];
      } else {
        const float s = sigmoidf_(G == 2 ? a + FORGET_BIAS_F : a);
        act[G][j] = s;
        if (TAN) actd[G][j] = s * (1.f - s) * ad[G][j];
      }
    }
  }
  V16 c, cd, cp, cpd, nc, ncd;
  float rc;
  ld16(p.C + prow * LH, c);
  if (TAN) ld16(p.C + trow * LH, cd);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    cp[j] = fmaf(c[j], act[2][j], act[0][j] * act[1][j]);
    if (TAN) cpd[j] = cd[j] * act[2][j] + c[j] * actd[2][j] + actd[0][j] * act[1][j] + act[0][j] * actd[1][j];
  }
  ln_norm(cp, nc, rc);
  if (TAN) ln_proj(nc, rc, cpd, ncd);
  ld16(p.ln.gamma[4], g);
  ld16(p.ln.beta[4], bt);
  // ---- phase B: cell-level reverse
  V16 hb, cnb, hdb, cndb;
  ld16_or_zero(p.XBn ? p.XBn + prow * p.ldXB + p.hoff : nullptr, hb);
  if (p.HB) {
    ld16(p.HB + prow * p.ldHB, x);
#pragma unroll
    for (int j = 0; j < 16; ++j) hb[j] += x[j];
  }
  const float yb = p.ybar_blk[(int)(prow / p.B) & 7];
  V16 wd;
  if (p.wdec) {
    ld16(p.wdec, wd);
#pragma unroll
    for (int j = 0; j < 16; ++j) hb[j] = fmaf(yb, wd[j], hb[j]);
  }
  ld16_or_zero(p.CBn ? p.CBn + prow * LH : nullptr, cnb);
  if (TAN) {
    ld16_or_zero(p.XBn ? p.XBn + trow * p.ldXB + p.hoff : nullptr, hdb);
    if (p.wdec) {
#pragma unroll
      for (int j = 0; j < 16; ++j) hdb[j] = fmaf(p.ydot_bar, wd[j], hdb[j]);
    }
    ld16_or_zero(p.CBn ? p.CBn + trow * LH : nullptr, cndb);
  }
  V16 ncb, ncdb;        // adjoints of nc (normalised state) and its tangent
  V16 yb_[4], ydb_[4];  // adjoints of the gate activations (and of their tangents)
  {
    V16 dgs, dbs, dwv;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float cn = fmaf(nc[j], g[j], bt[j]);
      const float tc = tanhf_(cn);
      const float dt = 1.f - tc * tc;
      const float so = act[3][j];
      float tcb = hb[j] * so;
      float sob = hb[j] * tc;
      float cnbar = cnb[j];
      float tcdb = 0.f, cndbar = 0.f;
      float h = tc * so;
      float hd = 0.f;
      if (TAN) {
        const float cnd = ncd[j] * g[j];
        const float tcd = dt * cnd;
        hd = tcd * so + tc * actd[3][j];
        tcb += hdb[j] * actd[3][j];
        sob += hdb[j] * tcd;
        tcdb = hdb[j] * so;
        ydb_[3][j] = hdb[j] * tc;
        cnbar += tcdb * (-2.f * tc * dt) * cnd;
        cndbar = cndb[j] + tcdb * dt;
      }
      cnbar += tcb * dt;
      yb_[3][j] = sob;
      ncb[j] = cnbar * g[j];
      dgs[j] = cnbar * nc[j];
      dbs[j] = cnbar;
      dwv[j] = h * yb;
      if (TAN) {
        ncdb[j] = cndbar * g[j];
        dgs[j] += cndbar * ncd[j];
        dwv[j] += hd * p.ydot_bar;
      }
    }
    if (wgrad) {
      slab_acc(slab + 4 * LH, dgs, first);
      slab_acc(slab + 9 * LH, dbs, first);
      slab_acc(slab + 10 * LH, dwv, first);
      dbd_acc += yb;
    }
  }
  // LN(state) reverse -> adjoint of c' (and of its tangent)
  V16 cpb, cpdb;
  ln_proj(nc, rc, ncb, cpb);
  if (TAN) {
    ln_proj(nc, rc, ncdb, cpdb);
    const float pp = dot16(nc, ncdb), qq = dot16(nc, cpd), ss = dot16(ncdb, ncd);
    const float k = rc * (1.0f / LH);
#pragma unroll
    for (int j = 0; j < 16; ++j) cpb[j] -= k * (nc[j] * ss + qq * cpdb[j] + pp * ncd[j]);
  }
  // c' = c*sf + si*tj
  V16 cbar, cdbar;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    cbar[j] = cpb[j] * act[2][j];
    yb_[2][j] = cpb[j] * c[j];
    yb_[0][j] = cpb[j] * act[1][j];
    yb_[1][j] = cpb[j] * act[0][j];
    if (TAN) {
      cbar[j] += cpdb[j] * actd[2][j];
      yb_[2][j] += cpdb[j] * cd[j];
      yb_[0][j] += cpdb[j] * actd[1][j];
      yb_[1][j] += cpdb[j] * actd[0][j];
      cdbar[j] = cpdb[j] * act[2][j];
      ydb_[2][j] = cpdb[j] * c[j];
      ydb_[0][j] = cpdb[j] * act[1][j];
      ydb_[1][j] = cpdb[j] * act[0][j];
    }
  }
  st16(p.CB + prow * LH, cbar);
  if (TAN) st16(p.CB + trow * LH, cdbar);
  // ---- phase C: per-gate reverse through the nonlinearity and its LayerNorm
#pragma unroll
  for (int G = 0; G < 4; ++G) {
    ld16(q + G * LH, x);
    ln_norm(x, n, r);
    ld16(p.ln.gamma[G], g);
    V16 nb, xb, xd, nd, ndb, xdb;
    if (TAN) {
      ld16(qd + G * LH, xd);
      ln_proj(n, r, xd, nd);
    }
    V16 dgs, dbs;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float y = act[G][j];
      const float d1 = (G == 1) ? (1.f - y * y) : y * (1.f - y);
      float abar = yb_[G][j] * d1;
      if (TAN) {
        const float d2 = (G == 1) ? (-2.f * y * d1) : d1 * (1.f - 2.f * y);
        abar += ydb_[G][j] * d2 * ad[G][j];
        const float adb = ydb_[G][j] * d1;
        ndb[j] = adb * g[j];
        dgs[j] = abar * n[j] + adb * nd[j];
      } else {
        dgs[j] = abar * n[j];
      }
      dbs[j] = abar;
      nb[j] = abar * g[j];
    }
    if (wgrad) {
      slab_acc(slab + G * LH, dgs, first);
      slab_acc(slab + (5 + G) * LH, dbs, first);
    }
    ln_proj(n, r, nb, xb);
    if (TAN) {
      ln_proj(n, r, ndb, xdb);
      const float pp = dot16(n, ndb), qq = dot16(n, xd), ss = dot16(ndb, nd);
      const float k = r * (1.0f / LH);
#pragma unroll
      for (int j = 0; j < 16; ++j) xb[j] -= k * (n[j] * ss + qq * xdb[j] + pp * nd[j]);
      st16_hl(p.QB + trow * p.ldQB + G * LH, p.qb_lo, xdb);
    }
    st16_hl(p.QB + prow * p.ldQB + G * LH, p.qb_lo, xb);
  }
}

__global__ void __launch_bounds__(LR_THREADS, 1) lstm_rev_kernel(const LstmRevParams p) {
  extern __shared__ __align__(16) float lr_smem[];   // [LR_WARPS][11][512] + [LR_WARPS] (only with parameter gradients)
  const bool wgrad = p.dgamma[0] != nullptr;
  const int warp = threadIdx.x >> 5;
  float* slab = wgrad ? lr_smem + warp * LR_SLAB : nullptr;
  const int nrows = p.n_plain + p.n_tan;
  const int gw = blockIdx.x * LR_WARPS + warp;
  const int stride = gridDim.x * LR_WARPS;
  bool first = true;
  float dbd = 0.f;
  // tangent rows first: they are the longest, so they start earliest
  for (int i = gw; i < nrows; i += stride) {
    if (i < p.n_tan) lstm_rev_row<true>(p, p.tan_prow0 + i, p.trow0 + i, slab, first, dbd);
    else lstm_rev_row<false>(p, p.prow0 + (i - p.n_tan), 0, slab, first, dbd);
    first = false;
  }
  if (!wgrad) return;
  float* dbds = lr_smem + LR_WARPS * LR_SLAB;
  if ((threadIdx.x & 31) == 0) dbds[warp] = dbd;
  __syncthreads();
  int nact = min(LR_WARPS, nrows - blockIdx.x * LR_WARPS);   // warps of this CTA that processed >= 1 row
  if (nact <= 0) return;
  for (int k = threadIdx.x; k < LR_SLAB; k += LR_THREADS) {
    float s = 0.f;
    for (int w = 0; w < nact; ++w) s += lr_smem[w * LR_SLAB + k];
    const int vec = k / LH, col = k % LH;
    float* dst = vec < 5 ? p.dgamma[vec] : (vec < 10 ? p.dbeta[vec - 5] : p.dwdec);
    if (dst) atomicAdd(dst + col, s);
  }
  if (p.dbdec && threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < nact; ++w) s += dbds[w];
    atomicAdd(p.dbdec, s);
  }
}

int lstm_fwd(const LstmFwdParams& p, cudaStream_t stream) {
  if (p.nrows <= 0) return 0;
  lstm_fwd_kernel<<<(p.nrows + LS_WARPS - 1) / LS_WARPS, LS_THREADS, 0, stream>>>(p);
  SGG_LAUNCHED();
  return 0;
}
int lstm_tan(const LstmTanParams& p, cudaStream_t stream) {
  if (p.nrows <= 0) return 0;
  lstm_tan_kernel<<<(p.nrows + LS_WARPS - 1) / LS_WARPS, LS_THREADS, 0, stream>>>(p);
  SGG_LAUNCHED();
  return 0;
}
int lstm_rev(const LstmRevParams& p, cudaStream_t stream) {
  const int nrows = p.n_plain + p.n_tan;
  if (nrows <= 0) return 0;
  const bool wgrad = p.dgamma[0] != nullptr;
  const size_t smem = wgrad ? (size_t)(LR_WARPS * LR_SLAB + LR_WARPS) * sizeof(float) : 0;
  static bool configured = false;
  if (!configured) {
    SGG_CUDA(cudaFuncSetAttribute(lstm_rev_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)((LR_WARPS * LR_SLAB + LR_WARPS) * sizeof(float))));
    configured = true;
  }
  int grid = (nrows + LR_WARPS - 1) / LR_WARPS;
  if (grid > 148) grid = 148;
  lstm_rev_kernel<<<grid, LR_THREADS, smem, stream>>>(p);
  SGG_LAUNCHED();
  return 0;
}

}  // namespace sgg

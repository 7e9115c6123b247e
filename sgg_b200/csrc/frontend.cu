// HBM-bound half of the convolutional front-end (gen:29-68): tf.contrib.layers.layer_norm(activation_fn=tf.nn.elu) on an
// NHWC tensor -- moments over all of (H, W, C) per sample, epsilon 1e-12, per-channel gamma / beta (begin_params_axis = -1),
// then ELU (gen:30,32,36,...) -- forward and reverse, as two passes each over the sample:
//
//   forward   pass 1  per-chunk (count, mean, M2)                                   reads x
//             pass 2  Chan combination of the sample's chunks; y = elu(n*gamma+beta)  reads x, writes y      (3 N floats)
//   reverse   pass 1  per-chunk sums of dn and dn*n, per-channel dgamma / dbeta      reads x, dy
//             pass 2  dx = rstd (dn - mean(dn) - n mean(dn n))                       reads x, dy, writes dx   (5 N floats)
//
// with n = (x - mean) rstd, z = n gamma + beta, dz = dy * (z > 0 ? 1 : exp(z)), dn = dz gamma.  The pre-activation is
// recomputed from x, so x is the only saved tensor (a framework's group_norm + elu pair keeps two and moves 5 + 8 N floats).
// The convolutions around it stay library calls (sgg_b200/frontend.py).
#include "common.cuh"
#include "../../include/sgg_b200.h"

namespace sgg {

constexpr int LE_THREADS = 256;
constexpr int LE_STEP = LE_THREADS * 4;       // elements per CTA pass; a multiple of every channel count (C | 1024)

// Elements per CTA: a multiple of LE_STEP chosen so that the grid is about 8 CTAs per SM (a CTA's prologue -- the
// combination of its sample's chunk partials -- is then small against its streaming time), between 4 and 64 passes.
static inline long long le_chunk_elems(long long B, long long n) {
  long long c = (B * n + 148 * 8 - 1) / (148 * 8);
  c = (c + LE_STEP - 1) / LE_STEP * LE_STEP;
  if (c < 4 * LE_STEP) c = 4 * LE_STEP;
  if (c > 64 * LE_STEP) c = 64 * LE_STEP;
  return c;
}
static inline long long le_chunks(long long B, long long n) { const long long c = le_chunk_elems(B, n); return (n + c - 1) / c; }

struct LnEluParams {
  const float* x; const float* dy; const float* gamma; const float* beta;
  float* y; float* dx;
  float* stats;              // [B, 2] mean, rstd
  float* part;               // forward: [B, chunks, 3] (count, mean, M2); reverse: [B, chunks, 2] (sum dn, sum dn n)
  float* cpart;              // reverse: [B * chunks, 2, C] per-CTA dgamma / dbeta
  float* dgamma; float* dbeta;
  long long N;               // H * W * C elements per sample
  long long chunk;           // elements per CTA (multiple of LE_STEP)
  int C, chunks;
  float eps;
};

__device__ __forceinline__ float block_sum_256(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LE_THREADS / 32; ++i) s += red[i];
  return s;
}

// expm1(z) for z <= 0 without libdevice's branches (the apply kernel issued 67 % of its cycles with expm1f): the degree-6
// Taylor polynomial down to -0.25 (truncation 5e-8 relative), exp(z) - 1 from one MUFU.EX2 below (no cancellation there)
__device__ __forceinline__ float expm1_neg(float z) {
  const float p = z * (1.f + z * (0.5f + z * (1.f / 6.f + z * (1.f / 24.f + z * (1.f / 120.f + z * (1.f / 720.f))))));
  const float e = fast_exp(z) - 1.f;
  return z > -0.25f ? p : e;
}

// ------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(LE_THREADS) ln_elu_stats_kernel(const LnEluParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[LE_THREADS / 32];
  const int b = blockIdx.y, ch = blockIdx.x;
  const long long e0 = (long long)ch * p.chunk, e1 = min(p.N, e0 + p.chunk);
  const float* xs = p.x + (long long)b * p.N;
  const float shift = xs[e0];                 // sums about a value of the chunk: no cancellation in M2
  float s = 0.f, q = 0.f;
#pragma unroll 4
  for (long long e = e0 + threadIdx.x * 4; e < e1; e += LE_STEP) {
    const float4 v = *reinterpret_cast<const float4*>(xs + e);
    const float a0 = v.x - shift, a1 = v.y - shift, a2 = v.z - shift, a3 = v.w - shift;
    s += (a0 + a1) + (a2 + a3);
    q = fmaf(a0, a0, q); q = fmaf(a1, a1, q); q = fmaf(a2, a2, q); q = fmaf(a3, a3, q);
  }
  s = block_sum_256(s, red);
  q = block_sum_256(q, red);
  if (threadIdx.x == 0) {
    const float n = (float)(e1 - e0);
    float* o = p.part + ((long long)b * p.chunks + ch) * 3;
    o[0] = n; o[1] = shift + s / n; o[2] = fmaxf(q - s * s / n, 0.f);
  }
}

// (mean, rstd) of sample b from its chunk partials: n = sum n_i, mean = sum n_i m_i / n, M2 = sum (M2_i + n_i (m_i - mean)^2)
// (Chan et al., all chunks at once), the chunks spread over the threads
__device__ __forceinline__ void ln_elu_combine(const LnEluParams& p, int b, float* red, float& mean, float& rstd) {
  const float* q = p.part + (long long)b * p.chunks * 3;
  float n = 0.f, nm = 0.f;
  for (int i = threadIdx.x; i < p.chunks; i += LE_THREADS) { n += q[3 * i]; nm = fmaf(q[3 * i], q[3 * i + 1], nm); }
  n = block_sum_256(n, red);
  nm = block_sum_256(nm, red);
  mean = nm / n;
  float m2 = 0.f;
  for (int i = threadIdx.x; i < p.chunks; i += LE_THREADS) {
    const float d = q[3 * i + 1] - mean;
    m2 += q[3 * i + 2] + q[3 * i] * d * d;
  }
  m2 = block_sum_256(m2, red);
  rstd = rsqrtf(m2 / n + p.eps);
}

__global__ void __launch_bounds__(LE_THREADS) ln_elu_apply_kernel(const LnEluParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[LE_THREADS / 32];
  const int b = blockIdx.y, ch = blockIdx.x;
  float mean, rstd;
  ln_elu_combine(p, b, red, mean, rstd);
  if (ch == 0 && threadIdx.x == 0) { p.stats[2 * b] = mean; p.stats[2 * b + 1] = rstd; }
  const long long e0 = (long long)ch * p.chunk, e1 = min(p.N, e0 + p.chunk);
  const float* xs = p.x + (long long)b * p.N;
  float* ys = p.y + (long long)b * p.N;
  const int c0 = (threadIdx.x * 4) % p.C;     // LE_STEP and the chunk size are multiples of C: the channel group never changes
  const float4 g = *reinterpret_cast<const float4*>(p.gamma + c0), be = *reinterpret_cast<const float4*>(p.beta + c0);
#pragma unroll 4
  for (long long e = e0 + threadIdx.x * 4; e < e1; e += LE_STEP) {
    const float4 v = *reinterpret_cast<const float4*>(xs + e);
    float z[4] = {fmaf((v.x - mean) * rstd, g.x, be.x), fmaf((v.y - mean) * rstd, g.y, be.y),
                  fmaf((v.z - mean) * rstd, g.z, be.z), fmaf((v.w - mean) * rstd, g.w, be.w)};
#pragma unroll
    for (int j = 0; j < 4; ++j) z[j] = z[j] > 0.f ? z[j] : expm1_neg(z[j]);
    *reinterpret_cast<float4*>(ys + e) = make_float4(z[0], z[1], z[2], z[3]);
  }
}

// ------------------------------------------------------------------------------------ reverse
__global__ void __launch_bounds__(LE_THREADS) ln_elu_bwd_reduce_kernel(const LnEluParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[LE_THREADS / 32];
  __shared__ float cg[512], cb[512];
  const int b = blockIdx.y, ch = blockIdx.x;
  for (int c = threadIdx.x; c < p.C; c += LE_THREADS) { cg[c] = 0.f; cb[c] = 0.f; }
  const float mean = p.stats[2 * b], rstd = p.stats[2 * b + 1];
  const long long e0 = (long long)ch * p.chunk, e1 = min(p.N, e0 + p.chunk);
  const float* xs = p.x + (long long)b * p.N;
  const float* ds = p.dy + (long long)b * p.N;
  const int c0 = (threadIdx.x * 4) % p.C;
  const float4 g4 = *reinterpret_cast<const float4*>(p.gamma + c0), b4 = *reinterpret_cast<const float4*>(p.beta + c0);
  const float g[4] = {g4.x, g4.y, g4.z, g4.w}, be[4] = {b4.x, b4.y, b4.z, b4.w};
  float s1 = 0.f, s2 = 0.f, dg[4] = {0.f, 0.f, 0.f, 0.f}, db[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (long long e = e0 + threadIdx.x * 4; e < e1; e += LE_STEP) {
    const float4 v = *reinterpret_cast<const float4*>(xs + e);
    const float4 d = *reinterpret_cast<const float4*>(ds + e);
    const float xv[4] = {v.x, v.y, v.z, v.w}, dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float n = (xv[j] - mean) * rstd;
      const float z = fmaf(n, g[j], be[j]);
      const float dz = dv[j] * (z > 0.f ? 1.f : fast_exp(z));
      const float dn = dz * g[j];
      dg[j] = fmaf(dz, n, dg[j]);
      db[j] += dz;
      s1 += dn;
      s2 = fmaf(dn, n, s2);
    }
  }
  __syncthreads();           // cg / cb cleared
#pragma unroll
  for (int j = 0; j < 4; ++j) { atomicAdd(&cg[c0 + j], dg[j]); atomicAdd(&cb[c0 + j], db[j]); }
  s1 = block_sum_256(s1, red);
  s2 = block_sum_256(s2, red);      // (its barriers also order the shared atomics before the read-out below)
  const long long cta = (long long)b * p.chunks + ch;
  if (threadIdx.x == 0) { p.part[2 * cta] = s1; p.part[2 * cta + 1] = s2; }
  float* o = p.cpart + cta * 2 * p.C;
  for (int c = threadIdx.x; c < p.C; c += LE_THREADS) { o[c] = cg[c]; o[p.C + c] = cb[c]; }
}

__global__ void __launch_bounds__(LE_THREADS) ln_elu_bwd_apply_kernel(const LnEluParams p, int nctas) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y, ch = blockIdx.x;
  if (b == gridDim.y - 1 && ch >= p.chunks) {   // the extra CTAs of the last grid row: dgamma / dbeta = sum over the per-CTA partials
    const int k = (ch - p.chunks) * LE_THREADS + threadIdx.x;
    if (k < 2 * p.C) {
      float a0 = 0.f, a1 = 0.f;
      int i = 0;
      for (; i + 2 <= nctas; i += 2) { a0 += p.cpart[(long long)i * 2 * p.C + k]; a1 += p.cpart[(long long)(i + 1) * 2 * p.C + k]; }
      if (i < nctas) a0 += p.cpart[(long long)i * 2 * p.C + k];
      (k < p.C ? p.dgamma[k] : p.dbeta[k - p.C]) = a0 + a1;
    }
    return;
  }
  if (ch >= p.chunks) return;
  __shared__ float red[LE_THREADS / 32];
  float t1 = 0.f, t2 = 0.f;
  for (int i = threadIdx.x; i < p.chunks; i += LE_THREADS) {
    t1 += p.part[2 * ((long long)b * p.chunks + i)];
    t2 += p.part[2 * ((long long)b * p.chunks + i) + 1];
  }
  const float m1 = block_sum_256(t1, red) / (float)p.N, m2 = block_sum_256(t2, red) / (float)p.N;
  const float mean = p.stats[2 * b], rstd = p.stats[2 * b + 1];
  const long long e0 = (long long)ch * p.chunk, e1 = min(p.N, e0 + p.chunk);
  const float* xs = p.x + (long long)b * p.N;
  const float* ds = p.dy + (long long)b * p.N;
  float* os = p.dx + (long long)b * p.N;
  const int c0 = (threadIdx.x * 4) % p.C;
  const float4 g4 = *reinterpret_cast<const float4*>(p.gamma + c0), b4 = *reinterpret_cast<const float4*>(p.beta + c0);
  const float g[4] = {g4.x, g4.y, g4.z, g4.w}, be[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll 4
  for (long long e = e0 + threadIdx.x * 4; e < e1; e += LE_STEP) {
    const float4 v = *reinterpret_cast<const float4*>(xs + e);
    const float4 d = *reinterpret_cast<const float4*>(ds + e);
    const float xv[4] = {v.x, v.y, v.z, v.w}, dv[4] = {d.x, d.y, d.z, d.w};
    float r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float n = (xv[j] - mean) * rstd;
      const float z = fmaf(n, g[j], be[j]);
      const float dn = dv[j] * (z > 0.f ? 1.f : fast_exp(z)) * g[j];
      r[j] = rstd * (dn - m1 - n * m2);
    }
    *reinterpret_cast<float4*>(os + e) = make_float4(r[0], r[1], r[2], r[3]);
  }
}

static int le_check(int64_t B, int64_t HW, int32_t C, const char* who) {
  SGG_CHECK(B >= 1 && HW >= 1 && C >= 4 && C <= 512 && (LE_STEP % C) == 0, "%s: needs 1 <= B, HW and C in {4, 8, ..., 512} dividing 1024 (got B=%lld HW=%lld C=%d)",
            who, (long long)B, (long long)HW, C);
  SGG_CHECK(le_chunks(B, HW * C) <= 65535 && B <= 65535, "%s: sample too large for one launch", who);
  return 0;
}

}  // namespace sgg

using namespace sgg;

extern "C" int64_t sgg_ln_elu_scratch_floats(int64_t B, int64_t HW, int32_t C) {
  if (le_check(B, HW, C, "sgg_ln_elu_scratch_floats") != 0) return -1;
  const long long ctas = B * le_chunks(B, HW * C);
  return ctas * 3 + ctas * 2 * C;
}

extern "C" int sgg_ln_elu_forward(const float* x, const float* gamma, const float* beta, int64_t B, int64_t HW, int32_t C, float eps,
                                  float* y, float* stats, float* scratch, sgg_stream_t stream) {
  SGG_CHECK(x && gamma && beta && y && stats && scratch, "sgg_ln_elu_forward: null argument");
  SGG_TRY(le_check(B, HW, C, "sgg_ln_elu_forward"));
  LnEluParams p{};
  p.x = x; p.gamma = gamma; p.beta = beta; p.y = y; p.stats = stats; p.part = scratch;
  p.N = HW * C; p.C = C; p.chunk = le_chunk_elems(B, p.N); p.chunks = (int)le_chunks(B, p.N); p.eps = eps;
  const dim3 grid(p.chunks, (unsigned)B);
  SGG_LAUNCH(ln_elu_stats_kernel, grid, LE_THREADS, 0, (cudaStream_t)stream, p);
  SGG_LAUNCH(ln_elu_apply_kernel, grid, LE_THREADS, 0, (cudaStream_t)stream, p);
  return 0;
}

extern "C" int sgg_ln_elu_backward(const float* x, const float* dy, const float* gamma, const float* beta, const float* stats,
                                   int64_t B, int64_t HW, int32_t C, float* dx, float* dgamma, float* dbeta, float* scratch,
                                   sgg_stream_t stream) {
  SGG_CHECK(x && dy && gamma && beta && stats && dx && dgamma && dbeta && scratch, "sgg_ln_elu_backward: null argument");
  SGG_TRY(le_check(B, HW, C, "sgg_ln_elu_backward"));
  LnEluParams p{};
  p.x = x; p.dy = dy; p.gamma = gamma; p.beta = beta; p.dx = dx; p.dgamma = dgamma; p.dbeta = dbeta;
  p.stats = const_cast<float*>(stats);
  p.N = HW * C; p.C = C; p.chunk = le_chunk_elems(B, p.N); p.chunks = (int)le_chunks(B, p.N);
  const long long ctas = B * p.chunks;
  p.part = scratch; p.cpart = scratch + ctas * 3;
  SGG_LAUNCH(ln_elu_bwd_reduce_kernel, dim3(p.chunks, (unsigned)B), LE_THREADS, 0, (cudaStream_t)stream, p);
  // apply grid: the last row carries extra CTAs that finish the per-channel gradients
  const int extra = (2 * C + LE_THREADS - 1) / LE_THREADS;
  SGG_LAUNCH(ln_elu_bwd_apply_kernel, dim3(p.chunks + extra, (unsigned)B), LE_THREADS, 0, (cudaStream_t)stream, p, (int)ctas);
  return 0;
}

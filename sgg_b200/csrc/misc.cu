// Small HBM-bound / elementwise kernels of the hot path: mean-pool initial state (gen:76-77),
// embedding gather + interpolate mixing (disc:86-87, tfgan interpolates), hi/lo packing,
// WGAN-GP slopes/penalty (train:245-250 via tfgan), Wasserstein losses (train:252-253),
// TF-form Adam over flat buckets (train:258-259), Philox RNG (gen:81, tfgan alpha).
#include "common.cuh"
#include "../../include/sgg_b200.h"

namespace sgg {
static inline long long llmin(long long a, long long b) { return a < b ? a : b; }

// ------------------------------------------------------------------------------------ mean pool
// c0[b,:] = mean_r a[b,r,:]  -> fp32 C0, hi/lo CH0, and h0 hi/lo into X0's h columns, replicated
// into `nblk` stream row blocks.
struct MeanPoolParams {
  const __nv_bfloat16* a; int B, R, nblk;
  float* C0;                                              // [nblk*B, 512]
  __nv_bfloat16* CH; long long ldCH; long long ch_lo;     // [nblk*B, ...]
  __nv_bfloat16* X; long long ldX; long long x_lo; int hoff;
};

__global__ void __launch_bounds__(256) meanpool_kernel(const MeanPoolParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[4][512];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int cg = tid & 63, rg = tid >> 6;
  const uint4* base = reinterpret_cast<const uint4*>(p.a + (size_t)b * p.R * 512);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 7
  for (int r = rg; r < p.R; r += 4) {
    const uint4 pk = __ldg(base + (size_t)r * 64 + cg);
    const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      acc[2 * e] += bf16lo_to_f32(w[e]);
      acc[2 * e + 1] += bf16hi_to_f32(w[e]);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[rg][cg * 8 + e] = acc[e];
  __syncthreads();
  const float inv = 1.0f / p.R;
  for (int c = tid; c < 512; c += 256) {
    const float m = (red[0][c] + red[1][c] + red[2][c] + red[3][c]) * inv;
    __nv_bfloat16 h, l;
    split_bf16(m, h, l);
    for (int s = 0; s < p.nblk; ++s) {
      const long long row = (long long)s * p.B + b;
      p.C0[row * 512 + c] = m;
      if (p.CH) {
        p.CH[row * p.ldCH + c] = h;
        p.CH[row * p.ldCH + p.ch_lo + c] = l;
      }
      if (p.X) {
        p.X[row * p.ldX + p.hoff + c] = h;
        p.X[row * p.ldX + p.x_lo + p.hoff + c] = l;
      }
    }
  }
}

int meanpool(const MeanPoolParams& p, cudaStream_t stream) {
  SGG_LAUNCH(meanpool_kernel, p.B, 256, 0, stream, p);
  return 0;
}

// ------------------------------------------------------------------------------------ pack hi/lo
// dst[r, c] (hi) / dst[r, lo_off + c] (lo) = scale[r % smod] * (src[r, c] + mix[r % mmod] * src2[r, c])
struct PackParams {
  int rows, cols;
  int rpg;                             // rows per group (row r -> group r / rpg, index r % rpg); 0 = one group
  const float* src; long long ld; long long gstride;     // src row address = src + group*gstride + index*ld
  const float* src2; long long ld2; long long gstride2;  // optional second source
  const float* mix; int mmod;          // optional per-row multiplier on src2 (row % mmod)
  const float* scale; int smod;        // optional per-row scale
  __nv_bfloat16* dst; long long ldd; long long lo_off;
  int reps; long long rep_stride;      // the result is written `reps` times, rep_stride elements apart (0/1 = once)
  float4* zero_p; long long zero_n4;   // optional: also clear zero_n4 float4 at zero_p (a later kernel accumulates there)
};
__global__ void pack_hl_kernel(const PackParams p) {
  pdl_trigger();
  pdl_wait();
  const long long n = (long long)p.rows * p.cols;
  if (p.zero_p)
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.zero_n4; i += (long long)gridDim.x * blockDim.x)
      p.zero_p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / p.cols), c = (int)(i % p.cols);
    const int grp = p.rpg > 0 ? r / p.rpg : 0, idx = p.rpg > 0 ? r % p.rpg : r;
    float v = p.src[grp * p.gstride + (long long)idx * p.ld + c];
    if (p.src2) v += (p.mix ? p.mix[r % p.mmod] : 1.0f) * p.src2[grp * p.gstride2 + (long long)idx * p.ld2 + c];
    if (p.scale) v *= p.scale[r % p.smod];
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    const int reps = p.reps > 1 ? p.reps : 1;
    for (int k = 0; k < reps; ++k) {
      p.dst[k * p.rep_stride + (long long)r * p.ldd + c] = h;
      p.dst[k * p.rep_stride + (long long)r * p.ldd + p.lo_off + c] = l;
    }
  }
}
int pack_hl(const PackParams& p, cudaStream_t stream) {
  const long long n = (long long)p.rows * p.cols;
  if (n <= 0) return 0;
  const int grid = (int)llmin((n + 255) / 256, 148 * 8);
  SGG_LAUNCH(pack_hl_kernel, grid, 256, 0, stream, p);
  return 0;
}

// ------------------------------------------------------------------------------------ embeddings
// Per (t, b): u_fake (GEMM result, fp32, row t*B+b), u_real = W_emb[label] (fp32 master: exactly
// one_hot @ W of disc:86-87), u_int = u_real + alpha_b (u_fake - u_real) (linearity of disc:87 in
// tfgan's interpolates).  Written hi/lo into the x buffers' u columns of the stream row blocks.
struct EmbedMixParams {
  int B, T, E;
  int V;                                // label ids are clamped to [0, V): a bad id cannot read outside W_emb (the host validates)
  const float* Uf; long long ldUf;      // [T*B, E] or null (no fake stream)
  const int64_t* labels;                // [B, T] or null (no real stream)
  const float* Wemb;                    // [V, E] fp32 master
  const float* gp_alpha;                // [B] or null (no interp stream)
  int blk_fake, blk_real, blk_int;      // row blocks (or -1)
  __nv_bfloat16* X; long long ldX; long long x_lo; long long strideT; int uoff; int NR;
};
__global__ void embed_mix_kernel(const EmbedMixParams p) {
  pdl_trigger();
  pdl_wait();
  const long long n = (long long)p.T * p.B * p.E;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i % p.E);
    const long long tb = i / p.E;
    const int b = (int)(tb % p.B), t = (int)(tb / p.B);
    const float uf = p.Uf ? p.Uf[tb * p.ldUf + e] : 0.f;
    float ur = 0.f;
    if (p.labels) {
      long long id = p.labels[(long long)b * p.T + t];
      id = id < 0 ? 0 : (id >= p.V ? p.V - 1 : id);
      ur = p.Wemb[id * p.E + e];
    }
    __nv_bfloat16* xt = p.X + (long long)t * p.strideT + p.uoff + e;
    __nv_bfloat16 h, l;
    if (p.blk_fake >= 0) {
      split_bf16(uf, h, l);
      const long long row = (long long)p.blk_fake * p.B + b;
      xt[row * p.ldX] = h; xt[row * p.ldX + p.x_lo] = l;
    }
    if (p.blk_real >= 0) {
      split_bf16(ur, h, l);
      const long long row = (long long)p.blk_real * p.B + b;
      xt[row * p.ldX] = h; xt[row * p.ldX + p.x_lo] = l;
    }
    if (p.blk_int >= 0) {
      split_bf16(ur + p.gp_alpha[b] * (uf - ur), h, l);
      const long long row = (long long)p.blk_int * p.B + b;
      xt[row * p.ldX] = h; xt[row * p.ldX + p.x_lo] = l;
    }
  }
}
int embed_mix(const EmbedMixParams& p, cudaStream_t stream) {
  const long long n = (long long)p.T * p.B * p.E;
  const int grid = (int)llmin((n + 255) / 256, 148 * 8);
  SGG_LAUNCH(embed_mix_kernel, grid, 256, 0, stream, p);
  return 0;
}

// dW_emb[label[b,t], :] += u_bar_real + (1 - alpha_b) u_bar_int   (one-hot rows of x^T u_bar)
struct EmbedScatterParams {
  int B, T, E;
  int V;                                // ids clamped like embed_mix
  const int64_t* labels; const float* gp_alpha;
  const float* XB; long long ldXB; long long strideT; int uoff;   // u_bar = XB[t][row, uoff:uoff+E]
  int blk_real, blk_int;
  float* dWemb;
};
__global__ void embed_scatter_kernel(const EmbedScatterParams p) {
  pdl_trigger();
  pdl_wait();
  const long long n = (long long)p.T * p.B * p.E;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i % p.E);
    const long long tb = i / p.E;
    const int b = (int)(tb % p.B), t = (int)(tb / p.B);
    const float* xb = p.XB + (long long)t * p.strideT + p.uoff + e;
    float v = xb[((long long)p.blk_real * p.B + b) * p.ldXB];
    if (p.blk_int >= 0) v += (1.0f - p.gp_alpha[b]) * xb[((long long)p.blk_int * p.B + b) * p.ldXB];
    long long id = p.labels[(long long)b * p.T + t];
    id = id < 0 ? 0 : (id >= p.V ? p.V - 1 : id);
    atomicAdd(p.dWemb + id * p.E + e, v);
  }
}
int embed_scatter(const EmbedScatterParams& p, cudaStream_t stream) {
  const long long n = (long long)p.T * p.B * p.E;
  const int grid = (int)llmin((n + 255) / 256, 148 * 8);
  SGG_LAUNCH(embed_scatter_kernel, grid, 256, 0, stream, p);
  return 0;
}

// ------------------------------------------------------------------------------------ GP slopes
// g [T*B, ld] (row t*B+b).  slopes_b = sqrt(sum_{t,v} g^2 + 1e-10); pen = mean_b max(0, s-1)^2;
// coef_b = lam_scale * (2/Bglobal) max(0, s-1)/s   (d pen / d g = coef * g).   scal[2] += pen part.
struct GpSlopesParams {
  int B, T, V; const float* g; long long ld;
  float* slopes; float* coef; float* scal;  // scal[2] (gp) accumulated atomically
  float inv_Bglobal;
  __nv_bfloat16* vhl; long long ldv; long long v_lo;   // optional: v = coef_b * g as a bf16 hi/lo pair (the tangent direction)
};
__global__ void __launch_bounds__(256) gp_slopes_kernel(const GpSlopesParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8];
  const int b = blockIdx.x;
  float s = 0.f;
  for (int t = 0; t < p.T; ++t) {
    const float* row = p.g + ((long long)t * p.B + b) * p.ld;
    for (int v = threadIdx.x; v < p.V; v += 256) s = fmaf(row[v], row[v], s);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < 8; ++w) tot += red[w];
    const float slope = sqrtf(tot + 1e-10f);
    const float ex = fmaxf(slope - 1.0f, 0.f);
    p.slopes[b] = slope;
    p.coef[b] = 2.0f * p.inv_Bglobal * ex / slope;
    red[0] = p.coef[b];
    atomicAdd(p.scal + 2, ex * ex * p.inv_Bglobal);
  }
  if (p.vhl == nullptr) return;
  __syncthreads();
  const float coef = red[0];   // the rows of this sample were just read: they come back from L1 / L2
  for (int t = 0; t < p.T; ++t) {
    const float* row = p.g + ((long long)t * p.B + b) * p.ld;
    __nv_bfloat16* dst = p.vhl + ((long long)t * p.B + b) * p.ldv;
    for (int v = threadIdx.x; v < p.V; v += 256) {
      __nv_bfloat16 h, l;
      split_bf16(coef * row[v], h, l);
      dst[v] = h;
      dst[p.v_lo + v] = l;
    }
  }
}
int gp_slopes(const GpSlopesParams& p, cudaStream_t stream) {
  SGG_LAUNCH(gp_slopes_kernel, p.B, 256, 0, stream, p);
  return 0;
}

// ------------------------------------------------------------------------------------ losses
// Y [NR, T].  scal[1] += (sum Y[blk_fake] - sum Y[blk_real]) * inv ; scal[3] += -sum Y[blk_fake] * inv
struct LossParams { int B, T; const float* Y; int blk_fake, blk_real; float inv; float* scal; };
__global__ void __launch_bounds__(256) loss_kernel(const LossParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[2][8];
  float sf = 0.f, sr = 0.f;
  const int n = p.B * p.T;
  for (int i = threadIdx.x; i < n; i += 256) {
    if (p.blk_fake >= 0) sf += p.Y[(long long)p.blk_fake * n + i];
    if (p.blk_real >= 0) sr += p.Y[(long long)p.blk_real * n + i];
  }
  sf = warp_sum(sf); sr = warp_sum(sr);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sf; red[1][threadIdx.x >> 5] = sr; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; }
    atomicAdd(p.scal + 1, (a - b) * p.inv);
    atomicAdd(p.scal + 3, -a * p.inv);
  }
}
int losses(const LossParams& p, cudaStream_t stream) {
  SGG_LAUNCH(loss_kernel, 1, 256, 0, stream, p);
  return 0;
}

// column sums: out[c] += sum_r src[r, c]
__global__ void colsum_kernel(const float* src, long long ld, int rows, int cols, float* out) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float s = 0.f;
  for (int r = blockIdx.y; r < rows; r += gridDim.y) s += src[(long long)r * ld + c];
  atomicAdd(out + c, s);
}
int colsum(const float* src, long long ld, int rows, int cols, float* out, cudaStream_t stream) {
  dim3 grid((cols + 127) / 128, min(rows, 32));
  SGG_LAUNCH(colsum_kernel, grid, 128, 0, stream, src, ld, rows, cols, out);
  return 0;
}

// ------------------------------------------------------------------------------------ annotation adjoint
// What a conv front-end (gen:29-68) back-propagates: d cost / d self.downsampled.  The tile a[b] enters the recurrent half
// through z_t = sum_r alpha_r a_r (gen:17; the gradient penalty's tangent adds zdot_t = sum_r adot_r a_r), through the
// hoisted projection P = flat(a) W_a (gen:14-15) and through c0 = h0 = mean_r a_r (gen:76-77):
//   a_bar[b, r, :] = sum_{t, v} alpha[t][v][b][r] * z_bar[t][v][b][:]  +  (P_bar W_a^T)[b, r, :]
//                  + (1/R) sum_{primal v} (c0_bar + h0_bar)[v][b][:]
// The middle term is a GEMM (plan.cu net_ann_grad) that runs first and overwrites `out`; this kernel adds the other two:
// a rank-(T * streams) update per sample from the saved alpha rows and the z_bar columns of x_bar.  The tangent vector's
// "alpha" row is adot (zero at t = 0, where its buffer is not written) and its start state is the constant 0.
constexpr int AG_KT = 16;      // (t, v) pairs per pass over the output chunk
constexpr int AG_RCH = 49;     // regions per CTA (196 = 4 * 49)
struct AnnGradParams {
  int B, R, T, nv;
  int row_blk[8];              // row block of vector v in alpha / XB / CB0
  int tan_v;                   // index of the tangent vector, or -1
  const float* alpha; long long ldA, strideA;            // alpha + t * strideA + (row_blk[v] * B + b) * ldA
  const float* XB; long long ldXB, strideXB; int hoff;   // z_bar = XB[t][row][0:512], h0_bar = XB[0][row][hoff:hoff+512]
  const float* CB0;            // [rows, 512] total c_bar of step 0
  float* out;                  // [B, R, 512] fp32, accumulated into
};
__global__ void __launch_bounds__(128) ann_grad_kernel(const AnnGradParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ float al[AG_KT][AG_RCH];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int r0 = blockIdx.y * AG_RCH;
  const int nr = min(AG_RCH, p.R - r0);
  const int K = p.T * p.nv;
  float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int v = 0; v < p.nv; ++v) {
    if (v == p.tan_v) continue;
    const long long row = (long long)p.row_blk[v] * p.B + b;
    const float4 c = *reinterpret_cast<const float4*>(p.CB0 + row * 512 + tid * 4);
    const float4 h = *reinterpret_cast<const float4*>(p.XB + row * p.ldXB + p.hoff + tid * 4);
    s0.x += c.x + h.x; s0.y += c.y + h.y; s0.z += c.z + h.z; s0.w += c.w + h.w;
  }
  const float invR = 1.0f / (float)p.R;
  s0.x *= invR; s0.y *= invR; s0.z *= invR; s0.w *= invR;
  float* out_b = p.out + ((long long)b * p.R + r0) * 512 + tid * 4;
  for (int k0 = 0; k0 < K; k0 += AG_KT) {
    float4 zb[AG_KT];
#pragma unroll
    for (int kk = 0; kk < AG_KT; ++kk) {
      const int k = k0 + kk, t = k / p.nv, v = k % p.nv;
      zb[kk] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < K && !(v == p.tan_v && t == 0))
        zb[kk] = *reinterpret_cast<const float4*>(p.XB + t * p.strideXB + ((long long)p.row_blk[v] * p.B + b) * p.ldXB + tid * 4);
    }
    __syncthreads();   // the previous pass has finished reading al
    for (int i = tid; i < AG_KT * nr; i += 128) {
      const int kk = i / nr, rr = i - kk * nr;
      const int k = k0 + kk, t = k / p.nv, v = k % p.nv;
      float a = 0.f;
      if (k < K && !(v == p.tan_v && t == 0))
        a = p.alpha[t * p.strideA + ((long long)p.row_blk[v] * p.B + b) * p.ldA + r0 + rr];
      al[kk][rr] = a;
    }
    __syncthreads();
    for (int rr = 0; rr < nr; ++rr) {
      float4 acc = *reinterpret_cast<const float4*>(out_b + (long long)rr * 512);
      if (k0 == 0) { acc.x += s0.x; acc.y += s0.y; acc.z += s0.z; acc.w += s0.w; }
#pragma unroll
      for (int kk = 0; kk < AG_KT; ++kk) {
        const float a = al[kk][rr];
        acc.x = fmaf(a, zb[kk].x, acc.x); acc.y = fmaf(a, zb[kk].y, acc.y);
        acc.z = fmaf(a, zb[kk].z, acc.z); acc.w = fmaf(a, zb[kk].w, acc.w);
      }
      *reinterpret_cast<float4*>(out_b + (long long)rr * 512) = acc;
    }
  }
}
int ann_grad(const AnnGradParams& p, cudaStream_t stream) {
  dim3 grid(p.B, (p.R + AG_RCH - 1) / AG_RCH);
  SGG_LAUNCH(ann_grad_kernel, grid, 128, 0, stream, p);
  return 0;
}

}  // namespace sgg

// ------------------------------------------------------------------------------------ Adam
// tf.train.AdamOptimizer update (train:258-259): m,v moments; lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
// theta -= lr_t * m / (sqrt(v) + eps).  Flat fp32 buckets; also refreshes the bf16 shadow of
// each tensor (the GEMM B operands), whose rows are padded to a 16-byte pitch.
namespace sgg {
// sh2_off >= 0: a second shadow copy with the columns in the fused gate kernel's interleaved order (gates.cu
// perm_gate_col: groups of 32 hidden units, [i | j | f | o] inside a group); same pitch and lo offset.
struct AdamSeg { long long off; long long sh_off; int cols; int pitch; long long n; long long lo_off; long long sh2_off; };
__host__ __device__ __forceinline__ unsigned perm_gate_col4(unsigned c) { return ((c & 511u) >> 5) * 128u + (c >> 9) * 32u + (c & 31u); }
constexpr int ADAM_MAX_SEG = 24;
constexpr int ADAM_UNROLL = 4;                         // float4 quadruples in flight per thread and array
constexpr int ADAM_TILE = 256 * 4 * ADAM_UNROLL;       // floats per CTA
struct AdamParams {
  float* theta; const float* grad; float* m; float* v; __nv_bfloat16* shadow;
  float lr_t, b1, b2, eps, gscale;
  // device-side step counter (graph replay): step = iter[0] * step_mul + step_add, lr_t recomputed in the kernel
  const long long* iter; long long step_mul, step_add; float lr;
  int stream_l2;                                       // streaming (evict-first) accesses to the buckets
  int nseg; AdamSeg seg[ADAM_MAX_SEG];
  int blk_start[ADAM_MAX_SEG + 1];                     // CTA range of each segment (one flat 1-D grid)
};
// One CTA updates ADAM_TILE consecutive floats of one tensor: all loads of the tile are issued before the first
// use (16 x 16 B in flight per thread), the bias-corrected step size is computed by one thread meanwhile.
__global__ void __launch_bounds__(256) adam_kernel(const __grid_constant__ AdamParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_lr;
  int s = 0;
  while (s + 1 < p.nseg && (int)blockIdx.x >= p.blk_start[s + 1]) ++s;
  const AdamSeg sg = p.seg[s];
  const long long base = (long long)((int)blockIdx.x - p.blk_start[s]) * ADAM_TILE + threadIdx.x * 4;
  float th[ADAM_UNROLL][4], g[ADAM_UNROLL][4], m[ADAM_UNROLL][4], v[ADAM_UNROLL][4];
  const bool aligned = (sg.off & 3) == 0;
#pragma unroll
  for (int u = 0; u < ADAM_UNROLL; ++u) {
    const long long i = base + u * 1024;
    const long long gi = sg.off + i;
    if (aligned && i + 4 <= sg.n) {
      const float4 g4 = __ldcs(reinterpret_cast<const float4*>(p.grad + gi));   // gradients are dead after this read
      float4 t4, m4, v4;
      if (p.stream_l2) {   // the whole bucket is touched once per step: do not displace the annotation tiles in L2
        t4 = __ldcs(reinterpret_cast<const float4*>(p.theta + gi));
        m4 = __ldcs(reinterpret_cast<const float4*>(p.m + gi));
        v4 = __ldcs(reinterpret_cast<const float4*>(p.v + gi));
      } else {
        t4 = *reinterpret_cast<const float4*>(p.theta + gi);
        m4 = *reinterpret_cast<const float4*>(p.m + gi);
        v4 = *reinterpret_cast<const float4*>(p.v + gi);
      }
      th[u][0] = t4.x; th[u][1] = t4.y; th[u][2] = t4.z; th[u][3] = t4.w;
      g[u][0] = g4.x; g[u][1] = g4.y; g[u][2] = g4.z; g[u][3] = g4.w;
      m[u][0] = m4.x; m[u][1] = m4.y; m[u][2] = m4.z; m[u][3] = m4.w;
      v[u][0] = v4.x; v[u][1] = v4.y; v[u][2] = v4.z; v[u][3] = v4.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool ok = i + e < sg.n;
        th[u][e] = ok ? p.theta[gi + e] : 0.f; g[u][e] = ok ? p.grad[gi + e] : 0.f;
        m[u][e] = ok ? p.m[gi + e] : 0.f; v[u][e] = ok ? p.v[gi + e] : 0.f;
      }
    }
  }
  if (threadIdx.x == 0) {
    float lr_t = p.lr_t;
    if (p.iter) {
      const double step = (double)(p.iter[0] * p.step_mul + p.step_add);
      lr_t = (float)((double)p.lr * sqrt(1.0 - pow((double)p.b2, step)) / (1.0 - pow((double)p.b1, step)));
    }
    s_lr = lr_t;
  }
  __syncthreads();
  const float lr_t = s_lr;
#pragma unroll
  for (int u = 0; u < ADAM_UNROLL; ++u) {
    const long long i = base + u * 1024;
    if (i >= sg.n) continue;
    const long long gi = sg.off + i;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float gg = g[u][e] * p.gscale;
      m[u][e] = p.b1 * m[u][e] + (1.0f - p.b1) * gg;
      v[u][e] = p.b2 * v[u][e] + (1.0f - p.b2) * gg * gg;
      th[u][e] -= lr_t * m[u][e] / (sqrtf(v[u][e]) + p.eps);
    }
    const bool vec = aligned && i + 4 <= sg.n;
    if (vec) {
      if (p.stream_l2) {
        __stcs(reinterpret_cast<float4*>(p.theta + gi), make_float4(th[u][0], th[u][1], th[u][2], th[u][3]));
        __stcs(reinterpret_cast<float4*>(p.m + gi), make_float4(m[u][0], m[u][1], m[u][2], m[u][3]));
        __stcs(reinterpret_cast<float4*>(p.v + gi), make_float4(v[u][0], v[u][1], v[u][2], v[u][3]));
      } else {
        *reinterpret_cast<float4*>(p.theta + gi) = make_float4(th[u][0], th[u][1], th[u][2], th[u][3]);
        *reinterpret_cast<float4*>(p.m + gi) = make_float4(m[u][0], m[u][1], m[u][2], m[u][3]);
        *reinterpret_cast<float4*>(p.v + gi) = make_float4(v[u][0], v[u][1], v[u][2], v[u][3]);
      }
    } else {
      for (int e = 0; e < 4; ++e)
        if (i + e < sg.n) { p.theta[gi + e] = th[u][e]; p.m[gi + e] = m[u][e]; p.v[gi + e] = v[u][e]; }
    }
    if (p.shadow && sg.sh_off >= 0) {
      // bf16 hi/lo shadow of the updated weights (row pitch may exceed the column count)
      unsigned r = (unsigned)i / (unsigned)sg.cols, c = (unsigned)i - r * (unsigned)sg.cols;
      __nv_bfloat16 h[4], l[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) split_bf16(th[u][e], h[e], l[e]);
      if (vec && ((sg.cols | sg.pitch) & 3) == 0) {   // the 4 elements share a row and are 8-byte aligned
        __nv_bfloat16* dst = p.shadow + sg.sh_off + (long long)r * sg.pitch + c;
        *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
        *reinterpret_cast<uint2*>(dst + sg.lo_off) = make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
        if (sg.sh2_off >= 0) {   // 4 consecutive units of one gate stay consecutive in the interleaved order
          __nv_bfloat16* d2 = p.shadow + sg.sh2_off + (long long)r * sg.pitch + perm_gate_col4(c);
          *reinterpret_cast<uint2*>(d2) = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
          *reinterpret_cast<uint2*>(d2 + sg.lo_off) = make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
        }
      } else {
        for (int e = 0; e < 4; ++e)
          if (i + e < sg.n) {
            __nv_bfloat16* dst = p.shadow + sg.sh_off + (long long)r * sg.pitch + c;
            dst[0] = h[e];
            dst[sg.lo_off] = l[e];
            if (sg.sh2_off >= 0) {
              __nv_bfloat16* d2 = p.shadow + sg.sh2_off + (long long)r * sg.pitch + perm_gate_col4(c);
              d2[0] = h[e];
              d2[sg.lo_off] = l[e];
            }
            if (++c == (unsigned)sg.cols) { c = 0; ++r; }
          }
      }
    }
  }
}

int adam(AdamParams p, cudaStream_t stream) {
  int nb = 0;
  for (int s = 0; s < p.nseg; ++s) {
    p.blk_start[s] = nb;
    nb += (int)((p.seg[s].n + ADAM_TILE - 1) / ADAM_TILE);
  }
  p.blk_start[p.nseg] = nb;
  if (nb == 0) return 0;
  p.stream_l2 = l2_policy_enabled() ? 1 : 0;
  SGG_LAUNCH(adam_kernel, nb, 256, 0, stream, p);
  return 0;
}

// fp32 master -> bf16 shadow refresh only (initialisation / after loading parameters)
__global__ void shadow_kernel(const float* theta, __nv_bfloat16* shadow, AdamSeg sg) {
  pdl_trigger();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < sg.n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / sg.cols, c = i % sg.cols;
    __nv_bfloat16 h, l;
    split_bf16(theta[sg.off + i], h, l);
    shadow[sg.sh_off + r * sg.pitch + c] = h;
    shadow[sg.sh_off + sg.lo_off + r * sg.pitch + c] = l;
    if (sg.sh2_off >= 0) {
      const long long c2 = perm_gate_col4((unsigned)c);
      shadow[sg.sh2_off + r * sg.pitch + c2] = h;
      shadow[sg.sh2_off + sg.lo_off + r * sg.pitch + c2] = l;
    }
  }
}
int refresh_shadow(const float* theta, __nv_bfloat16* shadow, const AdamSeg& sg, cudaStream_t stream) {
  if (sg.sh_off < 0 || sg.n <= 0) return 0;
  const int grid = (int)llmin((sg.n + 255) / 256, 148 * 8);
  SGG_LAUNCH(shadow_kernel, grid, 256, 0, stream, theta, shadow, sg);
  return 0;
}

// ------------------------------------------------------------------------------------ Philox RNG
// mode 0: uniform [0,1) ; mode 1: standard normal (Box-Muller)
__global__ void rng_fill_kernel(float* out, long long n, uint64_t seed, uint64_t offset, int mode,
                                const long long* iter, uint64_t per_iter) {
  pdl_trigger();
  pdl_wait();
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q * 4 >= n) return;
  if (iter) offset += (uint64_t)iter[0] * per_iter;   // device-side stream position (graph replay)
  const uint64_t ctr = offset + (uint64_t)q;
  const uint4 r = philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  float o[4];
  if (mode == 0) {
    o[0] = (r.x >> 8) * (1.0f / 16777216.0f); o[1] = (r.y >> 8) * (1.0f / 16777216.0f);
    o[2] = (r.z >> 8) * (1.0f / 16777216.0f); o[3] = (r.w >> 8) * (1.0f / 16777216.0f);
  } else {
    const float r0 = sqrtf(-2.0f * logf(u01(r.x))), r1 = sqrtf(-2.0f * logf(u01(r.z)));
    float s0, c0, s1, c1;
    sincosf(6.283185307179586f * u01(r.y), &s0, &c0);
    sincosf(6.283185307179586f * u01(r.w), &s1, &c1);
    o[0] = r0 * c0; o[1] = r0 * s0; o[2] = r1 * c1; o[3] = r1 * s1;
  }
  for (int e = 0; e < 4; ++e)
    if (q * 4 + e < n) out[q * 4 + e] = o[e];
}
int rng_fill(float* out, long long n, uint64_t seed, uint64_t offset, int mode, cudaStream_t stream,
             const long long* iter = nullptr, uint64_t per_iter = 0) {
  if (n <= 0) return 0;
  const long long quads = (n + 3) / 4;
  SGG_LAUNCH(rng_fill_kernel, (int)((quads + 255) / 256), 256, 0, stream, out, n, seed, offset, mode, iter, per_iter);
  return 0;
}

// iteration counter that drives the device-side RNG position and Adam step numbers
__global__ void bump_counter_kernel(long long* ctr) {
  pdl_trigger();
  pdl_wait();
  ctr[0] += 1;
}
int bump_counter(long long* ctr, cudaStream_t stream) {
  SGG_LAUNCH(bump_counter_kernel, 1, 1, 0, stream, ctr);
  return 0;
}

// Row-sharded attention projection: columns [p*Ks, (p+1)*Ks) of flat(a) [B, K] -> send[p][b][0:Ks] (zero beyond K).
__global__ void slab_pack_kernel(const uint4* a, uint4* send, int B, long long K8, long long Ks8, int world) {
  pdl_trigger();
  pdl_wait();
  const long long n = (long long)world * B * Ks8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long j = i % Ks8, pb = i / Ks8;
    const int b = (int)(pb % B), p = (int)(pb / B);
    const long long col = p * Ks8 + j;
    send[i] = col < K8 ? __ldg(a + (long long)b * K8 + col) : make_uint4(0u, 0u, 0u, 0u);
  }
}
int slab_pack(const __nv_bfloat16* a, __nv_bfloat16* send, int B, long long K, long long Ks, int world, cudaStream_t st) {
  if ((K & 7) || (Ks & 7)) { set_error("slab_pack: K and Ks must be multiples of 8"); return -1; }
  const long long n = (long long)world * B * (Ks / 8);
  const int grid = (int)llmin((n + 255) / 256, 148 * 16);
  SGG_LAUNCH(slab_pack_kernel, grid, 256, 0, st, reinterpret_cast<const uint4*>(a), reinterpret_cast<uint4*>(send), B, K / 8,
             Ks / 8, world);
  return 0;
}

// argmax keys of the decoder GEMM's sampling epilogue -> token ids (train:270)
__global__ void decode_keys_kernel(const unsigned long long* keys, int32_t* tokens, long long n) {
  pdl_trigger();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) tokens[i] = (int32_t)(0xFFFFFFFFu - (uint32_t)(keys[i] & 0xFFFFFFFFull));
}
int decode_keys(const unsigned long long* keys, int32_t* tokens, long long n, cudaStream_t st) {
  if (n <= 0) return 0;
  SGG_LAUNCH(decode_keys_kernel, (int)((n + 255) / 256), 256, 0, st, keys, tokens, n);
  return 0;
}

}  // namespace sgg

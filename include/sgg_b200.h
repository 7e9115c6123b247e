/* sgg_b200 -- C ABI of the B200-native Scene-Graph-GAN training hot path.
 *
 * This is the drop-in boundary for the reference's recurrent attention-LSTM generator /
 * discriminator and its WGAN-GP step.  The reference (pure Python + TensorFlow 1.x) has no
 * FFI layer of its own; each entry point below names the reference code it replaces
 * (gen: = architectures/generator_with_attention.py, disc: =
 * architectures/discriminator_with_attention.py, train: = train.py).  INTEGRATION.md shows
 * the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes only; every function returns 0 on success and a
 *    negative status on failure, with the message available from sgg_last_error().
 *  - The caller owns every buffer (parameters, gradients, workspace, outputs).  Nothing here
 *    calls cudaMalloc or synchronises the device; all work is enqueued on `stream`
 *    (graph-capturable).  Not thread-safe per workspace.
 *  - All device pointers must be 256-byte aligned.
 */
#ifndef SGG_B200_H_
#define SGG_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* sgg_stream_t; /* cudaStream_t */

/* ----------------------------------------------------------------------------------------
 * Library
 * -------------------------------------------------------------------------------------- */
const char* sgg_last_error(void);
int sgg_version(void);
/* Number of CUDA kernels this library has launched (or recorded into a stream capture) in this
 * process so far.  Lets a caller report how much of a timed region ran in these kernels. */
int64_t sgg_launch_count(void);
/* Run-time switches, for A/B comparisons inside one process.  "fused_gates": 0 = gate GEMM + cell kernel (default),
 * 1 = gate GEMM with the LayerNorm / cell epilogue fused in (csrc/gates.cu), 16 / 8 = force its cluster shape.
 * "attn_persistent": -1 = choose the attention kernel family by batch size (default), 0 = one CTA per sample,
 * 1 = persistent CTAs with the TMA ring running across samples. */
int sgg_set_option(const char* name, int32_t value);
/* Launches per kernel entry point so far, as "demangled kernel name;launches" lines (NUL-terminated, truncated to cap;
 * returns the size needed).  reset != 0 clears the counters.  Lets a test assert WHICH kernel variants a call ran. */
int64_t sgg_kernel_counts(char* buf, int64_t cap, int32_t reset);
/* Profiling aid: with SGG_TIMING=1 in the environment every eager (non-captured) kernel launch is bracketed by CUDA
 * events.  This call synchronises the device, writes one "kernel;grid;block;launches;total_us" line per distinct
 * launch shape timed since the previous call into buf (NUL-terminated, truncated to cap) and returns the size needed. */
int64_t sgg_timing_report(char* buf, int64_t cap);

/* ----------------------------------------------------------------------------------------
 * Problem dimensions.  Reference values: R=196 (14x14, gen:74-75), C=512 (gen:68), H=512
 * (gen:79), E=300 (train:63,70), T=3 (gen:85), U = C for G (noise, gen:81) / E for D.
 * -------------------------------------------------------------------------------------- */
typedef struct sgg_dims {
  int32_t B; /* per-GPU batch */
  int32_t T; /* timesteps (3 * triples) */
  int32_t V; /* vocabulary */
  int32_t R; /* annotation regions */
  int32_t C; /* annotation channels */
  int32_t H; /* LSTM units */
  int32_t E; /* embedding width */
  int32_t S; /* fake-logit slots kept in the workspace = max critic steps per sgg_train_iteration (0/1: one) */
} sgg_dims_t;

/* ----------------------------------------------------------------------------------------
 * Dense contraction on tcgen05 tensor cores (bf16 in, fp32 accumulate in TMEM).
 *   D[m,n] = alpha * sum_seg sum_k A_seg[m,k] * B_seg[n,k]  (+ bias[n]) (+ addm[m % add_mod, n])
 * Replaces every tf MatMul on the path: gen:15 (dense 196), LayerNormBasicLSTMCell kernel
 * (gen:79,87), gen:88 / disc:90 (decoder), disc:87 (embedding) and their registered
 * gradients.  Operands are 2-D row-major bf16 tensors; *_mn_major = 0 means the tensor is
 * [MN, K] (K contiguous), 1 means it is [K, MN] (MN contiguous; e.g. a TF [in,out] kernel
 * used as B, or an activation matrix used as A of a weight gradient).  Up to 4 K-segments
 * let one launch sum products of hi/lo bf16 split operands (x ~= hi + lo).
 * -------------------------------------------------------------------------------------- */
#define SGG_GEMM_MAX_SEG 4
typedef struct sgg_gemm_desc {
  const void* A; int64_t a_rows, a_cols, a_ld; int32_t a_mn_major;
  const void* B; int64_t b_rows, b_cols, b_ld; int32_t b_mn_major;
  int32_t M, N;
  int32_t nseg;
  int32_t seg_a_k[SGG_GEMM_MAX_SEG], seg_a_mn[SGG_GEMM_MAX_SEG];
  int32_t seg_b_k[SGG_GEMM_MAX_SEG], seg_b_mn[SGG_GEMM_MAX_SEG];
  int32_t seg_klen[SGG_GEMM_MAX_SEG]; /* k-blocks of 64; a ragged tail must fall outside both tensors */
  /* epilogue */
  float* C; int64_t ldc; int32_t atomic;        /* fp32 output (optional); 1 => red.add into the existing contents;
                                                   2 => the output is known to be zero-filled (no clear needed) */
  void* Chl; int64_t ld_hl; int64_t lo_off;     /* bf16 hi/lo split output (optional) */
  const float* bias;                            /* [N] optional */
  const float* addm; int64_t ld_addm; int32_t add_mod; /* optional broadcast add */
  float alpha;
  int32_t block_n; /* 0 = auto, else 64/128/256 */
  int32_t splits;  /* 0/1 = none; >1 requires atomic */
  /* optional output row permutation (out_d0 > 0): result row m is stored at row
   *   (m / out_d0) * out_s0 + ((m % out_d0) / out_d1) * out_s1 + (m % out_d1)
   * of C / Chl (e.g. [t][stream][b] rows -> [stream][t][b]). */
  int32_t out_d0, out_d1;
  int64_t out_s0, out_s1;
  /* optional sampling epilogue (train:270 tf.argmax over the vocabulary; Gumbel-max sampling as an extension):
   * argmax_keys[row * argmax_stride] (device uint64, zero-filled by the caller) receives, by 64-bit atomic max over
   * the n-tiles, (orderable bits of max_n D[row,n] (+ Gumbel noise)) << 32 | (0xFFFFFFFF - argmax column); the lowest
   * column wins ties.  gumbel != 0 adds g = -log(-log u), u = Philox4x32-10(gumbel_seed; gumbel_offset + row, n). */
  uint64_t* argmax_keys; int64_t argmax_stride;
  int32_t gumbel; uint64_t gumbel_seed, gumbel_offset;
} sgg_gemm_desc_t;

int sgg_gemm(const sgg_gemm_desc_t* d, sgg_stream_t stream);
/* Host-side query, launches nothing: the tile width and split-K sgg_gemm would choose for this descriptor, and the number
 * of 64-wide k-blocks one accumulator then runs over.  Split-K is also the accuracy device of this library: the tcgen05
 * fp32 accumulator truncates, so an automatically split contraction never keeps one accumulator for more than 128 k-blocks. */
int sgg_gemm_plan(const sgg_gemm_desc_t* d, int32_t* block_n, int32_t* splits, int32_t* k_blocks_per_split);

/* ----------------------------------------------------------------------------------------
 * Attention step (gen:16-17 / disc:16-17): for each of `nv` streams that share the annotation tile
 * of sample b (fake / real / interpolate rows v*B+b), alpha = softmax(e) over R and
 * z_hat = sum_r alpha_r a[b,r,:].  One pass over `a` ([B,R,512] bf16) serves all streams.
 *   e      [nv*B, ld_e] fp32 scores (= P + c W_h + bias, from sgg_gemm)
 *   alpha  [nv*B, ld_e] fp32 out (may alias e)
 *   z_hl   [nv*B, ld_z] bf16 out: z_hat hi part at columns [0,512), lo part at [lo_off, lo_off+512)
 * -------------------------------------------------------------------------------------- */
int sgg_attn_forward(const void* a, int32_t B, int32_t R, int32_t nv, const float* e, float* alpha, int64_t ld_e,
                     void* z_hl, int64_t ld_z, int64_t lo_off, sgg_stream_t stream);

/* Reverse of the attention step (the gradient TF registers for gen:16-17) for `nv` <= 4 streams sharing a tile:
 *   alpha_bar[r] = <z_bar, a[b,r,:]>,  e_bar = alpha * (alpha_bar - <alpha, alpha_bar>)   (softmax reverse)
 *   z_bar    [nv*B, ld_zb]   fp32 upstream of z_hat (columns [0,512))
 *   alpha    [nv*B, ld_alpha] fp32 saved softmax
 *   e_bar_hl [nv*B, ld_eb]   bf16 out: hi part at columns [0,R), lo part at [lo_off, lo_off+R)
 *   p_bar    [B, ld_p]       fp32, optional: += sum over streams of e_bar (the gradient of P = flat(a) W_a) */
int sgg_attn_reverse(const void* a, int32_t B, int32_t R, int32_t nv, const float* z_bar, int64_t ld_zb,
                     const float* alpha, int64_t ld_alpha, void* e_bar_hl, int64_t ld_eb, int64_t lo_off,
                     float* p_bar, int64_t ld_p, sgg_stream_t stream);

/* ----------------------------------------------------------------------------------------
 * Parameters.  Each network keeps ONE flat fp32 bucket (master weights; gradients and Adam
 * moments use the same layout) plus a bf16 "shadow" bucket holding the GEMM operands as a
 * hi/lo pair (w ~= hi + lo to 2^-17: three bf16 tensor-core products reproduce the fp32 product),
 * rows padded to a 16-byte pitch.  The table reproduces the reference's TF variable names
 * (train:262-263 splits them by the "Generator"/"Discriminator" prefix):
 *   <scope>/attention_perceptron/{kernel [R*C+H, R], bias [R]}                  (gen:15)
 *   <scope>/layer_norm_basic_lstm_cell/{kernel [C+U+H, 4H], <ln>/{gamma,beta}}  (gen:79)
 *   <scope>/decoder/{kernel [H, V|1], bias}                                     (gen:88, disc:90)
 *   Discriminator/W [V, E]                                                      (train:70)
 * net: 0 = generator, 1 = discriminator.  Offsets are in floats / bf16 elements.
 * -------------------------------------------------------------------------------------- */
typedef struct sgg_param_entry {
  char name[96];
  int64_t offset;
  int32_t rows, cols;
  int64_t shadow_offset; /* -1: tensor has no bf16 shadow (used in fp32) */
  int32_t shadow_pitch;  /* row pitch (elements) of the shadow */
  int32_t shadow_rows;   /* hi part = shadow rows [0, rows); lo part = rows [shadow_rows, shadow_rows + rows) */
} sgg_param_entry_t;

int sgg_param_table(int net, const sgg_dims_t* d, sgg_param_entry_t* out, int max_entries,
                    int* n_entries, int64_t* n_floats, int64_t* n_shadow);

/* Rebuilds the bf16 shadow from the fp32 master bucket (after init / load). */
int sgg_refresh_shadow(int net, const sgg_dims_t* d, const float* theta, void* shadow, sgg_stream_t stream);

/* tf.train.AdamOptimizer update (train:258-259) over the flat bucket, epsilon outside the bias
 * correction: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t*m/(sqrt(v)+eps).  Also refreshes
 * the bf16 shadow.  step is 1-based.  grad_scale multiplies the gradient (1/world after a
 * sum-allreduce of mean-normalised shards is NOT needed: the kernels already normalise by the
 * global batch). */
int sgg_adam_step(int net, const sgg_dims_t* d, float* theta, const float* grad, float* m, float* v,
                  void* shadow, int64_t step, float lr, float beta1, float beta2, float eps,
                  float grad_scale, sgg_stream_t stream);

/* Optimiser-fused attention projection (single GPU, B <= 256, R % 4 == 0): the Adam update of the annotation rows
 * W_a = attention_perceptron/kernel[0 : R*C, :] (same rule as sgg_adam_step) and, from the updated weights while they
 * are still on the SM, the hoisted projection  P[b, :] = flat(a)[b, :] W_a  (gen:14-15) of the NEXT pass, by split-K
 * reduction into P [B, rup(R,64)] fp32 (cleared by the call).  The bf16 hi/lo shadow of W_a is rewritten only when
 * write_shadow != 0 (a later pass with other annotations needs it).  theta / grad / m / v / shadow are the bucket bases
 * of network `net`; every other tensor of the bucket is left to sgg_adam_step. */
int sgg_adam_project(int net, const sgg_dims_t* d, float* theta, const float* grad, float* m, float* v, void* shadow,
                     int64_t step, float lr, float beta1, float beta2, float eps, const void* ann, float* P,
                     int32_t write_shadow, sgg_stream_t stream);

/* Philox4x32-10 counter RNG: tf.random_normal (gen:81) / tfgan's random_uniform alpha. */
int sgg_rng_fill_normal(float* out, int64_t n, uint64_t seed, uint64_t offset, sgg_stream_t stream);
int sgg_rng_fill_uniform(float* out, int64_t n, uint64_t seed, uint64_t offset, sgg_stream_t stream);

/* ----------------------------------------------------------------------------------------
 * Training / inference steps.  The workspace (sgg_workspace_bytes) must be zero-filled once
 * by the caller after allocation; its contents are private to the library.
 * -------------------------------------------------------------------------------------- */
int64_t sgg_workspace_bytes(const sgg_dims_t* d);

#define SGG_FLAG_REFRESH_GEN_PROJ 1 /* recompute the generator's hoisted projection P and c0: set on
                                       the first call for a new batch and after a generator update */

typedef struct sgg_step_args {
  sgg_dims_t dims;
  int32_t world;          /* data-parallel world size: losses are means over B*world samples */
  float lam;              /* gradient_penalty_weight (train:249, --lambda) */
  const float* g_theta; const void* g_shadow; float* g_grad;
  const float* d_theta; const void* d_shadow; float* d_grad;
  const void* ann_g;      /* [B,R,C] bf16: generator's annotations (gen:68 self.downsampled) */
  const void* ann_d;      /* [B,R,C] bf16: discriminator's annotations (disc:68) */
  const int64_t* labels;  /* [B,T] class ids of the real triples (one-hot of train:173) */
  const float* noise;     /* [B,C] N(0,1), shared by all timesteps (gen:81,86) */
  const float* gp_alpha;  /* [B] U[0,1) interpolation coefficients (tfgan gradient penalty) */
  void* workspace; int64_t workspace_bytes;
  float* scalars;         /* [4] device floats: {-, w_disc, gp, gen_cost} */
  float* logits_out;      /* optional [B,T,V] fp32 generator logits */
  int32_t flags;
  /* Optional outputs for a caller that trains the convolutional front-end (gen:29-68 / disc:29-68, whose variables are
   * in the var_lists of train:262-263): the adjoint of the annotations, [B,R,C] fp32, overwritten.
   *   sgg_gen_step : ann_g_grad = d gen_cost / d ann_g      sgg_disc_step : ann_d_grad = d disc_cost / d ann_d
   * (the generator is a constant of the discriminator step and vice versa, so the other pointer is ignored by each).
   * Costs one GEMM (P_bar W_a^T, M = B, N = R*C, K = R) and one HBM pass over the output per step.  Step-level entry
   * points only: sgg_train_iteration updates no front-end between its steps and rejects non-NULL pointers. */
  float* ann_g_grad;
  float* ann_d_grad;
} sgg_step_args_t;

/* gen:74-91 (from self.downsampled): logits [B,T,V] into logits_out. */
int sgg_gen_forward(const sgg_step_args_t* a, sgg_stream_t stream);
/* disc:73-93 on caller-supplied float triples [B,T,V]: scores [B,T]. */
int sgg_disc_forward(const sgg_step_args_t* a, const float* triples, float* scores_out, sgg_stream_t stream);
/* train:365 minus the optimizer: d disc_cost / d Discriminator* into d_grad (overwritten),
 * disc_cost = scalars[1] + lam * scalars[2]  (train:245-253); optionally d disc_cost / d ann_d into ann_d_grad. */
int sgg_disc_step(const sgg_step_args_t* a, sgg_stream_t stream);
/* train:368 minus the optimizer: d gen_cost / d Generator* into g_grad, scalars[3] = gen_cost; optionally
 * d gen_cost / d ann_g into ann_g_grad. */
int sgg_gen_step(const sgg_step_args_t* a, sgg_stream_t stream);

/* ----------------------------------------------------------------------------------------
 * Generator sampling (inference): gen:74-91 forward only, then the reference's test-time decoding
 * (train:270: tf.argmax(generator_output, axis=2)) or, as an extension, Gumbel-max sampling
 * token ~ softmax(logits).  The batch is processed in chunks of `chunk` images whose buffers are
 * reused every timestep; the logits only reach HBM when logits_out is given.
 * -------------------------------------------------------------------------------------- */
#define SGG_SAMPLE_GREEDY 0
#define SGG_SAMPLE_GUMBEL 1
typedef struct sgg_sample_args {
  sgg_dims_t dims;            /* B = images of this call; S and E are ignored */
  const float* g_theta; const void* g_shadow;
  const void* ann_g;          /* [B,R,C] bf16 annotations */
  const float* noise;         /* [B,C] N(0,1) (gen:81), or NULL: drawn on the device from (seed, offset) */
  int32_t mode;               /* SGG_SAMPLE_GREEDY | SGG_SAMPLE_GUMBEL */
  int32_t chunk;              /* images per chunk (0 = default: B/2 when B >= 2048, else B); consecutive chunks
                                 alternate between `stream` and a library-owned second stream */
  uint64_t seed, offset;      /* Philox key / position of the noise and Gumbel draws */
  void* workspace; int64_t workspace_bytes;   /* sgg_sample_workspace_bytes(dims, chunk) */
  int32_t* tokens_out;        /* [B,T] token ids */
  float* logits_out;          /* optional [B,T,V] fp32 raw logits */
} sgg_sample_args_t;
int64_t sgg_sample_workspace_bytes(const sgg_dims_t* d, int32_t chunk);
int sgg_gen_sample(const sgg_sample_args_t* a, sgg_stream_t stream);

/* ----------------------------------------------------------------------------------------
 * Data-parallel exchange (the reference is single-GPU).  One process per GPU; the
 * batch is sharded over ranks and every loss is normalised by the GLOBAL batch (args.world), so
 * the only exchange is a sum of the flat gradient bucket before each optimiser step.  The
 * communicator is an NCCL communicator owned by this library (libnccl.so.2 is resolved at run
 * time from the process, i.e. the copy PyTorch ships); the 128-byte unique id is created on
 * rank 0 and distributed by the caller (torch.distributed broadcast).
 * -------------------------------------------------------------------------------------- */
#define SGG_COMM_ID_BYTES 128
int sgg_comm_unique_id(void* id_out /* SGG_COMM_ID_BYTES */);
int sgg_comm_init(const void* id, int32_t rank, int32_t world, void** comm_out);
int sgg_comm_destroy(void* comm);
int sgg_comm_allreduce_sum(void* comm, float* buf, int64_t n, sgg_stream_t stream);

/* ----------------------------------------------------------------------------------------
 * One training iteration (train:362-368 on one batch, train:185-187): critic_iters x
 * { D step, [allreduce], Adam(D) } then { G step, [allreduce], Adam(G) }.  Noise (gen:81) and
 * the gradient-penalty interpolation coefficients are drawn on the device, fresh for every
 * step.  The generator forwards of all critic steps are batched (the generator is constant
 * while the critic trains).  step.noise / step.gp_alpha / step.scalars / step.flags are ignored.
 * The call depends on the host only through its arguments: the iteration counter that drives
 * the RNG position and the Adam step numbers lives in `counters`, so a stream capture of one
 * call can be replayed as a CUDA graph.
 * -------------------------------------------------------------------------------------- */
/* Row-sharded attention projection (world > 1, optional but recommended).  Data parallelism replicates every tensor
 * except the annotation rows W_a of attention_perceptron/kernel (85 % of the parameters): rank r owns the contraction
 * indices [r*Ks, (r+1)*Ks), Ks = R*C/world, of  P = flat(a) W_a  for the GLOBAL batch.  Per iteration the ranks exchange
 * annotation column slabs (all-to-all); per optimiser step the block needs a reduce-scatter of the partial projections
 * [B*world, R] and an all-gather of P_bar instead of an all-reduce of its 79 MB gradient, and its HBM traffic per GPU
 * (K1 operand, dW_a, Adam) drops by 1/world.  Only rows [r*Ks, (r+1)*Ks) of W_a (theta, m, v, shadow) are maintained
 * on rank r; gather them before reading the full tensor.  Results equal the replicated scheme up to fp32 summation
 * order. */
typedef struct sgg_wa_shard {
  int32_t enabled;
  void* slab_g; void* slab_d;            /* device bf16 [B*world, Ks] each: sgg_wa_shard_slab_elems() elements */
  void* scratch; int64_t scratch_bytes;  /* sgg_wa_shard_scratch_bytes() */
} sgg_wa_shard_t;
int64_t sgg_wa_shard_scratch_bytes(const sgg_dims_t* d, int32_t world);
int64_t sgg_wa_shard_slab_elems(const sgg_dims_t* d, int32_t world);

typedef struct sgg_iter_args {
  sgg_step_args_t step;
  int32_t critic_iters;              /* train:30 CRITIC_ITERS; <= step.dims.S */
  float* g_m; float* g_v;            /* Adam moments, generator bucket layout */
  float* d_m; float* d_v;            /* Adam moments, discriminator bucket layout */
  float lr, beta1, beta2, eps;       /* train:258-259: 1e-4, 0.5, 0.9 (eps 1e-8, TF form) */
  uint64_t seed;                     /* Philox key of this rank */
  int64_t* counters;                 /* device int64[1]: iterations done so far (zero-initialised by the caller) */
  float* noise_all;                  /* device [(critic_iters+1), B, C] scratch */
  float* gp_alpha_all;               /* device [critic_iters, B] scratch */
  float* scalars_all;                /* device [(critic_iters+1), 4]: per-step {-, w_disc, gp, gen_cost} */
  void* comm;                        /* sgg_comm_init handle, or NULL for a single GPU */
  sgg_wa_shard_t shard;              /* row-sharded attention projection (ignored unless comm != NULL and world > 1) */
} sgg_iter_args_t;

int sgg_train_iteration(const sgg_iter_args_t* it, sgg_stream_t stream);

/* ----------------------------------------------------------------------------------------
 * Caller-side front-end helper (gen:29-68): tf.contrib.layers.layer_norm(activation_fn=tf.nn.elu) as used after every
 * convolution of the stack (gen:30,32,36,...): x [B, H*W, C] fp32 (NHWC), moments over all H*W*C elements of a sample
 * (biased variance, eps inside the rsqrt), gamma / beta [C], y = elu(gamma * (x - mean) * rstd + beta).  C must divide
 * 1024 and be a multiple of 4.  stats [B, 2] receives (mean, rstd) for the reverse pass, which recomputes the
 * pre-activation from x (x is the only tensor to keep) and returns dx, dgamma, dbeta (overwritten).
 * scratch: sgg_ln_elu_scratch_floats(B, HW, C) floats, contents private.  HBM-bound: 3 (forward) / 5 (reverse) passes.
 * -------------------------------------------------------------------------------------- */
int64_t sgg_ln_elu_scratch_floats(int64_t B, int64_t HW, int32_t C);
int sgg_ln_elu_forward(const float* x, const float* gamma, const float* beta, int64_t B, int64_t HW, int32_t C, float eps,
                       float* y, float* stats, float* scratch, sgg_stream_t stream);
int sgg_ln_elu_backward(const float* x, const float* dy, const float* gamma, const float* beta, const float* stats,
                        int64_t B, int64_t HW, int32_t C, float* dx, float* dgamma, float* dbeta, float* scratch,
                        sgg_stream_t stream);

/* Test/debug accessor: byte offset of a named intermediate buffer inside the workspace. */
int sgg_ws_lookup(const sgg_dims_t* d, const char* name, int64_t* offset_bytes, int64_t* elem_bytes);

#ifdef __cplusplus
}
#endif
#endif /* SGG_B200_H_ */

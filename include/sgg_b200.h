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
  int32_t seg_klen[SGG_GEMM_MAX_SEG]; /* multiple of 64 */
  /* epilogue */
  float* C; int64_t ldc; int32_t atomic;        /* fp32 output (optional); atomic => red.add */
  void* Chl; int64_t ld_hl; int64_t lo_off;     /* bf16 hi/lo split output (optional) */
  const float* bias;                            /* [N] optional */
  const float* addm; int64_t ld_addm; int32_t add_mod; /* optional broadcast add */
  float alpha;
  int32_t block_n; /* 0 = auto, else 64/128/256 */
  int32_t splits;  /* 0/1 = none; >1 requires atomic */
} sgg_gemm_desc_t;

int sgg_gemm(const sgg_gemm_desc_t* d, sgg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SGG_B200_H_ */

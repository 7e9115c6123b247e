// Unity translation unit of libsgg_b200.so (one nvcc invocation, sm_100a only).
#include "lib.cu"
#include "gemm.cu"
#include "attn.cu"
#include "lstm.cu"
#include "misc.cu"
#include "adamproj.cu"
#include "plan.cu"
#include "comm.cu"

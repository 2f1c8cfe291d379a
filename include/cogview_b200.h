/*
 * cogview_b200.h — C ABI of libcogview_b200.so, the sm_100a implementation of CogView's hot path.
 *
 * The reference (THUDM/CogView) has no FFI: its operator boundary is the Python `mpu` / `model` /
 * `vqvae` namespaces.  Each entry point below replaces the torch/apex/cuBLAS call sequence of one
 * reference function (cited as file:line, relative to the reference tree) and is what a ctypes stub in
 * the reference would bind (INTEGRATION.md shows the stubs).
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless marked host
 *   - functions never allocate or free; the caller owns inputs, outputs and workspaces
 *   - `stream` is a cudaStream_t passed as void*; all work is asynchronous on that stream
 *   - return 0 = ok, < 0 = argument error, > 0 = cudaError_t; cv_last_error() returns the message
 *   - bf16 = __nv_bfloat16 storage; matrices are row-major with a leading dimension in ELEMENTS
 */
#ifndef COGVIEW_B200_H
#define COGVIEW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CV_B200_VERSION 100

int cv_version(void);
const char* cv_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * GEMM  C[M,N] = op(A)[M,K] * op(B)[N,K]^T (+ bias[N]) (+ tanh-GELU)      tcgen05 + TMA + TMEM
 *   replaces F.linear in ColumnParallelLinear.forward (mpu/layers.py:239-249),
 *   RowParallelLinear.forward (mpu/layers.py:312-326), gelu_impl (mpu/sparse_transformer.py:172-176),
 *   the tied-weight logits GEMM (model/gpt2_modeling.py:117-118) and their autograd backward.
 *   a_mn_major = 0: A stored [M,K] (lda >= K);  1: A stored [K,M] (lda >= M)   (wgrad operand)
 *   b_mn_major = 0: B stored [N,K] (ldb >= K);  1: B stored [K,N] (ldb >= N)   (dgrad/wgrad operand)
 *   C: bf16 (c_is_f32 = 0) or fp32 (c_is_f32 = 1); C2 (optional, bf16, same ld): value before GELU
 *   bias: bf16 [N] or NULL;  act: 0 none, 1 tanh-GELU
 *   absmax: NULL or device float (must hold a non-negative value): atomic max of |C| — feeds the
 *           reference's abs-max pre-scaled LayerNorm (mpu/sparse_transformer.py:40-44)
 *   block_n: 0 = auto, or 128 / 256
 * ---------------------------------------------------------------------------------------------- */
int cv_gemm_bf16(const void* A, int a_mn_major, int64_t lda, const void* B, int b_mn_major, int64_t ldb,
                 void* C, int c_is_f32, int64_t ldc, void* C2, const void* bias, int act, float* absmax,
                 int M, int N, int K, int block_n, void* stream);

#ifdef __cplusplus
}
#endif
#endif

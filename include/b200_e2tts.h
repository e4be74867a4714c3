/*
 * b200_e2tts.h — C ABI of libb200e2tts.so: the sm_100a kernels behind the E2-TTS flow-matching hot path.
 *
 * This is the drop-in boundary described in SURVEY.md §8(b): plain pointers and sizes, no torch types.
 * Every entry point replaces an eager PyTorch op chain of the reference (file:line cited per function,
 * relative to /root/reference/e2_tts_pytorch/e2_tts.py; "A.n" = SURVEY.md Appendix A, the unvendored
 * x-transformers / hyper-connections leaves the reference composes).
 *
 * Conventions
 *   - all pointers are DEVICE pointers owned by the caller (inputs, outputs, saved tensors, workspaces);
 *     the library never allocates device memory and keeps no pointer after the call returns;
 *   - every function enqueues work on `stream` and returns immediately: 0 on success, negative on error
 *     (unsupported shape/flag, CUDA launch failure); `b200_last_error()` returns a thread-local message;
 *   - activations are bf16 (row-major, innermost dim contiguous), parameters that feed tensor-core GEMMs
 *     are bf16 packed by `b200_pack_weight`, small vectors / parameter gradients are fp32;
 *   - no host synchronisation, no allocation: every call is CUDA-graph capturable and re-entrant.
 */
#ifndef B200_E2TTS_H
#define B200_E2TTS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* b200_stream_t; /* cudaStream_t */

const char* b200_last_error(void);
int b200_version(void);
/* number of kernels launched by this library since load (per process; used for bench "gpu_launches") */
uint64_t b200_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Tensor-core GEMM (tcgen05 + TMEM + TMA):  D[M,N] = epilogue( sum_k A[m,k] * B[n,k] )
 * Replaces every nn.Linear on the path: to_q/to_k/to_v/to_out (A.4), FeedForward GLU proj + out (A.2),
 * skip_proj :649/:895-896, TextAudioCrossCondition :503-513, proj_in/cond_proj_in :1267-1277, to_pred
 * :1296, and all of their backward contractions (dX = dY*W, dW = dY^T*X).
 *
 * A: bf16. a_mn_major=0: A stored [M,K] (lda = row pitch in elements); 1: A stored [K,M] (M contiguous).
 *    Optional second K source (a_mn_major=0 only): k in [0,K1) from A, k in [K1,K) from A2 — the concat
 *    of skip / cross-condition inputs is never materialised. K1 % 64 == 0.
 * B: bf16. b_mn_major=0: stored [N,K] (an nn.Linear weight); 1: stored [K,N].
 * Row pitches must be multiples of 8 elements (16 B, TMA requirement); M, N, K otherwise arbitrary.
 * Epilogue (applied in this order, each optional):
 *   + bias[n] (fp32)   * colscale[(m / rows_per_batch), n] (fp32, AdaLNZero gate :346-351)
 *   zero rows where rowmask[m]==0 (A.4 step 6)   + resid[m,n] (bf16)
 *   geglu=1: B rows are packed [u(64) | gate(64)] per 128-column tile (b200_pack_weight mode 2);
 *            D2[M,N] <- pre-activation (bf16), D[M,N/2] <- u * gelu_erf(gate) * dropout  (A.2)
 *   split_k>1: fp32 atomic accumulation into D (D is zeroed by the call); d_fp32 must be 1.
 */
typedef struct {
    const void* A; int64_t lda;
    const void* A2; int64_t lda2; int64_t K1;
    const void* B; int64_t ldb;
    int64_t M, N, K;
    int32_t a_mn_major, b_mn_major;
    void* D; int64_t ldd; int32_t d_fp32;
    void* D2; int64_t ldd2;
    const float* bias;
    const float* colscale; int64_t rows_per_batch;
    const uint8_t* rowmask;
    const void* resid; int64_t ldr;
    int32_t geglu; float dropout_p; uint64_t seed;
    int32_t split_k;
} b200_gemm_args;
int b200_gemm(const b200_gemm_args* a, b200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif

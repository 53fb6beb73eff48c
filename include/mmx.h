/* libmmx — C ABI of the B200 (sm_100a) retrieval-scoring path.
 *
 * This header is the drop-in boundary for everything sariekr/multimodal-embedding's main.py does after the
 * encoders return embeddings. The reference has no FFI of its own (it is pure Python); each entry point below
 * names the reference call site it replaces, and multimodal-embedding_b200/ holds the ctypes binding plus the
 * Python functions with the reference's own signatures (INTEGRATION.md shows the stub a maintainer would add).
 *
 * Conventions
 *  - plain C types only; every pointer is a DEVICE pointer unless the name ends in _h;
 *  - every function enqueues on `stream` and returns immediately: 0 = ok, <0 = error, text via mmx_last_error();
 *  - nothing is retained between calls; workspaces are caller-owned;
 *  - matrices are row-major; bf16 is the raw 16-bit pattern (torch.bfloat16);
 *  - there is no CPU fallback: on a device that is not sm_100 every compute entry point fails.
 */
#ifndef MMX_H_
#define MMX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMX_ABI_VERSION 1

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

const char* mmx_last_error(void);
int mmx_abi_version(void);
/* 0 when the current device is compute capability 10.x (B200), error otherwise. */
int mmx_device_check(void);

/* ---------------------------------------------------------------------------------------------------------
 * K2  ColPali late-interaction MaxSim  — replaces main.py:451-470 (chunk loop around processor.score) and
 *     run_benchmark_grand_slam_v28_publication_ready.py:360-386 (T2I and role-swapped I2T).
 *
 *   scores[q, d] = sum_{t in [q_off[q], q_off[q+1])} max_{s in [d_off[d], d_off[d+1])} <q_tok[t,:], d_tok[s,:]>
 *
 * q_tok [n_qtok, dim] and d_tok [n_dtok, dim] are the flat concatenations of the reference's per-item
 * [L_i, 128] tensors; dim must be 128. d_zero (nullable, [nd]) marks documents for which pad_sequence added a
 * zero row inside the reference's 10-document chunk (the padded zero takes part in the max).
 * The role-swapped I2T orientation is the same call with images passed as q_* and captions as d_*.
 * Arithmetic: bf16 x bf16 products accumulated in fp32 on tcgen05, max and token sum in fp32.
 * workspace: mmx_maxsim_workspace_bytes(n_qtok, nd).
 */
int mmx_maxsim_workspace_bytes(int64_t n_qtok, int64_t nd, size_t* bytes);
int mmx_maxsim_scores(const void* q_tok, const int32_t* q_off, int64_t nq, int64_t n_qtok, const void* d_tok,
                      const int32_t* d_off, int64_t nd, int64_t n_dtok, int dim, const uint8_t* d_zero,
                      float* scores, int64_t ld_scores, void* workspace, size_t workspace_bytes,
                      cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MMX_H_ */

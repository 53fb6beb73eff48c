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
 * Documents shorter than 256 rows on average (the role-swapped orientation: captions as documents) take a variant
 * that first re-lays the documents out with lengths rounded up to a multiple of 4 rows (>= 16) by repeating a row of
 * the same document, which cannot change a maximum; the scores are the same expression.
 * workspace: mmx_maxsim_workspace_bytes(n_qtok, nd, n_dtok).
 */
int mmx_maxsim_workspace_bytes(int64_t n_qtok, int64_t nd, int64_t n_dtok, size_t* bytes);
int mmx_maxsim_scores(const void* q_tok, const int32_t* q_off, int64_t nq, int64_t n_qtok, const void* d_tok,
                      const int32_t* d_off, int64_t nd, int64_t n_dtok, int dim, const uint8_t* d_zero,
                      float* scores, int64_t ld_scores, void* workspace, size_t workspace_bytes,
                      cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Ranking passes shared by K1 and KR. Recall@K in the reference is `gt in torch.topk(row, k)` per row
 * (main.py:331-378, v28 :203-229); outside exact fp32 ties that is  #{j : S[q,j] > thr_q} < k  with
 * thr_q = max_{c in GT(q)} S[q,c]  (SURVEY 5.3). Ground truth is CSR: gt_idx[gt_off[q] .. gt_off[q+1]) are GLOBAL
 * gallery column indices; g_base is the global index of this call's column 0 (gallery shard offset).
 *   MMX_MODE_THRESHOLD  thr[q] = max(thr[q], S[q, c - g_base]) over the GT columns inside this shard.
 *                       The caller fills thr with -inf first; shards are combined with a max.
 *   MMX_MODE_COUNT      cnt_above[q] += #{j < ng : S[q,j] > thr[q]}  (caller zeroes cnt_above first);
 *                       above_bits (nullable, [nq, words_per_row] uint32): bit j of word w = S[q, 32w+j] > thr[q].
 */
#define MMX_MODE_THRESHOLD 0
#define MMX_MODE_COUNT 1

/* K1  dense similarity + rank — replaces torch.matmul(txt.float(), img.float().t()) at main.py:473 (v28 :390-391)
 * fused with the metric loops above. Q [nq, d] and G [ng, d] are bf16 row-major (ld = d, d % 8 == 0); products are
 * exact in fp32 and accumulated in fp32 on tcgen05. `scores` (nullable, fp32 [nq, ld_scores]) is only written in
 * MMX_MODE_COUNT and only when the caller needs the matrix itself (compute_similarity_scores drop-in).
 * workspace (MMX_MODE_THRESHOLD only): mmx_dense_workspace_bytes(nq, ng). */
int mmx_dense_workspace_bytes(int64_t nq, int64_t ng, size_t* bytes);
int mmx_dense_rank(const void* Q, int64_t nq, const void* G, int64_t ng, int d, const int32_t* gt_off,
                   const int32_t* gt_idx, int64_t g_base, int mode, float* thr, int32_t* cnt_above,
                   uint32_t* above_bits, int64_t words_per_row, float* scores, int64_t ld_scores, void* workspace,
                   size_t workspace_bytes, cudaStream_t stream);

/* KR  the same two passes over a materialised fp32 score matrix [nq, ng] (row pitch ld) — what
 * compute_metrics_detailed (main.py:305-383) / compute_metrics (v28 :203-229) receive, and what K2 produces. */
int mmx_rank_from_scores(const float* scores, int64_t nq, int64_t ng, int64_t ld, const int32_t* gt_off,
                         const int32_t* gt_idx, int64_t g_base, int mode, float* thr, int32_t* cnt_above,
                         uint32_t* above_bits, int64_t words_per_row, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * KT  per-query top-k lists — the `torch.topk(scores[i], k)` of main.py:333,362,372 (v28 :211,222) itself, for callers
 * that want the retrieved items and for the multi-GPU merge. Order: higher score first, ties by lower column index
 * (torch leaves tied order unspecified, tie class T2). top_s fp32 [nq, k], top_i int32 [nq, k] hold GLOBAL column ids
 * (g_base + local column); rows with fewer than k columns are padded with (-inf, -1). 1 <= k <= 32.
 *   mmx_topk_from_scores  one pass over a materialised fp32 matrix (what K2 produces).
 *   mmx_dense_topk        K1 without the matrix: a group-maximum pass bounds every row's k-th score, a second pass
 *                         collects the few candidates above that bound, a warp picks the k best.
 *                         workspace: mmx_dense_topk_workspace_bytes(nq, ng).
 *   mmx_topk_select       best k of per-row candidate lists cand_s/cand_i [nq, cap] (cand_n nullable [nq] = valid
 *                         entries per row; negative indices are padding).
 *   mmx_merge_topk        gathered per-shard lists, laid out [nq, n_shards * k], -> global top-k.
 */
int mmx_topk_from_scores(const float* scores, int64_t nq, int64_t ng, int64_t ld, int64_t g_base, int k, float* top_s,
                         int32_t* top_i, cudaStream_t stream);
int mmx_dense_topk_workspace_bytes(int64_t nq, int64_t ng, size_t* bytes);
int mmx_dense_topk(const void* Q, int64_t nq, const void* G, int64_t ng, int d, int64_t g_base, int k, float* top_s,
                   int32_t* top_i, void* workspace, size_t workspace_bytes, cudaStream_t stream);
/* mmx_dense_rank (MMX_MODE_COUNT) and the top-k lists in one call: the COUNT pass also stores the group maxima, so the
 * lists cost one extra GEMM pass instead of two. thr = merged thresholds, cnt_above zeroed by the caller. */
int mmx_dense_rank_topk(const void* Q, int64_t nq, const void* G, int64_t ng, int d, const int32_t* gt_off,
                        const int32_t* gt_idx, int64_t g_base, float* thr, int32_t* cnt_above, uint32_t* above_bits,
                        int64_t words_per_row, int k, float* top_s, int32_t* top_i, void* workspace,
                        size_t workspace_bytes, cudaStream_t stream);
int mmx_topk_select(const float* cand_s, const int32_t* cand_i, const int32_t* cand_n, int64_t nq, int cap, int k,
                    float* top_s, int32_t* top_i, cudaStream_t stream);
int mmx_merge_topk(const float* gathered_s, const int32_t* gathered_i, int n_shards, int64_t nq, int k, float* top_s,
                   int32_t* top_i, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * K3  bootstrap — replaces the resample/gather/rescore/metrics loop of main.py:586-634 by the multiplicity
 * identity of SURVEY 5.4 on the BASE score pass:
 *     m = bincount(sample_indices[b]);  c_q = sum_col m[col2unit[col]] * above[q, col];
 *     correct[b, k] = sum_q min(m[q_unit[q]], max(0, k - c_q)),  k in {1, 5, 10}.
 * mmx_bincount:     sample_idx int32 [B, n_draws] (the reference's np.random.choice(N, N) draws, main.py:588)
 *                   -> mult uint8 [B, n_units], saturated at 255 (exact: only min(., 10) is ever used);
 *                   scratch int32 [B, n_units].
 * mmx_above_lists:  first L set columns of every bitmask row, mapped through col2unit (nullable = identity),
 *                   stored transposed: list_t[i * nq + q].
 * mmx_boot_counts:  correct int32 [B, 3]. A row whose list was truncated (cnt_above > L) and ran out before the
 *                   running multiplicity reached 10 is recounted exactly from its bitmask row.
 *                   q_hits1 (nullable int32 [nq], caller-zeroed) += the query's R@1 hits over all iterations: the
 *                   `correct_r1` bits that aggregate_failure_analysis (main.py:669-722) sums per caption feature.
 */
int mmx_bincount(const int32_t* sample_idx, int B, int64_t n_draws, int64_t n_units, int32_t* scratch, uint8_t* mult,
                 cudaStream_t stream);
int mmx_above_lists(const uint32_t* above_bits, int64_t nq, int64_t words_per_row, const int32_t* col2unit, int L,
                    int32_t* list_t, cudaStream_t stream);
int mmx_boot_counts(const int32_t* list_t, int L, const int32_t* cnt_above, const uint32_t* above_bits,
                    int64_t words_per_row, const int32_t* col2unit, const int32_t* q_unit, const uint8_t* mult, int B,
                    int64_t n_units, int64_t nq, int32_t* correct, int32_t* q_hits1, cudaStream_t stream);
/* drawn int64 [B] = sum_q mult[b, q_unit[q]] (q_unit nullable = identity): the query copies iteration b drew — the
 * len(valid_indices) every recall of main.py:358/368/378 is divided by. */
int mmx_copies_drawn(const uint8_t* mult, int B, int64_t n_units, const int32_t* q_unit, int64_t nq, int64_t* drawn,
                     cudaStream_t stream);

/* Shard merge of the above-GT bitmask (the counts and bit columns of main.py:331-378's rank tests are additive /
 * concatenable over gallery shards). gathered uint32 [n_shards, nq, wmax] = every shard's bit columns as ONE all-gather
 * delivers them (bit j of word w of shard s = global column shard_begin_h[s] + 32 w + j; bits beyond a shard's width are
 * ignored); shard_begin_h = HOST array of n_shards + 1 global column offsets (<= 64 shards). Writes the merged bitmask
 * above_bits [nq, words_per_row] (a shard may start inside a 32-column word) and cnt_above [nq] (nullable) = the popcount
 * of every merged row, i.e. the summed strict-greater counts — no second collective. */
int mmx_place_bit_columns(const uint32_t* gathered, int n_shards, int64_t nq, int64_t wmax, const int64_t* shard_begin_h,
                          uint32_t* above_bits, int64_t words_per_row, int32_t* cnt_above, cudaStream_t stream);

/* LIST mode — K1 + K3 for galleries whose [nq, ng] above-GT bitmask does not fit in memory (10^6 x 10^6
 * would be 125 GB of bits). Same identity as above; the above-GT set of a query is kept as a list of at most `cap`
 * columns, which is all K3 ever reads for queries with cnt_above <= cap and enough for min(c_q, 10) otherwise
 * whenever the listed multiplicities already reach 10.
 * mmx_dense_rank_lists: the COUNT pass of mmx_dense_rank (thr = merged thresholds, input) with cnt_above int32 [nq]
 *                   (exact strict-greater count) and, instead of bits, list_i int32 [nq, cap] global column ids of the
 *                   first hits to arrive + list_n int32 [nq] cursor (valid entries = min(list_n, cap)). Both zeroed here.
 * mmx_boot_counts_lists: K3 straight from the per-shard lists list_i [n_shards, nq, cap] + cursors list_n [n_shards, nq]
 *                   (one all-gather when sharded; n_shards = 1 otherwise). Query q's list = its shards' lists in shard
 *                   order, cut at L <= cap entries (col2unit nullable = identity). One warp per query, the
 *                   iterations across the lanes (the multiplicities are transposed into `workspace` first, 128
 *                   iterations per 128-byte row): a list entry costs one coalesced load for 128 iterations.
 *                   (iteration, query) pairs whose truncated list ran out below a multiplicity of 10 are left OUT of
 *                   `correct` and reported as int32 pairs in unres [unres_cap, 2] (any order); *unres_n counts them all
 *                   (call again with more room if it exceeds unres_cap). The caller recounts them from the scores
 *                   (ops.dense_unlisted_resolver); with cap = 64 the probability of one is ~1e-17 per pair.
 *                   workspace: mmx_boot_counts_lists_workspace_bytes(B, n_units) device bytes, 16-byte aligned. */
int mmx_dense_rank_lists(const void* Q, int64_t nq, const void* G, int64_t ng, int d, int64_t g_base, const float* thr,
                         int32_t* cnt_above, int32_t* list_i, int32_t* list_n, int cap, cudaStream_t stream);
int mmx_boot_counts_lists_workspace_bytes(int B, int64_t n_units, size_t* bytes);
int mmx_boot_counts_lists(const int32_t* list_i, const int32_t* list_n, int n_shards, int64_t nq, int cap,
                          const int32_t* col2unit, int L, const int32_t* cnt_above, const int32_t* q_unit,
                          const uint8_t* mult, int B, int64_t n_units, int32_t* correct, int32_t* q_hits1,
                          int32_t* unres, int32_t* unres_n, int unres_cap, void* workspace, size_t workspace_bytes,
                          cudaStream_t stream);

/* K4  CI resample means — replaces the 10 000 x (np.random.choice(data, n) ; np.mean) loop of main.py:268-270.
 * idx int32 [R, n] is the reference's index stream (RandomState.randint(0, n, (R, n)) == R x choice);
 * means[r] = np.mean(data[idx[r]]) bit for bit (fp64, numpy's pairwise summation order). */
int mmx_ci_means(const double* data, int n, const int32_t* idx, int R, double* means, cudaStream_t stream);
/* The same for `batch` independent data vectors in one launch (the nine metric keys of a CSV row, main.py:644-651):
 * data [batch, n], idx [batch, R, n], means [batch, R]. */
int mmx_ci_means_batched(const double* data, int n, const int32_t* idx, int R, int batch, double* means,
                         cudaStream_t stream);

/* K5  permutation statistics — replaces the shuffle + two means loop of main.py:295-299.
 * perm int32 [P, n1+n2] are the cumulative np.random.shuffle states of an index array;
 * diffs[p] = |np.mean(combined[perm[p, :n1]]) - np.mean(combined[perm[p, n1:]])| bit for bit. */
int mmx_perm_stats(const double* combined, int n1, int n2, const int32_t* perm, int P, double* diffs,
                   cudaStream_t stream);

/* K6  numpy legacy MT19937 stream on the device — the draws behind np.random.choice(n, n) (main.py:588) and
 * np.random.choice(data, n) (main.py:269), both == RandomState.randint(0, high, count):
 * masked rejection on 32-bit MT19937 outputs. state: uint32 [625] = numpy's get_state() key[624] followed by pos;
 * it is advanced exactly as numpy would advance it, so it can be written back with set_state().
 * out int32 [count] = the next `count` values numpy would return for randint(0, high). 1 <= high < 2^31. */
int mmx_mt19937_randint(uint32_t* state, uint32_t high, int64_t count, int32_t* out, cudaStream_t stream);

/* K6p  the same contract, produced by all SMs at once through MT19937 jump-ahead. polys uint32 [n_poly, 624] holds
 * x^(k * seg_blocks * 624) mod phi(x), k = 1..n_poly (phi = MT19937's characteristic polynomial; the table is built by
 * multimodal-embedding_b200/mtjump.py): segment k of the stream starts from a GF(2) convolution of the first 33 blocks
 * with polys[k-1], every CTA regenerates one segment, a count pass and a write pass keep numpy's order under
 * rejection sampling. Requests beyond one launch's budget are chained; a request that outruns the 2 % slack of its
 * segment budget is finished by the sequential kernel, so the result never depends on the budget.
 * scratch: mmx_mt19937_par_scratch_bytes(n_poly). */
int mmx_mt19937_par_scratch_bytes(int n_poly, size_t* bytes);
/* Optional: positions of the set bits of every polynomial of `polys` (uint16 [n_poly][ld], ld % 4 == 0, rows padded with
 * the value 20592) and their counts (int32 [n_poly]), all device pointers kept by reference; lets the jump convolution
 * gather without scanning bits. One registration per table pointer (up to 16: short and long tables, several devices);
 * pass NULL positions to unregister a table. Results never depend on it. */
int mmx_mt19937_par_set_bit_positions(const uint32_t* polys, const uint16_t* bitpos, const int32_t* bitcnt, int ld);
/* Test hook: budget = count / acceptance * rel + blocks * 624 raw words (default 1.02, 64). Results never depend on it. */
int mmx_mt19937_par_set_slack(double rel, double blocks);
int mmx_mt19937_randint_par(uint32_t* state, uint32_t high, int64_t count, int32_t* out, const uint32_t* polys,
                            int n_poly, int seg_blocks, void* scratch, size_t scratch_bytes, cudaStream_t stream);
/* The same request, of which only the values [win_begin, win_begin + win_count) are stored (out int32 [win_count],
 * nullable when win_count == 0): the generator advances by all `count` values, but segments that hold nothing of the
 * window skip their write pass. A rank that bootstraps iterations [b0, b1) of a sharded run (main.py:586-588: one
 * np.random.choice(n, n) per iteration) draws the whole stream's positions and keeps its own rows only; win_count == 0
 * skips a request (ops.ShufflePrefetch runs ahead of the models' draws that way). */
int mmx_mt19937_randint_par_window(uint32_t* state, uint32_t high, int64_t count, int32_t* out, int64_t win_begin,
                                   int64_t win_count, const uint32_t* polys, int n_poly, int seg_blocks, void* scratch,
                                   size_t scratch_bytes, cudaStream_t stream);

/* The next `count` raw 32-bit outputs of the stream (genrand_int32), produced like K6p; out nullable = only advance
 * the generator state by `count` words. */
int mmx_mt19937_raw_par(uint32_t* state, int64_t count, uint32_t* out, const uint32_t* polys, int n_poly,
                        int seg_blocks, void* scratch, size_t scratch_bytes, cudaStream_t stream);

/* K6s  numpy's cumulative in-place shuffles — the `np.random.shuffle(combined)` of main.py:296, P times on the same
 * array — replayed on the device from raw stream words (legacy RandomState.shuffle: Fisher-Yates from the top,
 * j = random_interval(i) by masked rejection). words uint32 [n_words] = mmx_mt19937_raw_par output from the current
 * state; perm int32 [P, n] = the index array after every shuffle (what mmx_perm_stats takes);
 * info int64 [2] (device): info[1] = number of words the P shuffles consumed, or -1 when n_words was too small (perm
 * is then unspecified; call again with more words; -3: a mask epoch did not fit the 65 536-word window, which needs a
 * > 100-sigma run of rejections). 2 <= n <= 65536: up to 16 384 elements the table scheme (work ~ n^2 * P, all SMs),
 * beyond it one CTA that is parallel within a permutation (bracketing prefix sums per mask epoch).
 * workspace: mmx_perm_shuffle_workspace_bytes(n, P, n_words). */
int mmx_perm_shuffle_workspace_bytes(int n, int P, int64_t n_words, size_t* bytes);
int mmx_perm_shuffle(const uint32_t* words, int64_t n_words, int n, int P, int32_t* perm, int64_t* info,
                     void* workspace, size_t workspace_bytes, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MMX_H_ */

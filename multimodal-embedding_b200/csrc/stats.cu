// K3-K5 — the statistics stage: bootstrap resample counts, CI resample means, permutation-test statistics.
//
// K3 replaces the 1000-iteration resample / gather / rescore / topk loop of /root/reference/main.py:586-634 by the
//    multiplicity identity:  c_q = sum_j m[j] * [S[q,j] > thr_q],  correct_k = sum_q min(m[q], max(0,k-c_q)).
//    Only min(c_q, 10) matters, so each (iteration, query) walks the query's above-GT list until the running
//    multiplicity reaches 10; a truncated list that runs out first falls back to the exact bitmask row.
// K4 replaces the 10 000 x (np.random.choice + np.mean) loop of main.py:268-270: gathers + fp64 sums in numpy's
//    pairwise order, so the means (and the percentiles the host takes of them) are bit-identical.
// K5 replaces the 10 000 x (shuffle + two means) loop of main.py:295-299 on host-generated permutations.
// All three are byte/integer/fp64 streaming work bound by memory traffic, not tensor-core work.
#include <vector>

#include "../../include/mmx.h"
#include "mmx_host.cuh"

namespace mmx {
namespace stats {

// ------------------------------------------------------------------------------------------- multiplicities
__global__ void bincount_kernel(const int32_t* __restrict__ idx, int B, long long n_draws, long long n_units,
                                int32_t* __restrict__ counts) {
    const int b = blockIdx.y;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_draws;
         i += (long long)gridDim.x * blockDim.x) {
        const int u = idx[(long long)b * n_draws + i];
        if (u >= 0 && u < n_units) atomicAdd(&counts[(long long)b * n_units + u], 1);
    }
}
__global__ void saturate_u8_kernel(const int32_t* __restrict__ counts, long long n, uint8_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint8_t)min(counts[i], 255);
}

// ------------------------------------------------------------------------------------------- above-GT lists
// list_t[i * nq + q] = unit (col2unit applied) of the i-th set bit of row q, for i < min(cnt, L). One warp per row.
__global__ void above_lists_kernel(const uint32_t* __restrict__ bits, long long nq, long long wpr,
                                   const int32_t* __restrict__ col2unit, int L, int32_t* __restrict__ list_t) {
    const int lane = threadIdx.x & 31;
    const long long q = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    int found = 0;
    for (long long w0 = 0; w0 < wpr && found < L; w0 += 32) {
        const long long w = w0 + lane;
        const uint32_t word = w < wpr ? bits[q * wpr + w] : 0u;
        const int c = __popc(word);
        int incl = c;  // inclusive prefix over lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        int pos = found + incl - c;
        uint32_t x = word;
        while (x != 0u && pos < L) {
            const int bit = __ffs(x) - 1;
            x &= x - 1;
            const long long col = (w << 5) + bit;
            list_t[(long long)pos * nq + q] = col2unit != nullptr ? col2unit[col] : (int32_t)col;
            ++pos;
        }
        found += __shfl_sync(0xffffffffu, incl, 31);
    }
}

// LIST mode (K1 mmx_dense_rank_lists): per-shard column lists [n_shards, nq, cap] + cursors [n_shards, nq] ->
// list_t[i * nq + q] = unit of the i-th entry of query q's concatenated lists (shard order), i < L; total[q] = number
// of entries available (sum of min(cursor, cap) over shards, unclamped by L).
__global__ void lists_to_units_kernel(const int32_t* __restrict__ list_i, const int32_t* __restrict__ list_n,
                                      int n_shards, long long nq, int cap, const int32_t* __restrict__ col2unit, int L,
                                      int32_t* __restrict__ list_t, int32_t* __restrict__ avail) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    int pos = 0, tot = 0;
    for (int s = 0; s < n_shards; ++s) {
        const int n = min(list_n[(long long)s * nq + q], cap);
        tot += n;
        const int32_t* src = list_i + ((long long)s * nq + q) * cap;
        for (int i = 0; i < n && pos < L; ++i, ++pos) {
            const int32_t col = src[i];
            list_t[(long long)pos * nq + q] = col2unit != nullptr ? col2unit[col] : col;
        }
    }
    avail[q] = tot;
}

// One block per bootstrap iteration b. mult[b,:] (uint8, saturated) sits in shared memory when it fits.
__global__ void __launch_bounds__(256)
boot_counts_kernel(const int32_t* __restrict__ list_t, int L, const int32_t* __restrict__ cnt_above,
                   const uint32_t* __restrict__ bits, long long wpr, const int32_t* __restrict__ col2unit,
                   const int32_t* __restrict__ q_unit, const uint8_t* __restrict__ mult, long long n_units,
                   long long nq, int mult_in_smem, int32_t* __restrict__ correct, int32_t* __restrict__ q_hits1,
                   int32_t* __restrict__ unres, int32_t* __restrict__ unres_n, int unres_cap) {
    extern __shared__ uint8_t sm_mult[];
    __shared__ int red[3][8];
    const int b = blockIdx.x;
    const uint8_t* mrow = mult + (long long)b * n_units;
    if (mult_in_smem) {
        for (long long i = threadIdx.x; i < n_units; i += blockDim.x) sm_mult[i] = mrow[i];
        __syncthreads();
        mrow = sm_mult;
    }
    int c1 = 0, c5 = 0, c10 = 0;
    for (long long q = threadIdx.x; q < nq; q += blockDim.x) {
        const int mq = mrow[q_unit != nullptr ? q_unit[q] : q];
        if (mq == 0) continue;  // the query was not drawn in this iteration
        const int cnt = cnt_above[q];
        const int len = min(cnt, L);
        int c = 0;
        for (int i = 0; i < len && c < 10; ++i) c += mrow[list_t[(long long)i * nq + q]];
        if (c < 10 && cnt > L && bits == nullptr) {
            // LIST mode: the truncated list ran out before saturating and there is no bitmask to fall back to: leave
            // this (iteration, query) out of the sums and report it; the caller recounts it exactly from the scores
            const int slot = atomicAdd(unres_n, 1);
            if (slot < unres_cap) {
                unres[2 * slot] = b;
                unres[2 * slot + 1] = (int32_t)q;
            }
            continue;
        }
        if (c < 10 && cnt > L) {
            // truncated list ran out before saturating: exact walk over the bitmask row (rare)
            c = 0;
            for (long long w = 0; w < wpr && c < 10; ++w) {
                uint32_t x = bits[q * wpr + w];
                while (x != 0u && c < 10) {
                    const int bit = __ffs(x) - 1;
                    x &= x - 1;
                    const long long col = (w << 5) + bit;
                    c += mrow[col2unit != nullptr ? col2unit[col] : col];
                }
            }
        }
        const int h1 = min(mq, max(0, 1 - c));
        if (q_hits1 != nullptr && h1 > 0) atomicAdd(q_hits1 + q, h1);  // per-query R@1 hits over all iterations
        c1 += h1;
        c5 += min(mq, max(0, 5 - c));
        c10 += min(mq, max(0, 10 - c));
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c1 += __shfl_xor_sync(0xffffffffu, c1, o);
        c5 += __shfl_xor_sync(0xffffffffu, c5, o);
        c10 += __shfl_xor_sync(0xffffffffu, c10, o);
    }
    if (lane == 0) {
        red[0][warp] = c1;
        red[1][warp] = c5;
        red[2][warp] = c10;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        int s = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[threadIdx.x][w];
        correct[b * 3 + threadIdx.x] = s;
    }
}

// ------------------------------------------------------------------------------------------- numpy pairwise sums
// The recursion tree of numpy's pairwise sum for a given n, flattened on the host: leaves (<= 128 elements, eight
// interleaved accumulators) in left-to-right order, each followed by the number of stack merges it completes.
constexpr int MAX_LEAVES = 320;
struct Program {
    int n_leaves;
    uint16_t size[MAX_LEAVES];
    uint8_t merges[MAX_LEAVES];
};

static void gen_program(long long n, std::vector<std::pair<int, int>>& out) {
    if (n <= 128) {
        out.push_back({(int)n, 0});
        return;
    }
    long long n2 = n / 2;
    n2 -= n2 % 8;
    gen_program(n2, out);
    gen_program(n - n2, out);
    out.back().second += 1;
}
static int make_program(long long n, Program* p) {
    std::vector<std::pair<int, int>> v;
    gen_program(n, v);
    MMX_REQUIRE((int)v.size() <= MAX_LEAVES, "vector too long for the pairwise-sum program (n=%lld)", n);
    p->n_leaves = (int)v.size();
    for (size_t i = 0; i < v.size(); ++i) {
        p->size[i] = (uint16_t)v[i].first;
        p->merges[i] = (uint8_t)v[i].second;
    }
    return 0;
}

__device__ __forceinline__ double pairwise_gather(const double* __restrict__ data, const int32_t* __restrict__ idx,
                                                  const Program& prog) {
    double stack[24];
    int sp = 0;
    long long pos = 0;
    for (int l = 0; l < prog.n_leaves; ++l) {
        const int n = prog.size[l];
        const int32_t* ix = idx + pos;
        double res;
        if (n < 8) {
            res = 0.0;
            for (int i = 0; i < n; ++i) res += data[ix[i]];
        } else {
            double r0 = data[ix[0]], r1 = data[ix[1]], r2 = data[ix[2]], r3 = data[ix[3]];
            double r4 = data[ix[4]], r5 = data[ix[5]], r6 = data[ix[6]], r7 = data[ix[7]];
            int i = 8;
            for (; i < n - (n % 8); i += 8) {
                r0 += data[ix[i]];
                r1 += data[ix[i + 1]];
                r2 += data[ix[i + 2]];
                r3 += data[ix[i + 3]];
                r4 += data[ix[i + 4]];
                r5 += data[ix[i + 5]];
                r6 += data[ix[i + 6]];
                r7 += data[ix[i + 7]];
            }
            res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
            for (; i < n; ++i) res += data[ix[i]];
        }
        pos += n;
        stack[sp++] = res;
        for (int m = 0; m < prog.merges[l]; ++m) {
            --sp;
            stack[sp - 1] = stack[sp - 1] + stack[sp];
        }
    }
    return stack[0];
}

// out[r] = mean(data[idx[r, 0:n1]])                       (n2 == 0, K4)
// out[r] = | mean(data[idx[r, 0:n1]]) - mean(data[idx[r, n1:n1+n2]]) |   (n2 > 0, K5)
__global__ void __launch_bounds__(128)
gather_mean_kernel(const double* __restrict__ data, int n_data, const int32_t* __restrict__ idx, long long ld, int R,
                   int n1, int n2, const __grid_constant__ Program p1, const __grid_constant__ Program p2,
                   int data_in_smem, double* __restrict__ out) {
    extern __shared__ double sm_data[];
    // blockIdx.y = batch entry (its own data vector, index rows and outputs): nine CIs of a row in one launch
    data += (long long)blockIdx.y * n_data;
    idx += (long long)blockIdx.y * R * ld;
    out += (long long)blockIdx.y * R;
    const double* d = data;
    if (data_in_smem) {
        for (int i = threadIdx.x; i < n_data; i += blockDim.x) sm_data[i] = data[i];
        __syncthreads();
        d = sm_data;
    }
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const int32_t* row = idx + (long long)r * ld;
    const double m1 = pairwise_gather(d, row, p1) / (double)n1;
    if (n2 > 0) {
        const double m2 = pairwise_gather(d, row + n1, p2) / (double)n2;
        out[r] = fabs(m1 - m2);
    } else {
        out[r] = m1;
    }
}

}  // namespace stats
}  // namespace mmx

using namespace mmx;
using namespace mmx::stats;

static const char boot_tag = 0;  // boot_counts_kernel's shared-memory attribute: set once per device

extern "C" int mmx_bincount(const int32_t* sample_idx, int B, int64_t n_draws, int64_t n_units, int32_t* scratch,
                            uint8_t* mult, cudaStream_t stream) {
    MMX_REQUIRE(B >= 0 && n_draws >= 0 && n_units > 0, "bad shape");
    if (B == 0) return 0;
    MMX_CHECK_CUDA(cudaMemsetAsync(scratch, 0, (size_t)B * n_units * sizeof(int32_t), stream));
    if (n_draws > 0) {
        dim3 grid((unsigned)std::min<long long>((n_draws + 255) / 256, 64), (unsigned)B);
        bincount_kernel<<<grid, 256, 0, stream>>>(sample_idx, B, n_draws, n_units, scratch);
        MMX_CHECK_CUDA(cudaGetLastError());
    }
    const long long n = (long long)B * n_units;
    saturate_u8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(scratch, n, mult);
    MMX_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mmx_above_lists(const uint32_t* above_bits, int64_t nq, int64_t words_per_row, const int32_t* col2unit,
                               int L, int32_t* list_t, cudaStream_t stream) {
    MMX_REQUIRE(L >= 1 && L <= 1024, "L out of range");
    if (nq == 0) return 0;
    above_lists_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, stream>>>(above_bits, nq, words_per_row, col2unit, L, list_t);
    MMX_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mmx_boot_counts(const int32_t* list_t, int L, const int32_t* cnt_above, const uint32_t* above_bits,
                               int64_t words_per_row, const int32_t* col2unit, const int32_t* q_unit,
                               const uint8_t* mult, int B, int64_t n_units, int64_t nq, int32_t* correct,
                               int32_t* q_hits1, cudaStream_t stream) {
    MMX_REQUIRE(list_t && cnt_above && above_bits && mult && correct, "NULL argument");
    if (B == 0) return 0;
    const int in_smem = n_units <= 200 * 1024;
    const size_t smem = in_smem ? (size_t)((n_units + 15) & ~15ll) : 0;
    if (smem > 48 * 1024 && first_use_on_device(&boot_tag))
        MMX_CHECK_CUDA(cudaFuncSetAttribute(boot_counts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 16));
    boot_counts_kernel<<<B, 256, smem, stream>>>(list_t, L, cnt_above, above_bits, words_per_row, col2unit, q_unit, mult,
                                                 n_units, nq, in_smem, correct, q_hits1, nullptr, nullptr, 0);
    MMX_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mmx_lists_to_units(const int32_t* list_i, const int32_t* list_n, int n_shards, int64_t nq, int cap,
                                  const int32_t* col2unit, int L, int32_t* list_t, int32_t* avail, cudaStream_t stream) {
    MMX_REQUIRE(list_i && list_n && list_t && avail, "NULL argument");
    MMX_REQUIRE(n_shards >= 1 && cap >= 1 && L >= 1 && L <= 1024, "bad list geometry");
    if (nq == 0) return 0;
    lists_to_units_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, stream>>>(list_i, list_n, n_shards, nq, cap, col2unit,
                                                                           L, list_t, avail);
    MMX_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// K3 without a bitmask (LIST mode). list_t holds min(cnt_above[q], L) units of query q's above-GT set — all of it
// when cnt_above[q] <= L. (iteration, query) pairs whose truncated list did not reach a multiplicity of 10 are left
// out of `correct` and reported in unres [unres_cap, 2] / *unres_n (may exceed unres_cap: call again with more room).
extern "C" int mmx_boot_counts_lists(const int32_t* list_t, int L, const int32_t* cnt_above, const int32_t* q_unit,
                                     const uint8_t* mult, int B, int64_t n_units, int64_t nq, int32_t* correct,
                                     int32_t* q_hits1, int32_t* unres, int32_t* unres_n, int unres_cap,
                                     cudaStream_t stream) {
    MMX_REQUIRE(list_t && cnt_above && mult && correct && unres && unres_n && unres_cap >= 1, "NULL argument");
    MMX_CHECK_CUDA(cudaMemsetAsync(unres_n, 0, sizeof(int32_t), stream));
    if (B == 0) return 0;
    const int in_smem = n_units <= 200 * 1024;
    const size_t smem = in_smem ? (size_t)((n_units + 15) & ~15ll) : 0;
    if (smem > 48 * 1024 && first_use_on_device(&boot_tag))
        MMX_CHECK_CUDA(cudaFuncSetAttribute(boot_counts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 16));
    boot_counts_kernel<<<B, 256, smem, stream>>>(list_t, L, cnt_above, nullptr, 0, nullptr, q_unit, mult, n_units, nq,
                                                 in_smem, correct, q_hits1, unres, unres_n, unres_cap);
    MMX_CHECK_CUDA(cudaGetLastError());
    return 0;
}

static int launch_gather_mean(const double* data, int n_data, const int32_t* idx, int64_t ld, int R, int n1, int n2,
                              double* out, cudaStream_t stream, int batch = 1) {
    MMX_REQUIRE(n1 >= 1 && n2 >= 0 && ld >= (int64_t)n1 + n2, "bad segment sizes");
    MMX_REQUIRE(batch >= 1 && batch <= 65535, "batch out of range");
    if (R == 0) return 0;
    Program p1, p2;
    if (make_program(n1, &p1) != 0) return -2;
    p2.n_leaves = 0;
    if (n2 > 0 && make_program(n2, &p2) != 0) return -2;
    const int in_smem = (size_t)n_data * 8 <= 200 * 1024;
    const size_t smem = in_smem ? (size_t)n_data * 8 : 0;
    static const char gather_tag = 0;
    if (smem > 48 * 1024 && first_use_on_device(&gather_tag))
        MMX_CHECK_CUDA(cudaFuncSetAttribute(gather_mean_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    gather_mean_kernel<<<dim3((R + 127) / 128, batch), 128, smem, stream>>>(data, n_data, idx, ld, R, n1, n2, p1, p2,
                                                                            in_smem, out);
    MMX_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mmx_ci_means(const double* data, int n, const int32_t* idx, int R, double* means, cudaStream_t stream) {
    return launch_gather_mean(data, n, idx, n, R, n, 0, means, stream);
}

extern "C" int mmx_ci_means_batched(const double* data, int n, const int32_t* idx, int R, int batch, double* means,
                                    cudaStream_t stream) {
    return launch_gather_mean(data, n, idx, n, R, n, 0, means, stream, batch);
}

extern "C" int mmx_perm_stats(const double* combined, int n1, int n2, const int32_t* perm, int P, double* diffs,
                              cudaStream_t stream) {
    MMX_REQUIRE(n2 >= 1, "n2 must be >= 1");
    return launch_gather_mean(combined, n1 + n2, perm, (int64_t)n1 + n2, P, n1, n2, diffs, stream);
}

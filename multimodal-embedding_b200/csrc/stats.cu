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

// drawn[b] = sum_q mult[b, q_unit[q]]: the query copies iteration b drew (the denominator of main.py:358's recall).
// One CTA per iteration; the iteration's multiplicity row (n_units bytes) stays in L1/L2 while the queries stream by.
__global__ void __launch_bounds__(256)
copies_drawn_kernel(const uint8_t* __restrict__ mult, long long n_units, const int32_t* __restrict__ q_unit,
                    long long nq, long long* __restrict__ drawn) {
    __shared__ long long red[8];
    const uint8_t* mrow = mult + (long long)blockIdx.x * n_units;
    long long sum = 0;
    for (long long q = threadIdx.x; q < nq; q += blockDim.x) sum += mrow[q_unit != nullptr ? q_unit[q] : q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long tot = 0;
        for (int w = 0; w < 8; ++w) tot += red[w];
        drawn[blockIdx.x] = tot;
    }
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

// LIST mode (K1 mmx_dense_rank_lists): per-shard column lists [n_shards, nq, cap] + cursors [n_shards, nq]. Query q's
// list is the concatenation of its shards' lists in shard order, cut at L entries.
//
// The walk is laid out the other way round from boot_counts_kernel: ONE WARP PER QUERY, the bootstrap iterations across
// the lanes. The multiplicities are transposed first (mult_t[chunk][unit][128] uint8: the 128 iterations of a chunk of one
// unit are one 128-byte line), so a list entry costs the warp ONE coalesced 128-byte load that serves 128 iterations, the
// lists are read once per chunk instead of once per iteration, and up to LOOKAHEAD loads are in flight per warp (the
// addresses come from the list, not from the running sums). With one thread per (iteration, query) a 10^6-query,
// 125-iteration block re-read the 256 MB list array 125 times with one random byte gather per entry: 54 ms; this
// layout: see profiles/.
constexpr int CHUNK_B = 128;     // iterations per chunk = bytes per mult_t row
constexpr int LOOKAHEAD = 8;

// mult [B, n_units] -> mult_t [n_chunks][n_units][128] (iterations beyond B: 0 = "query not drawn", which the walk skips)
__global__ void __launch_bounds__(256)
transpose_mult_kernel(const uint8_t* __restrict__ mult, int B, long long n_units, uint8_t* __restrict__ mult_t) {
    __shared__ uint8_t tile[CHUNK_B][CHUNK_B + 4];
    const int chunk = blockIdx.y;
    const long long u0 = (long long)blockIdx.x * CHUNK_B;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int it = warp; it < CHUNK_B; it += 8) {
        const int b = chunk * CHUNK_B + it;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long u = u0 + k * 32 + lane;
            tile[it][k * 32 + lane] = (b < B && u < n_units) ? mult[(long long)b * n_units + u] : (uint8_t)0;
        }
    }
    __syncthreads();
    uint32_t* dst = reinterpret_cast<uint32_t*>(mult_t + ((long long)chunk * n_units + u0) * CHUNK_B);
    for (int u = warp; u < CHUNK_B && u0 + u < n_units; u += 8) {
        const uint32_t v = (uint32_t)tile[4 * lane][u] | ((uint32_t)tile[4 * lane + 1][u] << 8) |
                           ((uint32_t)tile[4 * lane + 2][u] << 16) | ((uint32_t)tile[4 * lane + 3][u] << 24);
        dst[u * 32 + lane] = v;
    }
}

__device__ __forceinline__ uint32_t bytes_ge(uint32_t a, uint32_t b) { return __vcmpgeu4(a, b); }  // 0xff per byte a >= b

__global__ void __launch_bounds__(256)
boot_counts_lists_kernel(const int32_t* __restrict__ list_i, const int32_t* __restrict__ list_n, int n_shards,
                         long long nq, int cap, const int32_t* __restrict__ col2unit, int L,
                         const int32_t* __restrict__ cnt_above, const int32_t* __restrict__ q_unit,
                         const uint8_t* __restrict__ mult_t, int B, long long n_units, int32_t* __restrict__ correct,
                         int32_t* __restrict__ q_hits1, int32_t* __restrict__ unres, int32_t* __restrict__ unres_n,
                         int unres_cap) {
    __shared__ int red[8][3][CHUNK_B];
    const int chunk = blockIdx.y, b0 = chunk * CHUNK_B;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t* mt = reinterpret_cast<const uint32_t*>(mult_t + (long long)chunk * n_units * CHUNK_B);
    int acc[3][4];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[k][j] = 0;
    const long long n_warps = (long long)gridDim.x * 8;
    for (long long q = (long long)blockIdx.x * 8 + warp; q < nq; q += n_warps) {
        const long long uq = q_unit != nullptr ? q_unit[q] : q;
        const uint32_t mq4 = mt[uq * 32 + lane];                       // this lane's four iterations' copies of the query
        const uint32_t drawn = ~bytes_ge(0u, mq4);                     // 0xff where the query was drawn (mq > 0)
        if (__ballot_sync(0xffffffffu, drawn != 0u) == 0u) continue;
        const int cnt = cnt_above[q];
        const int len = min(cnt, L);
        uint32_t c4 = 0;                                               // four running sums, saturating bytes
        bool done = __ballot_sync(0xffffffffu, (drawn & ~bytes_ge(c4, 0x0a0a0a0au)) != 0u) == 0u;
        int pos = 0;
        for (int s = 0; s < n_shards && pos < len && !done; ++s) {
            const int take = min(min(list_n[(long long)s * nq + q], cap), len - pos);
            const int32_t* src = list_i + ((long long)s * nq + q) * cap;
            for (int base = 0; base < take && !done; base += 32) {
                const int m = min(32, take - base);
                long long unit = 0;
                if (lane < m) {
                    const int32_t col = src[base + lane];
                    unit = col2unit != nullptr ? col2unit[col] : col;
                }
                for (int i0 = 0; i0 < m && !done; i0 += LOOKAHEAD) {
                    uint32_t v[LOOKAHEAD];
#pragma unroll
                    for (int j = 0; j < LOOKAHEAD; ++j) {
                        const long long u = __shfl_sync(0xffffffffu, unit, (i0 + j) & 31);
                        v[j] = (i0 + j < m) ? mt[u * 32 + lane] : 0u;
                    }
#pragma unroll
                    for (int j = 0; j < LOOKAHEAD; ++j) c4 = __vaddus4(c4, v[j]);
                    // every drawn iteration of every lane at 10 or more: the rest of the list cannot change anything
                    done = __ballot_sync(0xffffffffu, (drawn & ~bytes_ge(c4, 0x0a0a0a0au)) != 0u) == 0u;
                }
            }
            pos += take;
        }
        int h1_sum = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int mq = (int)((mq4 >> (8 * j)) & 0xffu);
            const int c = (int)((c4 >> (8 * j)) & 0xffu);
            if (mq == 0) continue;
            if (c < 10 && cnt > L) {
                // the truncated list ran out before saturating: leave this (iteration, query) out of the sums and report it;
                // the caller recounts it exactly from the scores
                const int slot = atomicAdd(unres_n, 1);
                if (slot < unres_cap) {
                    unres[2 * slot] = b0 + 4 * lane + j;
                    unres[2 * slot + 1] = (int32_t)q;
                }
                continue;
            }
            const int h1 = min(mq, max(0, 1 - c));
            h1_sum += h1;
            acc[0][j] += h1;
            acc[1][j] += min(mq, max(0, 5 - c));
            acc[2][j] += min(mq, max(0, 10 - c));
        }
        if (q_hits1 != nullptr) {                                      // per-query R@1 hits over all iterations
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) h1_sum += __shfl_xor_sync(0xffffffffu, h1_sum, o);
            if (lane == 0 && h1_sum > 0) atomicAdd(q_hits1 + q, h1_sum);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) red[warp][k][4 * lane + j] = acc[k][j];
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * CHUNK_B; i += blockDim.x) {
        const int k = i / CHUNK_B, it = i % CHUNK_B;
        int sum = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += red[w][k][it];
        if (b0 + it < B && sum != 0) atomicAdd(correct + (long long)(b0 + it) * 3 + k, sum);
    }
}

// One block per bootstrap iteration b. mult[b,:] (uint8, saturated) sits in shared memory when it fits.
__global__ void __launch_bounds__(256)
boot_counts_kernel(const int32_t* __restrict__ list_t, int L, const int32_t* __restrict__ cnt_above,
                   const uint32_t* __restrict__ bits, long long wpr, const int32_t* __restrict__ col2unit,
                   const int32_t* __restrict__ q_unit, const uint8_t* __restrict__ mult, long long n_units,
                   long long nq, int mult_in_smem, int32_t* __restrict__ correct, int32_t* __restrict__ q_hits1) {
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
// numpy's add.reduce over a contiguous fp64 vector is a pairwise sum (numpy/core/src/umath/loops_utils.h.src):
//   n < 8: sequential from 0.0;  n <= 128: eight interleaved accumulators r[i % 8], then ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)),
//   then the n % 8 leftovers sequentially;  n > 128: split at n2 = (n / 2) & ~7 and add the two halves.
// The recursion tree for a given n is flattened into its leaves in left-to-right order, each with the number of
// stack merges it completes; every CTA builds that list once in shared memory (one thread, an explicit stack — a few
// hundred leaves at most), so there is no limit on n. All leaves but the last have a multiple of 8 elements.
//
// gather_mean_kernel: G threads per resample row, 128 / G rows per CTA. The G = 2^d subtrees at depth d of the
// recursion are independent pairwise sums over consecutive column ranges; each thread walks ONE subtree in numpy's
// order and the G partial sums are added pairwise in the recursion's own order (shuffles), which is numpy's result bit
// for bit. More threads per row = more rows' worth of memory requests in flight: with one thread per row a
// [3, 10 000, 1000] CI batch was 30 000 threads for 148 SMs. The index columns of the 128 (row, subtree) pairs of a CTA
// are streamed through a 2-stage cp.async ring of [128 x 32] tiles (16-byte copies, 8 consecutive threads fetch 128
// contiguous bytes; the 16-byte chunks of a tile row are XOR-swizzled so that every thread reads its own tile row with
// conflict-free LDS.128), so the global traffic is fully coalesced although every thread walks its own column range.
constexpr int GR = 128;      // (row, subtree) pairs per CTA = threads
constexpr int GC = 32;       // columns per tile
constexpr int GST = 2;       // tiles in the ring (one in flight while one is consumed; 5 CTAs per SM)
constexpr int GMAX = 32;     // subtrees per row at most (one warp adds them up)
constexpr int DATA_SMEM_MAX = 96 * 1024;

// room for the leaf programs of the G subtrees of a vector of n: every subtree gets the same stride
__host__ __device__ __forceinline__ int prog_stride(long long n, int lg) { return (int)(((n >> lg) + 8) / 56 + 3); }
__host__ __device__ __forceinline__ int leaves_cap(long long n, int lg) { return prog_stride(n, lg) << lg; }

// depth to which pairwise_sum(n) splits EVERY node (all nodes above that depth are longer than 128)
static int full_split_depth(long long n) {
    std::vector<long long> level{n};
    int d = 0;
    while (d < 5) {
        std::vector<long long> next;
        for (long long m : level) {
            if (m <= 128) return d;
            long long n2 = m / 2;
            n2 -= n2 % 8;
            next.push_back(n2);
            next.push_back(m - n2);
        }
        level.swap(next);
        ++d;
    }
    return d;
}

// leaves of pairwise_sum(n): size[l] (<= 128) and merges[l]; returns the count. One thread.
__device__ int build_program(int n, uint8_t* size, uint8_t* merges) {
    int stack_n[40];
    uint8_t stack_m[40];
    int sp = 0, nl = 0;
    stack_n[sp] = n;
    stack_m[sp++] = 0;
    while (sp > 0) {
        --sp;
        const int s = stack_n[sp];
        const int m = stack_m[sp];
        if (s <= 128) {
            size[nl] = (uint8_t)s;      // 128 fits
            merges[nl++] = (uint8_t)m;
        } else {
            int n2 = s / 2;
            n2 -= n2 % 8;
            stack_n[sp] = s - n2;       // right half: finishes this node's merge and everything pending above it
            stack_m[sp++] = (uint8_t)(m + 1);
            stack_n[sp] = n2;           // left half first
            stack_m[sp++] = 0;
        }
    }
    return nl;
}

// column range and leaf program of each of the G subtrees of pairwise_sum(n) at depth lg = log2(G): thread g walks
// from the root to ITS subtree (bit k of g, from the top: right or left child) and flattens that subtree only, so the
// set-up costs one short serial program per thread instead of one long one per CTA.
struct SegPlan {
    int start[GMAX], len[GMAX], prog[GMAX];   // first column, columns, offset of the program in psize / pmerge
    int max_len;
};
__device__ void build_subtree(int n, int lg, int g, int stride, SegPlan* sp, uint8_t* psize, uint8_t* pmerge) {
    int m = n, col = 0;
    for (int k = lg - 1; k >= 0; --k) {
        int n2 = m / 2;
        n2 -= n2 % 8;
        if ((g >> k) & 1) {
            col += n2;
            m -= n2;
        } else {
            m = n2;
        }
    }
    sp->start[g] = col;
    sp->len[g] = m;
    sp->prog[g] = g * stride;
    const int nl = build_program(m, psize + g * stride, pmerge + g * stride);
    psize[g * stride + nl] = 0;                      // sentinel read after the subtree's last leaf
    if (g == (1 << lg) - 1) sp->max_len = m;        // right halves are never shorter: the rightmost subtree is the longest
}

__device__ __forceinline__ void cp_async_16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
                 : "memory");
}

// sum of data[row[0 .. n)] in numpy's pairwise order for the CTA's rows; `rows` = the CTA's first index row + the
// segment's first column. The result is valid in every thread of a row's group.
// Every thread consumes only its own tile row, and a tile row is fetched by the thread itself and its lane neighbour
// (the two request the two 16-byte halves of every 32-byte sector in the same instruction), so the ring needs no CTA
// barrier: cp.async.wait_group + __syncwarp.
__device__ double pairwise_rows(const double* __restrict__ data, const int32_t* __restrict__ rows, long long ld, int n_rows,
                                int lg, bool vec_ok, const SegPlan& sp, const uint8_t* psize, const uint8_t* pmerge,
                                int32_t* tiles) {
    const int tid = threadIdx.x;
    const int G = 1 << lg;
    const int n_tiles = (sp.max_len + GC - 1) / GC;
    // the two tile rows this thread helps to fetch: its own and its lane neighbour's
    const int vrA = tid & ~1, vrB = tid | 1, odd = tid & 1;
    const int lenA = (vrA >> lg) < n_rows ? sp.len[vrA & (G - 1)] : 0;
    const int lenB = (vrB >> lg) < n_rows ? sp.len[vrB & (G - 1)] : 0;
    const int32_t* srcA = rows + (long long)(vrA >> lg) * ld + sp.start[vrA & (G - 1)];
    const int32_t* srcB = rows + (long long)(vrB >> lg) * ld + sp.start[vrB & (G - 1)];
    auto load_tile = [&](int t) {
        int32_t* dst = tiles + (size_t)(t % GST) * GR * GC;
        const int col0 = t * GC;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int vr = k < 4 ? vrA : vrB, len = k < 4 ? lenA : lenB;
            const int ch = 2 * (k & 3) + odd;
            const int col = col0 + 4 * ch;
            if (col < len) {
                const int32_t* src = (k < 4 ? srcA : srcB) + col;
                int32_t* d = dst + vr * GC + ((ch ^ (vr & 7)) << 2);
                if (vec_ok && col + 4 <= len) {
                    cp_async_16(d, src);
                } else {
                    const int m = min(4, len - col);
                    for (int e = 0; e < m; ++e) cp_async_4(d + e, src + e);
                }
            }
        }
    };
    for (int t = 0; t < GST - 1; ++t) {
        if (t < n_tiles) load_tile(t);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const int my_g = tid & (G - 1);
    const bool active = (tid >> lg) < n_rows;
    const int len = active ? sp.len[my_g] : 0;
    psize += sp.prog[my_g];
    pmerge += sp.prog[my_g];
    double stack[32];
    int sptr = 0, li = 0;
    int full_left = psize[0] >> 3, tail = psize[0] & 7;
    // -0.0 is the one start value that leaves every first addend unchanged, bit for bit (numpy starts from a[0..7])
    double r0 = -0.0, r1 = -0.0, r2 = -0.0, r3 = -0.0, r4 = -0.0, r5 = -0.0, r6 = -0.0, r7 = -0.0, res = 0.0;
    bool have_res = false;
    for (int t = 0; t < n_tiles; ++t) {
        asm volatile("cp.async.wait_group %0;" ::"n"(GST - 2) : "memory");
        __syncwarp();  // the neighbour's half of tile t has landed too; both lanes are done with tile t - 1
        if (t + GST - 1 < n_tiles) load_tile(t + GST - 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const int32_t* trow = tiles + (size_t)(t % GST) * GR * GC + tid * GC;
        const int col0 = t * GC;
#pragma unroll
        for (int o = 0; o < GC / 8; ++o) {
            const int col = col0 + 8 * o;
            if (col >= len) break;
            const int4 ia = *reinterpret_cast<const int4*>(trow + (((2 * o) ^ (tid & 7)) << 2));
            const int4 ib = *reinterpret_cast<const int4*>(trow + (((2 * o + 1) ^ (tid & 7)) << 2));
            if (full_left > 0) {
                r0 += data[ia.x]; r1 += data[ia.y]; r2 += data[ia.z]; r3 += data[ia.w];
                r4 += data[ib.x]; r5 += data[ib.y]; r6 += data[ib.z]; r7 += data[ib.w];
                if (--full_left > 0) continue;
                res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
                r0 = r1 = r2 = r3 = r4 = r5 = r6 = r7 = -0.0;
                have_res = true;
                if (tail > 0) continue;  // the leftovers of the last leaf follow in the next octet
            } else {
                // fewer than 8 elements: a leaf shorter than 8 (sequential from 0.0) or the leftovers of the last leaf
                const int idx8[8] = {ia.x, ia.y, ia.z, ia.w, ib.x, ib.y, ib.z, ib.w};
                if (!have_res) res = 0.0;
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (e < tail) res += data[idx8[e]];
            }
            // ---- the leaf is complete: push, merge, next leaf
            stack[sptr++] = res;
            for (int m = pmerge[li]; m > 0; --m) {
                --sptr;
                stack[sptr - 1] = stack[sptr - 1] + stack[sptr];
            }
            ++li;
            full_left = psize[li] >> 3;   // one sentinel entry follows the last leaf
            tail = psize[li] & 7;
            have_res = false;
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();  // the lane pair's ring rows are free for the next segment
    // ---- the G subtree sums of a row are added in the recursion's order: neighbours first
    double s = active ? stack[0] : 0.0;
    for (int o = 1; o < G; o <<= 1) {
        const double other = __shfl_xor_sync(0xffffffffu, s, o);
        s = s + other;        // only the left member of each pair is used further up (and IEEE addition commutes)
    }
    return s;
}

// out[r] = mean(data[idx[r, 0:n1]])                       (n2 == 0, K4)
// out[r] = | mean(data[idx[r, 0:n1]]) - mean(data[idx[r, n1:n1+n2]]) |   (n2 > 0, K5)
template <bool DATA_SMEM>
__global__ void __launch_bounds__(GR)
gather_mean_kernel(const double* __restrict__ data, int n_data, const int32_t* __restrict__ idx, long long ld, int R,
                   int n1, int n2, int lg, int vec1, int vec2, double* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t gsm[];
    __shared__ SegPlan sp1, sp2;
    int32_t* tiles = reinterpret_cast<int32_t*>(gsm);                            // [GST][GR][GC]
    uint8_t* prog = gsm + (size_t)GST * GR * GC * 4;
    const int cap1 = leaves_cap(n1, lg), cap2 = n2 > 0 ? leaves_cap(n2, lg) : 0;
    const int stride1 = prog_stride(n1, lg), stride2 = prog_stride(n2, lg);
    uint8_t* ps1 = prog;
    uint8_t* pm1 = ps1 + cap1;
    uint8_t* ps2 = pm1 + cap1;
    uint8_t* pm2 = ps2 + cap2;
    double* sm_data = reinterpret_cast<double*>(gsm + (((size_t)GST * GR * GC * 4 + 2 * (size_t)(cap1 + cap2) + 15) & ~size_t(15)));
    // blockIdx.y = batch entry (its own data vector, index rows and outputs): nine CIs of a row in one launch
    data += (long long)blockIdx.y * n_data;
    idx += (long long)blockIdx.y * R * ld;
    out += (long long)blockIdx.y * R;
    const int G = 1 << lg;
    if ((int)threadIdx.x < G) build_subtree(n1, lg, threadIdx.x, stride1, &sp1, ps1, pm1);
    if (n2 > 0 && (int)threadIdx.x >= 32 && (int)threadIdx.x < 32 + G)
        build_subtree(n2, lg, threadIdx.x - 32, stride2, &sp2, ps2, pm2);
    const double* d = data;
    if (DATA_SMEM) {
        for (int i = threadIdx.x; i < n_data; i += GR) sm_data[i] = data[i];
        d = sm_data;
    }
    __syncthreads();
    // persistent over row blocks: the programs and the data vector are set up once per CTA
    const int rows_per_cta = GR >> lg;
    const int r = threadIdx.x >> lg;
    for (int row0 = blockIdx.x * rows_per_cta; row0 < R; row0 += gridDim.x * rows_per_cta) {
        const int n_rows = min(rows_per_cta, R - row0);
        const int32_t* rows = idx + (long long)row0 * ld;
        const double s1 = pairwise_rows(d, rows, ld, n_rows, lg, vec1 != 0, sp1, ps1, pm1, tiles);
        double res = s1 / (double)n1;
        if (n2 > 0) {
            const double s2 = pairwise_rows(d, rows + n1, ld, n_rows, lg, vec2 != 0, sp2, ps2, pm2, tiles);
            res = fabs(res - s2 / (double)n2);
        }
        if ((threadIdx.x & (G - 1)) == 0 && r < n_rows) out[row0 + r] = res;
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

extern "C" int mmx_copies_drawn(const uint8_t* mult, int B, int64_t n_units, const int32_t* q_unit, int64_t nq,
                                int64_t* drawn, cudaStream_t stream) {
    MMX_REQUIRE(mult != nullptr && drawn != nullptr && B >= 0 && n_units > 0 && nq >= 0, "bad argument");
    if (B == 0) return 0;
    copies_drawn_kernel<<<B, 256, 0, stream>>>(mult, n_units, q_unit, nq, reinterpret_cast<long long*>(drawn));
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
                                                 n_units, nq, in_smem, correct, q_hits1);
    MMX_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mmx_boot_counts_lists_workspace_bytes(int B, int64_t n_units, size_t* bytes) {
    MMX_REQUIRE(bytes != nullptr && B >= 0 && n_units > 0, "bad argument");
    *bytes = (size_t)((B + CHUNK_B - 1) / CHUNK_B) * (size_t)n_units * CHUNK_B;
    return 0;
}

// K3 without a bitmask (LIST mode), straight from the per-shard lists of K1. The first min(cnt_above[q], L) entries of
// query q's concatenated lists are walked — all of its above-GT set when cnt_above[q] <= L. (iteration, query) pairs whose
// truncated list did not reach a multiplicity of 10 are left out of `correct` and reported in unres [unres_cap, 2] /
// *unres_n (may exceed unres_cap: call again with more room).
extern "C" int mmx_boot_counts_lists(const int32_t* list_i, const int32_t* list_n, int n_shards, int64_t nq, int cap,
                                     const int32_t* col2unit, int L, const int32_t* cnt_above, const int32_t* q_unit,
                                     const uint8_t* mult, int B, int64_t n_units, int32_t* correct, int32_t* q_hits1,
                                     int32_t* unres, int32_t* unres_n, int unres_cap, void* workspace,
                                     size_t workspace_bytes, cudaStream_t stream) {
    MMX_REQUIRE(list_i && list_n && cnt_above && mult && correct && unres && unres_n && unres_cap >= 1, "NULL argument");
    MMX_REQUIRE(n_shards >= 1 && cap >= 1 && L >= 1 && L <= cap, "bad list geometry (1 <= L <= cap: all hits may sit in one shard)");
    MMX_CHECK_CUDA(cudaMemsetAsync(unres_n, 0, sizeof(int32_t), stream));
    if (B == 0) return 0;
    MMX_CHECK_CUDA(cudaMemsetAsync(correct, 0, (size_t)B * 3 * sizeof(int32_t), stream));
    if (nq == 0) return 0;
    size_t need = 0;
    mmx_boot_counts_lists_workspace_bytes(B, n_units, &need);
    MMX_REQUIRE(workspace != nullptr && workspace_bytes >= need, "workspace too small (mmx_boot_counts_lists_workspace_bytes)");
    MMX_REQUIRE(((uintptr_t)workspace & 15) == 0, "workspace must be 16-byte aligned");
    const int n_chunks = (B + CHUNK_B - 1) / CHUNK_B;
    uint8_t* mult_t = static_cast<uint8_t*>(workspace);
    transpose_mult_kernel<<<dim3((unsigned)((n_units + CHUNK_B - 1) / CHUNK_B), (unsigned)n_chunks), 256, 0, stream>>>(
        mult, B, n_units, mult_t);
    MMX_CHECK_CUDA(cudaGetLastError());
    int per_sm = 0;  // a persistent grid: every CTA resident at once, the warps stride over the queries
    MMX_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, boot_counts_lists_kernel, 256, 0));
    const long long resident = (long long)sm_count() * std::max(per_sm, 1);
    const long long ctas = std::min<long long>((nq + 7) / 8, std::max<long long>(resident / n_chunks, 1));
    boot_counts_lists_kernel<<<dim3((unsigned)ctas, (unsigned)n_chunks), 256, 0, stream>>>(
        list_i, list_n, n_shards, nq, cap, col2unit, L, cnt_above, q_unit, mult_t, B, n_units, correct, q_hits1, unres,
        unres_n, unres_cap);
    MMX_CHECK_CUDA(cudaGetLastError());
    return 0;
}

static int launch_gather_mean(const double* data, int n_data, const int32_t* idx, int64_t ld, int R, int n1, int n2,
                              double* out, cudaStream_t stream, int batch = 1) {
    MMX_REQUIRE(n1 >= 1 && n2 >= 0 && ld >= (int64_t)n1 + n2, "bad segment sizes");
    MMX_REQUIRE(batch >= 1 && batch <= 65535, "batch out of range");
    if (R == 0) return 0;
    // threads per row: as many subtrees as the recursion offers (both segments), until ~64 warps per SM are in flight
    int lg = full_split_depth(n1);
    if (n2 > 0) lg = std::min(lg, full_split_depth(n2));
    while (lg > 0 && ((long long)R * batch << (lg - 1)) >= (long long)sm_count() * 2048) --lg;
    const size_t prog = 2 * (size_t)(leaves_cap(n1, lg) + (n2 > 0 ? leaves_cap(n2, lg) : 0));
    const size_t base = (((size_t)GST * GR * GC * 4 + prog + 15) & ~size_t(15));
    MMX_REQUIRE(base <= 200 * 1024, "vectors too long for the pairwise-sum program (n1=%d n2=%d)", n1, n2);
    const bool in_smem = (size_t)n_data * 8 <= (size_t)DATA_SMEM_MAX && base + (size_t)n_data * 8 <= 200 * 1024;
    const size_t smem = base + (in_smem ? (size_t)n_data * 8 : 0);
    // 16-byte copies need every column range to start on a 16-byte boundary (subtree boundaries are multiples of 8)
    const bool al = (reinterpret_cast<uintptr_t>(idx) % 16 == 0) && (ld % 4 == 0) && ((int64_t)R * ld % 4 == 0);
    const int vec1 = al ? 1 : 0, vec2 = (al && n1 % 4 == 0) ? 1 : 0;
    static const char gather_tag = 0;
    if (first_use_on_device(&gather_tag)) {
        MMX_CHECK_CUDA(cudaFuncSetAttribute(gather_mean_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        MMX_CHECK_CUDA(cudaFuncSetAttribute(gather_mean_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    const int rows_per_cta = GR >> lg;
    const int row_blocks = (R + rows_per_cta - 1) / rows_per_cta;
    const int ctas_per_sm = std::max(1, (int)((220 * 1024) / (smem + 1024)));
    const dim3 grid(std::min(row_blocks, std::max(1, sm_count() * ctas_per_sm / batch)), batch);
    if (in_smem)
        gather_mean_kernel<true><<<grid, GR, smem, stream>>>(data, n_data, idx, ld, R, n1, n2, lg, vec1, vec2, out);
    else
        gather_mean_kernel<false><<<grid, GR, smem, stream>>>(data, n_data, idx, ld, R, n1, n2, lg, vec1, vec2, out);
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

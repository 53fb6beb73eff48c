// K2 — ColPali late-interaction MaxSim on tcgen05 / TMEM, fed by TMA.
//
//   score[q, d] = sum_{t in tokens(q)} max_{s in patches(d)} <Q[t, :], P[s, :]>          (d = 128, bf16 in, fp32 acc)
//
// Replaces the reference's 10x10 chunk loop around `processor.score`
// (/root/reference/main.py:451-470; colpali_engine score_multi_vector:
//  einsum("bnd,csd->bcns").max(dim=3)[0].sum(dim=2)) — the token x patch tensor never leaves TMEM/registers.
//
// Layout in HBM: both sides are FLAT row-major bf16 [sum(L), 128] token matrices plus int32 CSR offsets, so a
// 1030-patch document does not have to be padded to a tile multiple.
//
// CTA = one persistent worker. Work item = (pair of 128-row query-token tiles A0/A1, chunk of whole documents).
// The two A tiles stay resident in shared memory for the whole item; the chunk's patches stream through a
// 2-stage TMA ring as 256-row B tiles. Each B tile feeds two 128x256x128 MMAs (A0*B -> TMEM cols [0,256),
// A1*B -> TMEM cols [256,512)), so every byte of B pulled from L2 is used for 256 query rows and the two
// accumulators ping-pong: while the tensor core fills one, an epilogue warpgroup drains the other.
//
// Epilogue (thread == one query-token row, TMEM lane == row): running max over the row's 256 fp32 columns with
// FMNMX3, cut at document boundaries (columns are global patch indices, boundaries come from d_off). When a
// document ends, the 128 row maxima go through shared memory and are summed per query segment by a warp (fixed
// shuffle tree -> deterministic), one float per (query, doc) is stored. A query whose tokens straddle two row
// tiles gets its tail part in `carry[tile, doc]`, folded in by maxsim_fixup_kernel in tile order.
#include <cfloat>
#include <climits>

#include "../../include/mmx.h"
#include "mmx_host.cuh"
#include "sm100_ptx.cuh"

namespace mmx {
namespace maxsim {

constexpr int BM = 128;                          // query-token rows per MMA (TMEM lanes)
constexpr int BN = 256;                          // document patches per MMA (TMEM columns)
constexpr int DIM = 128;                         // embedding width (= K of the contraction)
constexpr int A_HALF_BYTES = BM * 128;           // one 64-wide K half of an A tile, 128 B rows
constexpr int B_HALF_BYTES = BN * 128;
constexpr int A_TILE_BYTES = 2 * A_HALF_BYTES;   // 32 KB
constexpr int B_STAGE_BYTES = 2 * B_HALF_BYTES;  // 64 KB
constexpr int NSTAGE = 2;
constexpr int NUM_THREADS = 384;                 // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 idle, 2 x 4 epilogue
constexpr int TMEM_COLS = 512;

struct Params {
    const int32_t* q_off;   // [nq+1] token offsets
    const int32_t* d_off;   // [nd+1] patch offsets
    const int32_t* tok2q;   // [n_qtok] owning query of each token row
    const uint8_t* d_zero;  // nullable [nd]: 1 = a zero-padded row takes part in this doc's max (pad_sequence)
    float* scores;          // [nq, ld_scores]
    long long ld_scores;
    float* carry;           // [n_mtiles, nd] partial sums of queries continued from the previous row tile
    int nq, nd, n_qtok, n_dtok;
    int n_mtiles, n_mpairs, n_chunks, docs_per_chunk;
};

struct __align__(1024) Smem {
    uint8_t a[2][A_TILE_BYTES];
    uint8_t b[NSTAGE][B_STAGE_BYTES];
    float stage[2][2][BM];  // [warpgroup][slot][row] row maxima of the document being closed
    int seg_start[2][BM + 1];
    int seg_q[2][BM];
    int warp_heads[2][4];
    uint64_t a_full, a_free;
    uint64_t b_full[NSTAGE], b_empty[NSTAGE];
    uint64_t t_full[2], t_empty[2];
    uint32_t tmem_base;
};

struct ItemGeom {
    int mp, d0, d1, p0, ntiles;
};

__device__ __forceinline__ ItemGeom item_geom(const Params& p, int item) {
    ItemGeom g;
    g.mp = item % p.n_mpairs;
    const int ch = item / p.n_mpairs;
    g.d0 = ch * p.docs_per_chunk;
    g.d1 = min(p.nd, g.d0 + p.docs_per_chunk);
    g.p0 = __ldg(p.d_off + g.d0);
    const int p1 = __ldg(p.d_off + g.d1);
    g.ntiles = (p1 - g.p0 + BN - 1) / BN;
    return g;
}

__device__ __forceinline__ float warp_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// Per-thread state of the epilogue's walk along the patch axis.
struct Walk {
    int cur_doc, cur_end, d1;
    float m;
    bool done;
    int nflush;
};

struct EpiCtx {
    Smem* s;
    const Params* p;
    int g, wq, lane, row, nseg, mt;
    bool cont_first;
};

// Close document `doc`: row maxima -> shared memory -> per-query-segment sums -> one float per (query, doc).
// Scalars by value so the walk state stays in registers across the (rare) call.
__device__ __noinline__ void flush_doc(Smem* sp, const Params* pp, int g, int wq, int lane, int row, int nseg,
                                       int mt, int cont_first, int doc, float m, int nflush) {
    const Params& p = *pp;
    Smem& s = *sp;
    if (p.d_zero != nullptr && p.d_zero[doc]) m = fmaxf(m, 0.0f);
    const int slot = nflush & 1;
    s.stage[g][slot][row] = m;
    named_bar_sync(1 + g, 128);
    for (int si = (wq + nflush) & 3; si < nseg; si += 4) {
        const int a = s.seg_start[g][si], b = s.seg_start[g][si + 1];
        float sum = 0.0f;
        for (int r = a + lane; r < b; r += 32) sum += s.stage[g][slot][r];
        sum = warp_sum(sum);
        if (lane == 0) {
            const int q = s.seg_q[g][si];
            if (q >= 0) {
                if (si == 0 && cont_first)
                    p.carry[(long long)mt * p.nd + doc] = sum;
                else
                    p.scores[(long long)q * p.ld_scores + doc] = sum;
            }
        }
    }
}

__device__ __forceinline__ void flush_doc(const EpiCtx& c, Walk& w) {
    flush_doc(c.s, c.p, c.g, c.wq, c.lane, c.row, c.nseg, c.mt, c.cont_first ? 1 : 0, w.cur_doc, w.m, w.nflush);
    w.nflush++;
}

// Fold 32 accumulator columns (global patch indices [pc, pc+32)) into the walk.
__device__ __forceinline__ void fold_chunk(const EpiCtx& c, Walk& w, const uint32_t (&v)[32], int pc) {
    if (w.done) return;
    int rem = w.cur_end - pc;  // columns of this chunk still owned by cur_doc
    if (rem > 32) {
        float m0 = w.m, m1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            m0 = fmaxf(m0, fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
            m1 = fmaxf(m1, fmaxf(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
        }
        w.m = fmaxf(m0, m1);
        return;
    }
    int lo = 0;
    while (true) {
        const int hi = min(rem, 32);
        float m = w.m;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float x = (j >= lo && j < hi) ? __uint_as_float(v[j]) : -INFINITY;
            m = fmaxf(m, x);
        }
        w.m = m;
        if (rem > 32) break;
        flush_doc(c, w);
        w.m = -INFINITY;
        lo = max(lo, hi);
        w.cur_doc++;
        if (w.cur_doc >= w.d1) {
            w.done = true;
            break;
        }
        w.cur_end = __ldg(c.p->d_off + w.cur_doc + 1);
        rem = w.cur_end - pc;
        if (lo >= 32) break;
    }
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
maxsim_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ Params p) {
    extern __shared__ uint8_t smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_d);
    }
    if (warp == 1 && elect_one()) {
        mbar_init(&s.a_full, 1);
        mbar_init(&s.a_free, 1);
        for (int i = 0; i < NSTAGE; ++i) {
            mbar_init(&s.b_full[i], 1);
            mbar_init(&s.b_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s.t_full[i], 1);
            mbar_init(&s.t_empty[i], 4);  // one arrival per epilogue warp
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(&s.tmem_base, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s.tmem_base;
    const int n_items = p.n_mpairs * p.n_chunks;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (elect_one()) {
            uint32_t it = 0, bcount = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const ItemGeom g = item_geom(p, item);
                if (it > 0) mbar_wait(&s.a_free, (it - 1) & 1);
                mbar_arrive_expect_tx(&s.a_full, 2 * A_TILE_BYTES);
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        tma_load_2d(s.a[m] + h * A_HALF_BYTES, &tmap_q, &s.a_full, h * 64, (g.mp * 2 + m) * BM);
                for (int t = 0; t < g.ntiles; ++t, ++bcount) {
                    const uint32_t st = bcount % NSTAGE, ph = (bcount / NSTAGE) & 1;
                    mbar_wait(&s.b_empty[st], ph ^ 1);
                    mbar_arrive_expect_tx(&s.b_full[st], B_STAGE_BYTES);
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        tma_load_2d(s.b[st] + h * B_HALF_BYTES, &tmap_d, &s.b_full[st], h * 64, g.p0 + t * BN);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (one thread)
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
            uint32_t it = 0, bcount = 0, use[2] = {0, 0};
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const ItemGeom g = item_geom(p, item);
                const int nm = (g.mp * 2 + 1 < p.n_mtiles) ? 2 : 1;
                mbar_wait(&s.a_full, it & 1);
                tc_fence_after();
                for (int t = 0; t < g.ntiles; ++t, ++bcount) {
                    const uint32_t st = bcount % NSTAGE, ph = (bcount / NSTAGE) & 1;
                    mbar_wait(&s.b_full[st], ph);
                    tc_fence_after();
                    for (int m = 0; m < nm; ++m) {
                        mbar_wait(&s.t_empty[m], (use[m] & 1) ^ 1);
                        tc_fence_after();
#pragma unroll
                        for (int k = 0; k < DIM / 16; ++k) {
                            const uint32_t a_addr = smem_u32(s.a[m] + (k >> 2) * A_HALF_BYTES) + (k & 3) * 32;
                            const uint32_t b_addr = smem_u32(s.b[st] + (k >> 2) * B_HALF_BYTES) + (k & 3) * 32;
                            umma_bf16_ss(tmem_base + m * BN, umma_desc_kmajor_sw128(a_addr),
                                         umma_desc_kmajor_sw128(b_addr), idesc, k > 0);
                        }
                        umma_commit(&s.t_full[m]);
                        use[m]++;
                    }
                    umma_commit(&s.b_empty[st]);
                }
                umma_commit(&s.a_free);
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue warpgroups
        EpiCtx c;
        c.s = &s;
        c.p = &p;
        c.g = (warp - 4) >> 2;
        c.wq = warp & 3;  // TMEM lane quarter this warp may read
        c.lane = lane;
        c.row = c.wq * 32 + lane;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(c.wq * 32) << 16) + c.g * BN;
        uint32_t use = 0;
        Walk w;
        w.nflush = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const ItemGeom g = item_geom(p, item);
            c.mt = g.mp * 2 + c.g;
            if (c.mt >= p.n_mtiles) continue;
            // ---- query-segment table of this row tile
            const int r0 = c.mt * BM;
            const int rg = r0 + c.row;
            const int q = rg < p.n_qtok ? __ldg(p.tok2q + rg) : -1;
            const int qprev = (c.row == 0) ? INT_MIN : (rg - 1 < p.n_qtok ? __ldg(p.tok2q + rg - 1) : -1);
            const bool head = q != qprev;
            named_bar_sync(1 + c.g, 128);  // previous item's flush readers are done with the table
            const uint32_t hmask = __ballot_sync(0xffffffffu, head);
            if (lane == 0) s.warp_heads[c.g][c.wq] = __popc(hmask);
            named_bar_sync(1 + c.g, 128);
            int prefix = 0, total = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int h = s.warp_heads[c.g][i];
                if (i < c.wq) prefix += h;
                total += h;
            }
            if (head) {
                const int idx = prefix + __popc(hmask & ((1u << lane) - 1u));
                s.seg_start[c.g][idx] = c.row;
                s.seg_q[c.g][idx] = q;
            }
            if (c.row == 0) s.seg_start[c.g][total] = BM;
            c.nseg = total;
            c.cont_first = (r0 > 0) && (r0 < p.n_qtok) && (__ldg(p.tok2q + r0 - 1) == __ldg(p.tok2q + r0));
            named_bar_sync(1 + c.g, 128);

            // ---- walk the chunk's patches
            w.cur_doc = g.d0;
            w.d1 = g.d1;
            w.cur_end = __ldg(p.d_off + g.d0 + 1);
            w.m = -INFINITY;
            w.done = g.d0 >= g.d1;
            for (int t = 0; t < g.ntiles; ++t, ++use) {
                mbar_wait(&s.t_full[c.g], use & 1);
                tc_fence_after();
                const int pt = g.p0 + t * BN;
                uint32_t va[32], vb[32];
                tmem_ld_32x32(t_lane, va);
                tmem_ld_wait();
#pragma unroll 1
                for (int ch = 0; ch < BN / 32; ch += 2) {
                    tmem_ld_32x32(t_lane + (ch + 1) * 32, vb);
                    fold_chunk(c, w, va, pt + ch * 32);
                    tmem_ld_wait();
                    if (ch + 2 < BN / 32) {
                        tmem_ld_32x32(t_lane + (ch + 2) * 32, va);
                    } else {
                        // every column of this accumulator is in registers: hand it back to the MMA warp
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&s.t_empty[c.g]);
                    }
                    fold_chunk(c, w, vb, pt + (ch + 1) * 32);
                    if (ch + 2 < BN / 32) tmem_ld_wait();
                }
            }
            // documents that end exactly at the chunk end were flushed above; what is left has no patches
            while (!w.done) {
                flush_doc(c, w);
                w.m = -INFINITY;
                w.cur_doc++;
                if (w.cur_doc >= w.d1) w.done = true;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// tok2q[t] = query that owns token row t (binary search in q_off).
__global__ void tok2q_kernel(const int32_t* __restrict__ q_off, int nq, int n_qtok, int32_t* __restrict__ tok2q) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_qtok) return;
    int lo = 0, hi = nq;  // find largest q with q_off[q] <= t
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(q_off + mid) <= t) lo = mid; else hi = mid;
    }
    // skip empty queries that share the same offset: the owner is the LAST q with q_off[q] <= t
    tok2q[t] = lo;
}

// scores[q, d] += carry[mt, d] for every row tile mt that starts in the middle of query q, in tile order.
__global__ void maxsim_fixup_kernel(const int32_t* __restrict__ tok2q, const int32_t* __restrict__ q_off,
                                    const float* __restrict__ carry, float* __restrict__ scores, long long ld,
                                    int n_qtok, int n_mtiles, int nd) {
    const int d = blockIdx.y * blockDim.x + threadIdx.x;
    const int mt = blockIdx.x + 1;
    if (d >= nd || mt >= n_mtiles) return;
    const int r0 = mt * BM;
    const int q = tok2q[r0];
    if (tok2q[r0 - 1] != q) return;             // tile starts with a fresh query
    if (q_off[q] < r0 - BM) return;             // an earlier tile owns the head of this chain
    float acc = scores[(long long)q * ld + d];
    for (int m2 = mt; m2 < n_mtiles && m2 * BM < n_qtok && tok2q[m2 * BM] == q; ++m2)
        acc += carry[(long long)m2 * nd + d];
    scores[(long long)q * ld + d] = acc;
}

}  // namespace maxsim
}  // namespace mmx

using namespace mmx;
using namespace mmx::maxsim;

extern "C" int mmx_maxsim_workspace_bytes(int64_t n_qtok, int64_t nd, size_t* bytes) {
    MMX_REQUIRE(bytes != nullptr, "bytes == NULL");
    const int64_t n_mtiles = (n_qtok + BM - 1) / BM;
    size_t b = 0;
    b += ((size_t)n_qtok * 4 + 255) & ~size_t(255);
    b += (size_t)n_mtiles * (size_t)nd * 4;
    *bytes = b + 256;
    return 0;
}

extern "C" int mmx_maxsim_scores(const void* q_tok, const int32_t* q_off, int64_t nq, int64_t n_qtok,
                                 const void* d_tok, const int32_t* d_off, int64_t nd, int64_t n_dtok, int dim,
                                 const uint8_t* d_zero, float* scores, int64_t ld_scores, void* workspace,
                                 size_t workspace_bytes, cudaStream_t stream) {
    MMX_REQUIRE(dim == DIM, "mmx_maxsim_scores: only dim=128 (ColPali) is implemented, got %d", dim);
    MMX_REQUIRE(nq >= 0 && nd >= 0 && n_qtok >= 0 && n_dtok >= 0, "negative size");
    MMX_REQUIRE(n_qtok < (1ll << 31) - 512 && n_dtok < (1ll << 31) - 512 && nq < (1ll << 31) && nd < (1ll << 31),
                "sizes exceed int32 offsets");
    MMX_REQUIRE(ld_scores >= nd, "ld_scores < nd");
    if (nq == 0 || nd == 0) return 0;
    if (mmx_device_check() != 0) return -3;
    // every (q, d) with at least one token is overwritten; empty queries score 0
    MMX_CHECK_CUDA(cudaMemsetAsync(scores, 0, (size_t)nq * ld_scores * sizeof(float), stream));
    if (n_qtok == 0) return 0;
    size_t need = 0;
    mmx_maxsim_workspace_bytes(n_qtok, nd, &need);
    MMX_REQUIRE(workspace != nullptr && workspace_bytes >= need, "workspace too small: need %zu bytes, got %zu", need,
                workspace_bytes);

    Params p;
    p.q_off = q_off;
    p.d_off = d_off;
    uintptr_t ws = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
    int32_t* tok2q = reinterpret_cast<int32_t*>(ws);
    ws += ((size_t)n_qtok * 4 + 255) & ~size_t(255);
    p.tok2q = tok2q;
    p.carry = reinterpret_cast<float*>(ws);
    p.d_zero = d_zero;
    p.scores = scores;
    p.ld_scores = ld_scores;
    p.nq = (int)nq;
    p.nd = (int)nd;
    p.n_qtok = (int)n_qtok;
    p.n_dtok = (int)n_dtok;
    p.n_mtiles = (int)((n_qtok + BM - 1) / BM);
    p.n_mpairs = (p.n_mtiles + 1) / 2;
    const int sms = sm_count();
    long long want_chunks = ((long long)sms * 32 + p.n_mpairs - 1) / p.n_mpairs;
    if (want_chunks < 1) want_chunks = 1;
    if (want_chunks > nd) want_chunks = nd;
    p.docs_per_chunk = (int)((nd + want_chunks - 1) / want_chunks);
    p.n_chunks = (int)((nd + p.docs_per_chunk - 1) / p.docs_per_chunk);

    CUtensorMap tq, td;
    if (make_tmap_bf16_sw128(&tq, q_tok, (uint64_t)n_qtok, DIM, DIM, BM) != 0) return -4;
    if (make_tmap_bf16_sw128(&td, d_tok, (uint64_t)(n_dtok > 0 ? n_dtok : 1), DIM, DIM, BN) != 0) return -4;

    tok2q_kernel<<<(unsigned)((n_qtok + 255) / 256), 256, 0, stream>>>(q_off, (int)nq, (int)n_qtok, tok2q);
    MMX_CHECK_CUDA(cudaGetLastError());

    const size_t smem_bytes = sizeof(Smem) + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        MMX_CHECK_CUDA(cudaFuncSetAttribute(maxsim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        attr_set = true;
    }
    const long long n_items = (long long)p.n_mpairs * p.n_chunks;
    const int grid = (int)(n_items < sms ? n_items : sms);
    maxsim_kernel<<<grid, NUM_THREADS, smem_bytes, stream>>>(tq, td, p);
    MMX_CHECK_CUDA(cudaGetLastError());

    if (p.n_mtiles > 1) {
        dim3 fg((unsigned)(p.n_mtiles - 1), (unsigned)((nd + 255) / 256));
        maxsim_fixup_kernel<<<fg, 256, 0, stream>>>(tok2q, q_off, p.carry, scores, ld_scores, (int)n_qtok,
                                                    p.n_mtiles, (int)nd);
        MMX_CHECK_CUDA(cudaGetLastError());
    }
    return 0;
}

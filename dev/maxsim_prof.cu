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
// Chunks are sized to stay resident in L2 (~24 MB of patches) while every row-tile pair sweeps over them, so HBM
// sees each patch about once. The two A tiles stay in shared memory for the whole item; the chunk's patches
// stream through a 2-stage TMA ring as 256-row B tiles. Each B tile feeds two 128x256x128 MMAs
// (A0*B -> TMEM cols [0,256), A1*B -> TMEM cols [256,512)): every byte of B pulled from L2 serves 256 query rows
// and the two accumulators ping-pong — while the tensor core fills one, the epilogue drains the other.
//
// Epilogue: 8 warps. TMEM lane == query-token row; warps e and e+4 share a lane quarter and split the 256 columns
// of the SAME accumulator in halves, so one accumulator is drained in half the time (the drain latency, not the
// instruction count, is what the tensor pipe waits for). Per 32-column chunk: tcgen05.ld.32x32b.x32 + a FMNMX3 tree,
// software-pipelined. Columns are global patch indices; document boundaries come from d_off. When a document
// closes, both halves park their 128 row maxima in shared memory, a warp takes max(half0, half1) per row and sums
// each query segment with a fixed shuffle tree (deterministic), and one float per (query, doc) is stored.
// A query whose tokens straddle two row tiles gets its tail part in `carry[tile, doc]`, folded in by
// maxsim_fixup_kernel in tile order — no floating-point atomics anywhere.
#include <cfloat>
#include <climits>

#include "../include/mmx.h"
#include "../multimodal-embedding_b200/csrc/mmx_host.cuh"
#include "../multimodal-embedding_b200/csrc/sm100_ptx.cuh"

namespace mmx {
namespace maxsim {
__device__ unsigned long long g_prof[148*16*8];
#define PROF(slot, expr) do { long long _t0 = clock64(); expr; prof[slot] += clock64() - _t0; } while (0)

constexpr int BM = 128;                          // query-token rows per MMA (TMEM lanes)
constexpr int BN = 256;                          // document patches per MMA (TMEM columns)
constexpr int DIM = 128;                         // embedding width (= K of the contraction)
constexpr int A_HALF_BYTES = BM * 128;           // one 64-wide K half of an A tile, 128 B rows
constexpr int B_HALF_BYTES = BN * 128;
constexpr int A_TILE_BYTES = 2 * A_HALF_BYTES;   // 32 KB
constexpr int B_STAGE_BYTES = 2 * B_HALF_BYTES;  // 64 KB
constexpr int NSTAGE = 2;
constexpr int NUM_THREADS = 384;                 // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 idle, 8 epilogue
constexpr int NUM_EPI_THREADS = 256;
constexpr int TMEM_COLS = 512;
constexpr int HALF_COLS = BN / 2;                // columns of one accumulator drained by one warp
constexpr long long L2_CHUNK_BYTES = 24ll << 20; // document chunk kept hot in the 126 MB L2

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
    float stage[2][2][2][BM];  // [accumulator][slot][column half][row] row maxima of the document being closed
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

// Per-thread walk along the patch axis for one accumulator (= one row tile): which document the next column
// belongs to, and the running row maximum of that document over this thread's column half.
struct Walk {
    int cur_doc, cur_end;
    float m;
    int done;
    int nflush, nseg, mt, cont_first;
    uint32_t use;  // how many times this accumulator has been drained (mbarrier phase)
};

struct EpiCtx {
    Smem* s;
    const Params* p;
    int e, wq, half, lane, row, d1;
};

// Close document `doc` of accumulator g: both column halves park their row maxima, then per query segment
// sum_r max(half0[r], half1[r]) -> one float per (query, doc). Scalars by value: the walk stays in registers.
__device__ __noinline__ void flush_doc(Smem* sp, const Params* pp, int g, int e, int half, int lane, int row, int nseg,
                                       int mt, int cont_first, int doc, float m, int nflush) {
    const Params& p = *pp;
    Smem& s = *sp;
    if (p.d_zero != nullptr && p.d_zero[doc]) m = fmaxf(m, 0.0f);
    const int slot = nflush & 1;
    s.stage[g][slot][half][row] = m;
    named_bar_sync(1, NUM_EPI_THREADS);
    for (int si = (e + nflush) & 7; si < nseg; si += 8) {
        const int a = s.seg_start[g][si], b = s.seg_start[g][si + 1];
        float sum = 0.0f;
        for (int r = a + lane; r < b; r += 32) sum += fmaxf(s.stage[g][slot][0][r], s.stage[g][slot][1][r]);
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

__device__ __forceinline__ void close_doc(const EpiCtx& c, Walk& w, int g) {
    // A prefetched tcgen05.ld may still be writing its destination registers: it must land before the call below
    // lets the compiler spill / reuse registers (no-op when nothing is in flight).
    tmem_ld_wait();
    flush_doc(c.s, c.p, g, c.e, c.half, c.lane, c.row, w.nseg, w.mt, w.cont_first, w.cur_doc, w.m, w.nflush);
    w.nflush++;
    w.m = -INFINITY;
    w.cur_doc++;
    if (w.cur_doc >= c.d1)
        w.done = 1;
    else
        w.cur_end = __ldg(c.p->d_off + w.cur_doc + 1);
}

__device__ __forceinline__ float max3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));   // FMNMX3
    return r;
}

// max(m, v[0..32)): four independent FMNMX3 chains of depth 4, then two combines (asm-pinned so that ptxas does
// not re-serialise the associative maxima into two long dependent chains).
__device__ __forceinline__ float max32(const uint32_t (&v)[32], float m) {
    float c0 = max3(m, __uint_as_float(v[0]), __uint_as_float(v[1]));
    float c1 = max3(__uint_as_float(v[2]), __uint_as_float(v[3]), __uint_as_float(v[4]));
    float c2 = max3(__uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
    float c3 = max3(__uint_as_float(v[8]), __uint_as_float(v[9]), __uint_as_float(v[10]));
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        c0 = max3(c0, __uint_as_float(v[11 + 8 * j]), __uint_as_float(v[12 + 8 * j]));
        c1 = max3(c1, __uint_as_float(v[13 + 8 * j]), __uint_as_float(v[14 + 8 * j]));
        c2 = max3(c2, __uint_as_float(v[15 + 8 * j]), __uint_as_float(v[16 + 8 * j]));
        c3 = max3(c3, __uint_as_float(v[17 + 8 * j]), __uint_as_float(v[18 + 8 * j]));
    }
    c0 = max3(c0, __uint_as_float(v[27]), __uint_as_float(v[28]));
    c1 = max3(c1, __uint_as_float(v[29]), __uint_as_float(v[30]));
    c2 = fmaxf(c2, __uint_as_float(v[31]));
    return max3(max3(c0, c1, c2), c3, c3);
}

// Fold 32 accumulator columns this thread owns (global patch indices [pc, pc+32)) into the walk; documents may end
// anywhere inside the chunk (general path, kept small: it runs once per document per row tile, not per tile).
__device__ __forceinline__ void fold_chunk_general(const EpiCtx& c, Walk& w, int g, const uint32_t (&v)[32], int pc) {
    if (w.done) return;
    int rem = w.cur_end - pc;  // columns of this chunk still owned by cur_doc
    if (rem > 32) {
        w.m = max32(v, w.m);
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
        close_doc(c, w, g);
        lo = max(lo, hi);
        if (w.done || lo >= 32) break;
        rem = w.cur_end - pc;
    }
}

// Columns [.., pend) belong to the other column half: nothing to fold, but documents that end there close here too.
__device__ __forceinline__ void skip_to(const EpiCtx& c, Walk& w, int g, int pend) {
    while (!w.done && w.cur_end <= pend) close_doc(c, w, g);
}

__device__ __forceinline__ void release_accumulator(Smem& s, int a, int lane) {
    // this warp's 128 columns are in registers: hand the accumulator back to the MMA warp
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&s.t_empty[a]);
}

// Drain this warp's 128 columns of accumulator `a` for the B tile that starts at patch `pt`.
__device__ __forceinline__ void drain_accumulator(const EpiCtx& c, Smem& s, Walk& w, int a, uint32_t t_lane, int pt, long long* prof) {
    { long long _t0 = clock64(); mbar_wait(&s.t_full[a], w.use & 1); prof[0] += clock64() - _t0; }
    w.use++;
    long long _td = clock64();
    tc_fence_after();
    const uint32_t ta = t_lane + a * BN;
    const int pc0 = pt + c.half * HALF_COLS;
    uint32_t va[32];
    if (!w.done && w.cur_end - pt > BN) {
        uint32_t vb[32];
        // interior tile (3 of 4 at 1030 patches per document): every column belongs to cur_doc — straight-line,
        // the next chunk's TMEM load is in flight while the current one goes through the FMNMX3 tree.
        tmem_ld_32x32(ta, va);
        tmem_ld_wait();
        tmem_ld_32x32(ta + 32, vb);
        float m = max32(va, w.m);
        tmem_ld_wait();
        tmem_ld_32x32(ta + 64, va);
        m = max32(vb, m);
        tmem_ld_wait();
        tmem_ld_32x32(ta + 96, vb);
        m = max32(va, m);
        tmem_ld_wait();
        release_accumulator(s, a, c.lane);
        prof[1] += clock64() - _td; prof[4]++;
        w.m = max32(vb, m);
        prof[2] += clock64() - _td;
    } else {
        // a document ends (or the chunk of documents ends) inside this tile: one compact loop, one chunk at a time
        if (c.half == 1) skip_to(c, w, a, pc0);  // documents that end inside the other half's columns
#pragma unroll 1
        for (int ch = 0; ch < HALF_COLS / 32; ++ch) {
            tmem_ld_32x32(ta + ch * 32, va);
            tmem_ld_wait();
            if (ch == HALF_COLS / 32 - 1) release_accumulator(s, a, c.lane);
            fold_chunk_general(c, w, a, va, pc0 + ch * 32);
        }
        prof[3] += clock64() - _td; prof[5]++;
        if (c.half == 0) skip_to(c, w, a, pt + BN);
    }
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
maxsim_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_d,
              const __grid_constant__ Params p) {
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
            mbar_init(&s.t_empty[i], 8);  // one arrival per epilogue warp
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
            uint32_t it = 0, bcount = 0, use[2] = {0, 0}; long long prof[8] = {0,0,0,0,0,0,0,0}; long long _tstart = clock64();
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const ItemGeom g = item_geom(p, item);
                const int nm = (g.mp * 2 + 1 < p.n_mtiles) ? 2 : 1;
                mbar_wait(&s.a_full, it & 1);
                tc_fence_after();
                for (int t = 0; t < g.ntiles; ++t, ++bcount) {
                    const uint32_t st = bcount % NSTAGE, ph = (bcount / NSTAGE) & 1;
                    PROF(0, mbar_wait(&s.b_full[st], ph));
                    tc_fence_after();
                    for (int m = 0; m < nm; ++m) {
                        PROF(1 + m, mbar_wait(&s.t_empty[m], (use[m] & 1) ^ 1));
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
            prof[6] = clock64() - _tstart;
            for (int i = 0; i < 8; ++i) g_prof[(blockIdx.x * 16 + 1) * 8 + i] = prof[i];
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue (8 warps, 256 threads)
        EpiCtx c;
        c.s = &s;
        c.p = &p;
        c.e = warp - 4;
        c.wq = warp & 3;        // TMEM lane quarter this warp may read
        c.half = c.e >> 2;      // which 128 columns of every accumulator this warp drains
        c.lane = lane;
        c.row = c.wq * 32 + lane;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(c.wq * 32) << 16) + c.half * HALF_COLS;
        long long prof[8] = {0,0,0,0,0,0,0,0}; long long _tstart = clock64();
        Walk w0, w1;
        w0.nflush = w1.nflush = 0;
        w0.use = w1.use = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const ItemGeom g = item_geom(p, item);
            const int nm = (g.mp * 2 + 1 < p.n_mtiles) ? 2 : 1;
            c.d1 = g.d1;
            // ---- query-segment tables: the threads of column half h build the table of row tile h
            {
                const int tb = c.half;
                const int mt = g.mp * 2 + tb;
                const int r0 = mt * BM;
                const int rg = r0 + c.row;
                const int q = rg < p.n_qtok ? __ldg(p.tok2q + rg) : -1;
                const int qprev = (c.row == 0) ? INT_MIN : (rg - 1 < p.n_qtok ? __ldg(p.tok2q + rg - 1) : -1);
                const bool head = q != qprev;
                named_bar_sync(1, NUM_EPI_THREADS);  // the previous item's readers are done with the tables
                const uint32_t hmask = __ballot_sync(0xffffffffu, head);
                if (lane == 0) s.warp_heads[tb][c.wq] = __popc(hmask);
                named_bar_sync(1, NUM_EPI_THREADS);
                int prefix = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (i < c.wq) prefix += s.warp_heads[tb][i];
                if (head) {
                    const int idx = prefix + __popc(hmask & ((1u << lane) - 1u));
                    s.seg_start[tb][idx] = c.row;
                    s.seg_q[tb][idx] = q;
                }
                const int total = s.warp_heads[tb][0] + s.warp_heads[tb][1] + s.warp_heads[tb][2] + s.warp_heads[tb][3];
                if (c.row == 0) s.seg_start[tb][total] = BM;
                named_bar_sync(1, NUM_EPI_THREADS);
            }
            auto init_walk = [&](Walk& w, int a) {
                const int mt = g.mp * 2 + a;
                const int r0 = mt * BM;
                w.mt = mt;
                w.nseg = s.warp_heads[a][0] + s.warp_heads[a][1] + s.warp_heads[a][2] + s.warp_heads[a][3];
                w.cont_first = (r0 > 0) && (r0 < p.n_qtok) && (__ldg(p.tok2q + r0 - 1) == __ldg(p.tok2q + r0));
                w.cur_doc = g.d0;
                w.cur_end = __ldg(p.d_off + g.d0 + 1);
                w.m = -INFINITY;
                w.done = (g.d0 >= g.d1) || (a >= nm);
            };
            init_walk(w0, 0);
            init_walk(w1, 1);

            // ---- drain accumulator 0, then accumulator 1, of every B tile
#pragma unroll 1
            for (int t = 0; t < g.ntiles; ++t) {
                const int pt = g.p0 + t * BN;
#pragma unroll 1
                for (int a = 0; a < nm; ++a) {
                    drain_accumulator(c, s, w0, a, t_lane, pt, prof);   // one instance of the drain code for both
                    if (nm == 2) {                                  // accumulators: swap the two walks instead
                        const Walk tmp = w0;
                        w0 = w1;
                        w1 = tmp;
                    }
                }
            }
            // documents that end exactly at the chunk end were closed above; what is left has no patches
            while (!w0.done) close_doc(c, w0, 0);
            while (!w1.done) close_doc(c, w1, 1);
        }
        prof[6] = clock64() - _tstart;
        if (lane == 0) for (int i = 0; i < 8; ++i) g_prof[(blockIdx.x * 16 + warp) * 8 + i] = prof[i];
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// tok2q[t] = query that owns token row t (binary search in q_off).
__global__ void tok2q_kernel(const int32_t* __restrict__ q_off, int nq, int n_qtok, int32_t* __restrict__ tok2q) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_qtok) return;
    int lo = 0, hi = nq;  // largest q with q_off[q] <= t (empty queries share an offset: the last one owns the row)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(q_off + mid) <= t) lo = mid; else hi = mid;
    }
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

extern "C" int mmxp_get_prof(unsigned long long* host) { return (int)cudaMemcpyFromSymbol(host, mmx::maxsim::g_prof, sizeof(unsigned long long)*148*16*8); }
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
    MMX_CHECK_CUDA(cudaMemset2DAsync(scores, (size_t)ld_scores * sizeof(float), 0, (size_t)nd * sizeof(float),
                                     (size_t)nq, stream));
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
    // Chunk of whole documents per work item: small enough to stay L2-resident while all row-tile pairs sweep it,
    // and numerous enough that the persistent grid is balanced (>= 32 items per SM when the problem allows).
    const long long avg_doc_bytes = (n_dtok * (long long)DIM * 2 + nd - 1) / nd;
    long long dpc = L2_CHUNK_BYTES / (avg_doc_bytes > 0 ? avg_doc_bytes : 1);
    const long long want_chunks = ((long long)sms * 32 + p.n_mpairs - 1) / p.n_mpairs;
    const long long dpc_balance = (nd + want_chunks - 1) / want_chunks;
    if (dpc > dpc_balance) dpc = dpc_balance;
    if (dpc < 1) dpc = 1;
    if (dpc > nd) dpc = nd;
    p.docs_per_chunk = (int)dpc;
    p.n_chunks = (int)((nd + dpc - 1) / dpc);
    MMX_REQUIRE((long long)p.n_mpairs * p.n_chunks < (1ll << 31), "too many work items");

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

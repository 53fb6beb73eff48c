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
// Warp roles (12 warps):
//   0      TMA producer (one elected thread)
//   1      MMA issuer (one elected thread, tcgen05.mma.cta_group::1.kind::f16)
//   2      TMEM allocator
//   3      reducer: turns the row maxima of a finished document into per-query sums and stores them
//   4..11  drain warps. TMEM lane == query-token row; warps e and e+4 share a lane quarter and split the 256
//          columns of the SAME accumulator in halves, so an accumulator is back with the tensor core after ~300 clk.
//          Per 32-column chunk: tcgen05.ld.32x32b.x32 + an FMNMX3 tree, software-pipelined.
// Columns are global patch indices; document boundaries come from d_off. When a document closes, the drain warps
// only park their 128 row maxima (per column half) in a shared-memory slot and arrive on an mbarrier; the reducer
// warp takes max(half0, half1) per row, sums each query segment with a fixed shuffle tree (deterministic) and stores
// one float per (query, doc) — the drain warps never wait for that work. A query whose tokens straddle two row
// tiles gets its tail part in `carry[tile, doc]`, folded in by maxsim_fixup_kernel in tile order — no floating-point
// atomics anywhere.
//
// SHORT-document variant (template flag; the role-swapped I2T orientation of v28 :375-386 / main.py:622, where the
// "documents" are captions of ~20 tokens and a document closes every ~20 accumulator columns): the documents are
// first re-laid out with every length rounded up to a multiple of 4 rows, at least 16, by repeating a row of the same
// document (a duplicate cannot change a maximum; an empty document becomes zero rows, whose maxima are 0 like the
// reference's zero padding). Boundaries then fall on 4-column groups and are the same for every thread, so the drain
// is branch-uniform: 8 group maxima per 32-column chunk, a running maximum, and one shared-memory store per closed
// document into M[accumulator][slot][row] (at most 16 documents end inside one 256-column tile). Warps 4-7 drain
// accumulator 0 and warps 8-11 accumulator 1; every thread walks the columns [0,128) and [128,256) of its row as two
// interleaved chains (instruction-level parallelism for the latency-bound running maximum) and joins them at the end
// of the tile. Two reducer warps (2 and 3,
// one per accumulator) sum the rows of a whole tile's documents at once, two lanes per document, instead of one
// latency-bound shuffle tree per document.
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cstdlib>

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
constexpr int NUM_THREADS = 384;
constexpr int TMEM_COLS = 512;
constexpr int HALF_COLS = BN / 2;                // columns of one accumulator drained by one warp
constexpr int NSLOT = 4;                         // parked-document slots per accumulator
constexpr long long L2_CHUNK_BYTES = 24ll << 20; // document chunk kept hot in the 126 MB L2
constexpr int SHORT_ALIGN = 4;                   // SHORT variant: padded document lengths are multiples of this ...
constexpr int SHORT_MIN = 16;                    // ... and at least this, so at most BN / SHORT_MIN documents end per tile
constexpr int MSLOTS = BN / SHORT_MIN;           // 16
constexpr int MSTRIDE = BM + 1;                  // row pitch of M in floats: slot * 129 + row is conflict-free both ways
constexpr uint32_t MBUF_BYTES = 2 * MSLOTS * MSTRIDE * 4;  // one M buffer (both accumulators)
constexpr long long SHORT_AVG_ROWS = 256;        // documents shorter than this on average take the SHORT variant

struct Params {
    const int32_t* q_off;   // [nq+1] token offsets
    const int32_t* d_off;   // [nd+1] patch offsets
    const int32_t* tok2q;   // [n_qtok] owning query of each token row
    const uint8_t* d_zero;  // nullable [nd]: 1 = a zero-padded row takes part in this doc's max (pad_sequence)
    float* scores;          // [nq, ld_scores]
    long long ld_scores;
    float* carry;           // [n_mtiles, nd] partial sums of queries continued from the previous row tile
    const uint32_t* ends;   // SHORT: bit g = a document ends after 4-row group g of the padded document matrix
    int nq, nd, n_qtok, n_dtok;
    int n_mtiles, n_mpairs, n_chunks, docs_per_chunk;
};

struct __align__(1024) Smem {
    uint8_t a[2][A_TILE_BYTES];
    uint8_t b[NSTAGE][B_STAGE_BYTES];
    union {
        float stage[2][NSLOT][2][BM];  // [accumulator][slot][column half][row] row maxima of a closed document
        // SHORT: [buffer][accumulator][document slot][row] row maxima of a tile's documents. Two buffers per
        // accumulator: the reducer sums tile t while the drain warps already fill tile t + 1 (with one buffer the two
        // were serialised, T_reduce + T_drain = 2900 clk per 2048 clk of MMA: tensor pipe 70 %)
        float M[2][2][MSLOTS][MSTRIDE];
    };
    uint8_t seg_start[2][BM + 4];  // reducer-private: query segments of the two row tiles (row boundaries 0..128)
    int seg_q[2][BM];
    uint64_t a_full, a_free;
    uint64_t b_full[NSTAGE], b_empty[NSTAGE];
    uint64_t pb_full[4], pb_empty[4];  // CTA-pair kernel: 4 half-size B stages
    uint64_t t_full4[4], t_empty4[4];  // CTA-pair kernel: four 128-column accumulator slots [accumulator][column half]
    uint64_t t_full[2], t_empty[2];
    uint64_t st_full[2][NSLOT], st_empty[2][NSLOT];
    uint64_t batch_full[2][2], batch_empty[2][2];  // SHORT, [accumulator][buffer]: M written by the 4 drain warps / consumed by the reducer
    uint32_t tmem_base;
};

static_assert(sizeof(Smem) + 1024 <= 227 * 1024, "Smem (+ alignment slack) exceeds the 227 KB a CTA can opt into");

struct ItemGeom {
    int mp, d0, d1, p0, p1, ntiles;
};

__device__ __forceinline__ ItemGeom item_geom(const Params& p, int item) {
    ItemGeom g;
    g.mp = item % p.n_mpairs;
    const int ch = item / p.n_mpairs;
    g.d0 = ch * p.docs_per_chunk;
    g.d1 = min(p.nd, g.d0 + p.docs_per_chunk);
    g.p0 = __ldg(p.d_off + g.d0);
    g.p1 = __ldg(p.d_off + g.d1);
    g.ntiles = (g.p1 - g.p0 + BN - 1) / BN;
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

// ---------------------------------------------------------------------------------------------- drain warps
// Per-thread walk along the patch axis for one accumulator (= one row tile): which document the next column
// belongs to, and the running row maximum of that document over this thread's column half.
struct Walk {
    int cur_doc, cur_end, next_end;  // next_end = end of cur_doc+1, prefetched so a boundary never waits on HBM
    float m;
    int done;
    uint32_t nposted;  // documents parked so far (slot = nposted % NSLOT)
    uint32_t use;      // how many times this accumulator has been drained (mbarrier phase)
};

struct EpiCtx {
    Smem* s;
    const Params* p;
    int half, lane, row, d1;
    // CTA-pair kernel: the accumulators are handed back to the LEADER CTA's MMA thread — shared::cluster addresses of
    // its t_empty barriers (0 = this CTA's own barrier, the one-CTA kernels)
    uint32_t t_empty_cluster[2];
    uint64_t* t_full[2];  // "accumulator a is complete" for the columns this warp drains
};

// cur_doc is complete for this thread's columns: park the row maximum for the reducer and move to the next doc.
__device__ __forceinline__ void close_doc(const EpiCtx& c, Walk& w, int g) {
    Smem& s = *c.s;
    const uint32_t slot = w.nposted % NSLOT, round = w.nposted / NSLOT;
    mbar_wait(&s.st_empty[g][slot], (round & 1) ^ 1);  // the reducer has consumed this slot's previous document
    s.stage[g][slot][c.half][c.row] = w.m;
    __syncwarp();
    if (c.lane == 0) mbar_arrive(&s.st_full[g][slot]);
    w.nposted++;
    w.m = -INFINITY;
    w.cur_doc++;
    w.cur_end = w.next_end;
    if (w.cur_doc >= c.d1) {
        w.done = 1;
    } else if (w.cur_doc + 1 < c.d1) {
        w.next_end = __ldg(c.p->d_off + w.cur_doc + 2);
    }
}

__device__ __forceinline__ float max3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));  // FMNMX3
    return r;
}

// max(m, v[0..32)): four independent FMNMX3 chains of depth 4 plus two combines (asm-pinned so that ptxas does not
// re-serialise the associative maxima into two long dependent chains).
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
    return fmaxf(max3(c0, c1, c2), c3);
}

// Fold 32 accumulator columns this thread owns (global patch indices [pc, pc+32)) into the walk; documents may end
// anywhere inside the chunk.
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

__device__ __forceinline__ void release_accumulator(const EpiCtx& c, Smem& s, int a, int lane) {
    // this warp's 128 columns are in registers: hand the accumulator back to the MMA warp
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
        if (c.t_empty_cluster[a] != 0u) mbar_arrive_cluster(c.t_empty_cluster[a]);
        else mbar_arrive(&s.t_empty[a]);
    }
}

// Drain this warp's 128 columns of accumulator `a` for the B tile that starts at patch `pt`.
__device__ __forceinline__ void drain_accumulator(const EpiCtx& c, Smem& s, Walk& w, int a, uint32_t t_lane, int pt) {
    mbar_wait(c.t_full[a], w.use & 1);
    w.use++;
    tc_fence_after();
    const uint32_t ta = t_lane + a * BN;
    const int pc0 = pt + c.half * HALF_COLS;
    uint32_t va[32], vb[32];
    tmem_ld_32x32(ta, va);
    if (!w.done && w.cur_end - pt > BN) {
        // interior tile (3 of 4 at 1030 patches per document): every column belongs to cur_doc — straight-line,
        // the next chunk's TMEM load is in flight while the current one goes through the FMNMX3 tree.
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
        release_accumulator(c, s, a, c.lane);
        w.m = max32(vb, m);
    } else {
        // a document (or the chunk of documents) ends inside this tile; closing a document costs a shared-memory
        // store and an mbarrier arrive, so the same pipeline shape still works
        tmem_ld_wait();
        tmem_ld_32x32(ta + 32, vb);
        if (c.half == 1) skip_to(c, w, a, pc0);  // documents that end inside the other half's columns
        fold_chunk_general(c, w, a, va, pc0);
        tmem_ld_wait();
        tmem_ld_32x32(ta + 64, va);
        fold_chunk_general(c, w, a, vb, pc0 + 32);
        tmem_ld_wait();
        tmem_ld_32x32(ta + 96, vb);
        fold_chunk_general(c, w, a, va, pc0 + 64);
        tmem_ld_wait();
        release_accumulator(c, s, a, c.lane);
        fold_chunk_general(c, w, a, vb, pc0 + 96);
        if (c.half == 0) skip_to(c, w, a, pt + BN);
    }
}

// ---------------------------------------------------------------------------------------------- reducer warp
// Query-segment table of one 128-row tile: seg_start[0..nseg] row boundaries, seg_q[s] query id (-1 = padding rows).
__device__ __forceinline__ int build_segments(Smem& s, const Params& p, int tb, int mt, int lane) {
    const int r0 = mt * BM;
    int nseg = 0;
#pragma unroll 1
    for (int k = 0; k < BM / 32; ++k) {
        const int row = k * 32 + lane;
        const int rg = r0 + row;
        const int q = rg < p.n_qtok ? __ldg(p.tok2q + rg) : -1;
        const int qprev = (row == 0) ? INT_MIN : (rg - 1 < p.n_qtok ? __ldg(p.tok2q + rg - 1) : -1);
        const bool head = q != qprev;
        const uint32_t hmask = __ballot_sync(0xffffffffu, head);
        if (head) {
            const int idx = nseg + __popc(hmask & ((1u << lane) - 1u));
            s.seg_start[tb][idx] = (uint8_t)row;
            s.seg_q[tb][idx] = q;
        }
        nseg += __popc(hmask);
    }
    if (lane == 0) s.seg_start[tb][nseg] = BM;
    __syncwarp();
    return nseg;
}

struct Reducer {
    int nseg, mt, cont_first;
    uint32_t nposted;
};

// One closed document of accumulator g: wait for the 8 drain warps, sum_r max(half0[r], half1[r]) per query segment.
__device__ __forceinline__ void reduce_doc(Smem& s, const Params& p, Reducer& r, int g, int doc, int lane) {
    const uint32_t slot = r.nposted % NSLOT, round = r.nposted / NSLOT;
    mbar_wait(&s.st_full[g][slot], round & 1);
    const bool zero = p.d_zero != nullptr && p.d_zero[doc] != 0;
    float v[BM / 32];
#pragma unroll
    for (int k = 0; k < BM / 32; ++k) {
        const int row = k * 32 + lane;
        float x = fmaxf(s.stage[g][slot][0][row], s.stage[g][slot][1][row]);
        if (x == -INFINITY) x = 0.0f;  // a document without patches: the reference pads it with zero rows -> 0
        v[k] = zero ? fmaxf(x, 0.0f) : x;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&s.st_empty[g][slot]);  // values are in registers: the slot can be reused
    r.nposted++;
#pragma unroll 1
    for (int si = 0; si < r.nseg; ++si) {
        const int a = s.seg_start[g][si], b = s.seg_start[g][si + 1];
        float sum = 0.0f;
#pragma unroll
        for (int k = 0; k < BM / 32; ++k) {
            const int row = k * 32 + lane;
            sum += (row >= a && row < b) ? v[k] : 0.0f;
        }
        sum = warp_sum(sum);
        if (lane == 0) {
            const int q = s.seg_q[g][si];
            if (q >= 0) {
                if (si == 0 && r.cont_first)
                    p.carry[(long long)r.mt * p.nd + doc] = sum;
                else
                    p.scores[(long long)q * p.ld_scores + doc] = sum;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- SHORT variant
// Which 4-column groups of the tile starting at (padded) row pt end a document: 64 bits, uniform over the CTA.
__device__ __forceinline__ uint64_t tile_ends(const Params& p, int pt, int p1) {
    const int g0 = pt >> 2;
    const int w = g0 >> 5, sh = g0 & 31;
    const uint32_t a = __ldg(p.ends + w), b = __ldg(p.ends + w + 1), c = __ldg(p.ends + w + 2);
    uint64_t m = (uint64_t)__funnelshift_r(a, b, sh) | ((uint64_t)__funnelshift_r(b, c, sh) << 32);
    const int valid = (p1 - pt) >> 2;  // groups of this tile that still belong to the chunk
    if (valid < 64) m &= (valid <= 0) ? 0ull : ((1ull << valid) - 1ull);
    return m;
}

__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

// 16 columns = 4 groups of 4 of EACH of two independent column chains (A: the tile's columns [0,128), B: [128,256));
// eA / eB = which of those groups end a document (the same for every thread of the CTA). addr = shared-space byte
// address of M[accumulator][next free slot][row] of the chain; a closed document's maximum is stored there and the
// address moves on one slot. Branch-free: one predicate per group drives the store, the reset of the running maximum
// and the address step. The chains are interleaved group by group: every chain is a serial FMNMX -> select -> FMNMX
// dependency through its running maximum, and one chain alone left the drain warp waiting on ALU latency for a third of its cycles (ncu: stall_wait 30 %,
// IPC 0.34 while draining). Two chains give the scheduler an independent instruction for every dependent one.
__device__ __forceinline__ void fold_short2(const uint32_t (&a)[16], const uint32_t (&b)[16], float& mA, float& mB,
                                            uint32_t eA, uint32_t eB, uint32_t& addrA, uint32_t& addrB) {
    float gA[4], gB[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        gA[k] = fmaxf(max3(__uint_as_float(a[4 * k]), __uint_as_float(a[4 * k + 1]), __uint_as_float(a[4 * k + 2])),
                      __uint_as_float(a[4 * k + 3]));
        gB[k] = fmaxf(max3(__uint_as_float(b[4 * k]), __uint_as_float(b[4 * k + 1]), __uint_as_float(b[4 * k + 2])),
                      __uint_as_float(b[4 * k + 3]));
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        mA = fmaxf(mA, gA[k]);
        mB = fmaxf(mB, gB[k]);
        asm volatile(
            "{\n\t.reg .pred p, q;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "setp.ne.b32 q, %5, 0;\n\t"
            "@p st.shared.f32 [%0], %1;\n\t"
            "@q st.shared.f32 [%2], %3;\n\t"
            "@p mov.b32 %1, 0xff800000;\n\t"
            "@q mov.b32 %3, 0xff800000;\n\t"
            "@p add.u32 %0, %0, %6;\n\t"
            "@q add.u32 %2, %2, %6;\n\t}"
            : "+r"(addrA), "+f"(mA), "+r"(addrB), "+f"(mB)
            : "r"(eA & (1u << k)), "r"(eB & (1u << k)), "n"(MSTRIDE * 4)
            : "memory");
    }
}

// SHORT reducer of accumulator a: per tile, two lanes per document (lane = slot + 16 * h; the halves take alternate
// blocks of 16 rows, which keeps the reads of M[slot][row] free of bank conflicts); every lane walks the same
// segment table and the partial sums are combined in a fixed order. Rows of a segment that fill whole 32-row pairs
// are summed without any range test; only the ragged edges of a segment (one tile in eight holds a query boundary
// when the queries are 1030-row images) take the predicated loop.
__device__ __forceinline__ float red_edge(uint32_t M, int r_lo, int r_hi, int h, float floor_v) {
    float acc = 0.0f;
    for (int r = r_lo; r < r_hi; ++r)
        if (((r >> 4) & 1) == h) acc += fmaxf(lds_f32(M + 4u * (uint32_t)r), floor_v);
    return acc;
}

__device__ __forceinline__ void short_reducer(Smem& s, const Params& p, int a, int lane, int n_items) {
    const int j = lane & (MSLOTS - 1), h = lane >> 4;
    const bool flags = p.d_zero != nullptr;
    uint32_t use = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const ItemGeom g = item_geom(p, item);
        const int nm = (g.mp * 2 + 1 < p.n_mtiles) ? 2 : 1;
        if (a >= nm) continue;
        const long long mt = g.mp * 2 + a;
        const int r0 = (int)mt * BM;
        const bool cont = (r0 > 0) && (r0 < p.n_qtok) && (__ldg(p.tok2q + r0 - 1) == __ldg(p.tok2q + r0));
        const int nseg = build_segments(s, p, a, (int)mt, lane);
        int first = g.d0;
        uint64_t ends_next = tile_ends(p, g.p0, g.p1);
        uint8_t zflag_next = (flags && first + j < p.nd) ? __ldg(p.d_zero + first + j) : (uint8_t)0;
        const uint32_t M0 = smem_u32(&s.M[0][a][j][0]);
        for (int t = 0; t < g.ntiles; ++t, ++use) {
            const uint64_t ends = ends_next;
            const uint8_t zflag = zflag_next;
            const int n = __popcll(ends);
            // next tile's boundary words and padding flags travel from L2 while this tile is reduced
            if (t + 1 < g.ntiles) {
                ends_next = tile_ends(p, g.p0 + (t + 1) * BN, g.p1);
                zflag_next = (flags && first + n + j < p.nd) ? __ldg(p.d_zero + first + n + j) : (uint8_t)0;
            }
            const bool act = j < n;
            const int doc = first + j;
            const float floor_v = (act && zflag) ? 0.0f : -INFINITY;  // max(x, -inf) = x: one code path for both
            const uint32_t M = M0 + (use & 1u) * MBUF_BYTES;
            mbar_wait(&s.batch_full[a][use & 1], (use >> 1) & 1);
#pragma unroll 1
            for (int si = 0; si < nseg; ++si) {
                const int ra = s.seg_start[a][si], rb = s.seg_start[a][si + 1];
                float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
                if (act) {
                    const int f0 = (ra + 31) & ~31, f1 = rb & ~31;  // whole 32-row pairs are [f0, f1)
                    if (f0 <= f1) {
                        a0 = red_edge(M, ra, f0, h, floor_v);
                        a1 = red_edge(M, f1, rb, h, floor_v);
#pragma unroll 1
                        for (int base = f0 + 16 * h; base < f1; base += 32) {
                            const uint32_t mb = M + 4u * (uint32_t)base;
                            if (flags) {
#pragma unroll
                                for (int i = 0; i < 16; i += 4) {
                                    a0 += fmaxf(lds_f32(mb + 4 * i), floor_v);
                                    a1 += fmaxf(lds_f32(mb + 4 * i + 4), floor_v);
                                    a2 += fmaxf(lds_f32(mb + 4 * i + 8), floor_v);
                                    a3 += fmaxf(lds_f32(mb + 4 * i + 12), floor_v);
                                }
                            } else {
#pragma unroll
                                for (int i = 0; i < 16; i += 4) {
                                    a0 += lds_f32(mb + 4 * i);
                                    a1 += lds_f32(mb + 4 * i + 4);
                                    a2 += lds_f32(mb + 4 * i + 8);
                                    a3 += lds_f32(mb + 4 * i + 12);
                                }
                            }
                        }
                    } else {
                        a0 = red_edge(M, ra, rb, h, floor_v);  // the segment lies inside one 32-row pair
                    }
                }
                float sum = (a0 + a1) + (a2 + a3);
                sum += __shfl_xor_sync(0xffffffffu, sum, 16);
                const int q = s.seg_q[a][si];
                if (act && h == 0 && q >= 0) {
                    if (si == 0 && cont)
                        p.carry[mt * p.nd + doc] = sum;
                    else
                        p.scores[(long long)q * p.ld_scores + doc] = sum;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s.batch_empty[a][use & 1]);
            first += n;
        }
    }
}

template <bool SHORT>
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
            mbar_init(&s.t_empty[i], SHORT ? 4 : 8);  // one arrival per drain warp of this accumulator
            for (int j = 0; j < 2; ++j) {
                mbar_init(&s.batch_full[i][j], 4);
                mbar_init(&s.batch_empty[i][j], 1);
            }
            for (int j = 0; j < NSLOT; ++j) {
                mbar_init(&s.st_full[i][j], 8);
                mbar_init(&s.st_empty[i][j], 1);
            }
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
    } else if ((warp == 2 || warp == 3) && SHORT) {
        // ------------------------------------------------------------------ SHORT reducers: warp 2 finishes
        // accumulator 0, warp 3 accumulator 1
        short_reducer(s, p, warp - 2, lane, n_items);
    } else if (warp == 3) {
        // ------------------------------------------------------------------ reducer: replays the drain warps'
        // sequence of closed documents (it depends on d_off only) and finishes each one
        Reducer r[2];
        r[0].nposted = r[1].nposted = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const ItemGeom g = item_geom(p, item);
            const int nm = (g.mp * 2 + 1 < p.n_mtiles) ? 2 : 1;
            int cur[2];
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const int mt = g.mp * 2 + a;
                const int r0 = mt * BM;
                r[a].mt = mt;
                r[a].cont_first = (r0 > 0) && (r0 < p.n_qtok) && (__ldg(p.tok2q + r0 - 1) == __ldg(p.tok2q + r0));
                r[a].nseg = (a < nm) ? build_segments(s, p, a, mt, lane) : 0;
                cur[a] = (a < nm) ? g.d0 : g.d1;
            }
            for (int t = 0; t < g.ntiles; ++t) {
                const int pend = g.p0 + (t + 1) * BN;
#pragma unroll
                for (int a = 0; a < 2; ++a)
                    while (cur[a] < g.d1 && __ldg(p.d_off + cur[a] + 1) <= pend) {
                        reduce_doc(s, p, r[a], a, cur[a], lane);
                        cur[a]++;
                    }
            }
#pragma unroll
            for (int a = 0; a < 2; ++a)
                while (cur[a] < g.d1) {  // trailing documents without patches
                    reduce_doc(s, p, r[a], a, cur[a], lane);
                    cur[a]++;
                }
        }
    } else if (warp >= 4 && SHORT) {
        // ------------------------------------------------------------------ SHORT drain: warps 4-7 own accumulator 0,
        // warps 8-11 accumulator 1 (all 256 columns); document boundaries are uniform 4-column group ends
        const int a = (warp - 4) >> 2;
        const int row = (warp & 3) * 32 + lane;
        const uint32_t ta = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + a * BN;
        const uint32_t Mrow0 = smem_u32(&s.M[0][a][0][0]) + 4u * (uint32_t)row;
        uint32_t use = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const ItemGeom g = item_geom(p, item);
            const int nm = (g.mp * 2 + 1 < p.n_mtiles) ? 2 : 1;
            if (a >= nm) continue;
            float m = -INFINITY;
            uint64_t ends_next = tile_ends(p, g.p0, g.p1);
#pragma unroll 1
            for (int t = 0; t < g.ntiles; ++t, ++use) {
                const uint64_t ends = ends_next;
                // the next tile's boundary words travel from L2 while this tile is drained
                if (t + 1 < g.ntiles) ends_next = tile_ends(p, g.p0 + (t + 1) * BN, g.p1);
                const uint32_t elo = (uint32_t)ends, ehi = (uint32_t)(ends >> 32);
                mbar_wait(&s.t_full[a], use & 1);
                tc_fence_after();
                // two column chains per thread: A walks the tile's columns [0,128), B [128,256), 16 columns of each
                // per step; B's documents take the slots after A's (their number is known from the boundary word)
                // NBUF register buffers per chain: the loads run NBUF - 1 steps ahead of the folds, so the accumulator
                // goes back to the tensor core after 9 - NBUF of the eight folds
                constexpr int NBUF = 4;
                uint32_t va[NBUF][16], vb[NBUF][16];
#pragma unroll
                for (int i = 0; i < NBUF - 1; ++i) {
                    tmem_ld_32x16(ta + 16 * i, va[i]);
                    tmem_ld_32x16(ta + 128 + 16 * i, vb[i]);
                }
                // the reducer is done with the tile that used this buffer two tiles ago
                mbar_wait(&s.batch_empty[a][use & 1], ((use >> 1) & 1) ^ 1);
                const uint32_t Mrow = Mrow0 + (use & 1u) * MBUF_BYTES;
                uint32_t addrA = Mrow;
                const uint32_t addrB0 = Mrow + (uint32_t)__popc(elo) * (MSTRIDE * 4);
                uint32_t addrB = addrB0;
                float mB = -INFINITY;
#pragma unroll
                for (int st = 0; st < 8; ++st) {
                    // everything issued so far has landed (the later steps were issued at least a whole fold ago)
                    if (st + NBUF - 1 <= 8) tmem_ld_wait();
                    if (st + NBUF - 1 < 8) {
                        tmem_ld_32x16(ta + 16 * (st + NBUF - 1), va[(st + NBUF - 1) % NBUF]);
                        tmem_ld_32x16(ta + 128 + 16 * (st + NBUF - 1), vb[(st + NBUF - 1) % NBUF]);
                    } else if (st + NBUF - 1 == 8) {
                        // all 256 columns of this row are in registers
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&s.t_empty[a]);
                    }
                    fold_short2(va[st % NBUF], vb[st % NBUF], m, mB, (elo >> (4 * st)) & 0xfu, (ehi >> (4 * st)) & 0xfu,
                                addrA, addrB);
                }
                // join the chains: the document that was open at column 128 belongs to B's first closed document, or,
                // if B closed none, to the maximum carried into the next tile
                if (ehi != 0u) {
                    sts_f32(addrB0, fmaxf(lds_f32(addrB0), m));
                    m = mB;
                } else {
                    m = fmaxf(m, mB);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&s.batch_full[a][use & 1]);
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ drain warps (8 warps, 256 threads)
        EpiCtx c;
        c.s = &s;
        c.p = &p;
        c.half = (warp - 4) >> 2;  // which 128 columns of every accumulator this warp drains
        c.t_empty_cluster[0] = c.t_empty_cluster[1] = 0u;
        c.t_full[0] = &s.t_full[0];
        c.t_full[1] = &s.t_full[1];
        c.lane = lane;
        c.row = (warp & 3) * 32 + lane;  // TMEM lane quarter = warp % 4
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + c.half * HALF_COLS;
        Walk w0, w1;
        w0.nposted = w1.nposted = 0;
        w0.use = w1.use = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const ItemGeom g = item_geom(p, item);
            const int nm = (g.mp * 2 + 1 < p.n_mtiles) ? 2 : 1;
            c.d1 = g.d1;
            const int e0 = __ldg(p.d_off + g.d0 + 1);
            const int e1 = (g.d0 + 1 < g.d1) ? __ldg(p.d_off + g.d0 + 2) : e0;
            w0.cur_doc = w1.cur_doc = g.d0;
            w0.cur_end = w1.cur_end = e0;
            w0.next_end = w1.next_end = e1;
            w0.m = w1.m = -INFINITY;
            w0.done = (g.d0 >= g.d1);
            w1.done = (g.d0 >= g.d1) || (nm < 2);
#pragma unroll 1
            for (int t = 0; t < g.ntiles; ++t) {
                const int pt = g.p0 + t * BN;
                drain_accumulator(c, s, w0, 0, t_lane, pt);
                if (nm == 2) drain_accumulator(c, s, w1, 1, t_lane, pt);
            }
            // documents that end exactly at the chunk end were closed above; what is left has no patches
            while (!w0.done) close_doc(c, w0, 0);
            while (!w1.done) close_doc(c, w1, 1);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------- CTA-pair variant
// The same pipeline on a cluster of two CTAs (one TPC) with tcgen05 cta_group::2: the pair works on ONE chunk of
// documents and FOUR 128-row query tiles (two per CTA). Every B tile of 256 patches is split — each CTA's TMA fetches
// 128 of them into its own shared memory — and one `tcgen05.mma.cta_group::2` (M = 256: 128 rows of each CTA, N = 256)
// issued by the leader CTA fills the same accumulator in both CTAs' TMEM: every CTA pulls half the B bytes from L2 and
// its tensor core reads half the B rows from its own shared memory, the other half arrives from the peer. Everything
// behind the accumulators (drain warps, reducer, carry) is per CTA and unchanged.
//   barriers (all in the LEADER's shared memory unless noted): a_full / pb_full[stage] count the TMA bytes of BOTH CTAs
//   (the peer's loads complete on the leader's barrier); the accumulators are FOUR slots of 128 columns
//   [row tile][column half], each drained by 4 + 4 warps of the two CTAs that arrive on the leader's t_empty4[slot]
//   (the peer's over the cluster); tcgen05.commit multicasts t_full4[slot], pb_empty[stage] and a_free to both CTAs.
// With half-size B stages the ring is 4 deep in the same shared memory.
constexpr int PSTAGE = 4;
constexpr int BP_ROWS = BN / 2;                    // patches per CTA and B tile
constexpr int BP_HALF_BYTES = BP_ROWS * 128;       // one 64-wide K half: 16 KB
constexpr int BP_STAGE_BYTES = 2 * BP_HALF_BYTES;  // 32 KB
static_assert(PSTAGE * BP_STAGE_BYTES == NSTAGE * B_STAGE_BYTES, "the pair kernel reuses the B ring's shared memory");

struct PairGeom {
    int mp, nm, nm_pair, d0, d1, p0, p1, ntiles;
};
__device__ __forceinline__ PairGeom pair_geom(const Params& p, int item, int n_mpp, int rank) {
    PairGeom g;
    const int mpp = item % n_mpp, ch = item / n_mpp;
    g.mp = 2 * mpp + rank;
    g.nm = max(0, min(2, p.n_mtiles - 2 * g.mp));          // this CTA's valid row tiles (the last pair may be ragged)
    g.nm_pair = max(0, min(2, p.n_mtiles - 4 * mpp));      // the leader's: MMAs are issued for these accumulators
    g.d0 = ch * p.docs_per_chunk;
    g.d1 = min(p.nd, g.d0 + p.docs_per_chunk);
    g.p0 = __ldg(p.d_off + g.d0);
    g.p1 = __ldg(p.d_off + g.d1);
    g.ntiles = (g.p1 - g.p0 + BN - 1) / BN;
    return g;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
maxsim_pair_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_d,
                   const __grid_constant__ Params p) {
    extern __shared__ uint8_t smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    uint8_t* const bring = &s.b[0][0];  // [PSTAGE][2 K halves][128 rows x 128 B]

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_d);
    }
    if (warp == 1 && elect_one()) {
        mbar_init(&s.a_full, 1);
        mbar_init(&s.a_free, 1);
        for (int i = 0; i < PSTAGE; ++i) {
            mbar_init(&s.pb_full[i], 1);
            mbar_init(&s.pb_empty[i], 1);
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(&s.t_full4[i], 1);
            mbar_init(&s.t_empty4[i], 8);  // 4 drain warps of each CTA (only the leader's copy is used)
        }
        for (int i = 0; i < 2; ++i) {
            for (int j = 0; j < NSLOT; ++j) {
                mbar_init(&s.st_full[i][j], 8);
                mbar_init(&s.st_empty[i][j], 1);
            }
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc_2cta(&s.tmem_base, TMEM_COLS);
        tmem_relinquish_2cta();
    }
    tc_fence_before();
    cluster_sync_all();  // both CTAs' barriers are initialised before anything arrives on them remotely
    tc_fence_after();
    const uint32_t tmem_base = s.tmem_base;
    const int n_mpp = (p.n_mpairs + 1) / 2;
    const int n_items = n_mpp * p.n_chunks;
    const int cl = blockIdx.x >> 1, ncl = gridDim.x >> 1;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (one per CTA)
        if (elect_one()) {
            const uint32_t a_full_leader = mapa_shared(smem_u32(&s.a_full), 0);
            uint32_t it = 0, bcount = 0;
            for (int item = cl; item < n_items; item += ncl, ++it) {
                const PairGeom g = pair_geom(p, item, n_mpp, rank);
                if (it > 0) mbar_wait(&s.a_free, (it - 1) & 1);
                if (leader) mbar_arrive_expect_tx(&s.a_full, 2 * 2 * A_TILE_BYTES);
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        tma_load_2d_2cta(s.a[m] + h * A_HALF_BYTES, &tmap_q, a_full_leader, h * 64, (g.mp * 2 + m) * BM);
                for (int t = 0; t < g.ntiles; ++t, ++bcount) {
                    const uint32_t st = bcount % PSTAGE, ph = (bcount / PSTAGE) & 1;
                    mbar_wait(&s.pb_empty[st], ph ^ 1);
                    if (leader) mbar_arrive_expect_tx(&s.pb_full[st], 2 * BP_STAGE_BYTES);
                    const uint32_t full_leader = mapa_shared(smem_u32(&s.pb_full[st]), 0);
                    // this CTA's 64 + 64 patches of the tile: rows [0, 64) of its stage pair up with the peer's rows
                    // [0, 64) into the tile's first 128 columns, rows [64, 128) into the second 128
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int part = 0; part < 2; ++part)
                            tma_load_2d_2cta(bring + st * BP_STAGE_BYTES + h * BP_HALF_BYTES + part * (BP_ROWS / 2) * 128,
                                             &tmap_d, full_leader, h * 64,
                                             g.p0 + t * BN + part * (BN / 2) + (int)rank * (BP_ROWS / 2));
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (one thread of the leader CTA)
        if (leader && elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN / 2);
            uint32_t it = 0, bcount = 0, use[4] = {0, 0, 0, 0};
            for (int item = cl; item < n_items; item += ncl, ++it) {
                const PairGeom g = pair_geom(p, item, n_mpp, rank);
                mbar_wait(&s.a_full, it & 1);
                tc_fence_after();
                for (int t = 0; t < g.ntiles; ++t, ++bcount) {
                    const uint32_t st = bcount % PSTAGE, ph = (bcount / PSTAGE) & 1;
                    mbar_wait(&s.pb_full[st], ph);
                    tc_fence_after();
                    // four accumulator slots of 128 columns: a slot is written again three MMA groups (~1500 clk) after
                    // it completed, which covers the drain plus the peer's remote release
                    for (int m = 0; m < g.nm_pair; ++m)
#pragma unroll
                        for (int hf = 0; hf < 2; ++hf) {
                            const int slot = 2 * m + hf;
                            mbar_wait(&s.t_empty4[slot], (use[slot] & 1) ^ 1);
                            tc_fence_after();
#pragma unroll
                            for (int k = 0; k < DIM / 16; ++k) {
                                const uint32_t a_addr = smem_u32(s.a[m] + (k >> 2) * A_HALF_BYTES) + (k & 3) * 32;
                                const uint32_t b_addr = smem_u32(bring + st * BP_STAGE_BYTES + (k >> 2) * BP_HALF_BYTES +
                                                                 hf * (BP_ROWS / 2) * 128) + (k & 3) * 32;
                                umma_bf16_ss_2cta(tmem_base + m * BN + hf * (BN / 2), umma_desc_kmajor_sw128(a_addr),
                                                  umma_desc_kmajor_sw128(b_addr), idesc, k > 0);
                            }
                            umma_commit_2cta(&s.t_full4[slot]);
                            use[slot]++;
                        }
                    umma_commit_2cta(&s.pb_empty[st]);
                }
                umma_commit_2cta(&s.a_free);
            }
        }
    } else if (warp == 3) {
        // ------------------------------------------------------------------ reducer (per CTA, as in the one-CTA kernel)
        Reducer r[2];
        r[0].nposted = r[1].nposted = 0;
        for (int item = cl; item < n_items; item += ncl) {
            const PairGeom g = pair_geom(p, item, n_mpp, rank);
            int cur[2];
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const int mt = g.mp * 2 + a;
                const int r0 = mt * BM;
                r[a].mt = mt;
                r[a].cont_first = (r0 > 0) && (r0 < p.n_qtok) && (__ldg(p.tok2q + r0 - 1) == __ldg(p.tok2q + r0));
                r[a].nseg = (a < g.nm) ? build_segments(s, p, a, mt, lane) : 0;
                cur[a] = (a < g.nm) ? g.d0 : g.d1;
            }
            for (int t = 0; t < g.ntiles; ++t) {
                const int pend = g.p0 + (t + 1) * BN;
#pragma unroll
                for (int a = 0; a < 2; ++a)
                    while (cur[a] < g.d1 && __ldg(p.d_off + cur[a] + 1) <= pend) {
                        reduce_doc(s, p, r[a], a, cur[a], lane);
                        cur[a]++;
                    }
            }
#pragma unroll
            for (int a = 0; a < 2; ++a)
                while (cur[a] < g.d1) {  // trailing documents without patches
                    reduce_doc(s, p, r[a], a, cur[a], lane);
                    cur[a]++;
                }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ drain warps (8 warps per CTA)
        EpiCtx c;
        c.s = &s;
        c.p = &p;
        c.half = (warp - 4) >> 2;
        c.lane = lane;
        c.row = (warp & 3) * 32 + lane;
        c.t_empty_cluster[0] = mapa_shared(smem_u32(&s.t_empty4[0 + c.half]), 0);
        c.t_empty_cluster[1] = mapa_shared(smem_u32(&s.t_empty4[2 + c.half]), 0);
        c.t_full[0] = &s.t_full4[0 + c.half];
        c.t_full[1] = &s.t_full4[2 + c.half];
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + c.half * HALF_COLS;
        Walk w0, w1;
        w0.nposted = w1.nposted = 0;
        w0.use = w1.use = 0;
        for (int item = cl; item < n_items; item += ncl) {
            const PairGeom g = pair_geom(p, item, n_mpp, rank);
            c.d1 = g.d1;
            const int e0 = __ldg(p.d_off + g.d0 + 1);
            const int e1 = (g.d0 + 1 < g.d1) ? __ldg(p.d_off + g.d0 + 2) : e0;
            w0.cur_doc = w1.cur_doc = g.d0;
            w0.cur_end = w1.cur_end = e0;
            w0.next_end = w1.next_end = e1;
            w0.m = w1.m = -INFINITY;
            // a row tile beyond the matrix (ragged last pair) is still drained — the MMA wrote it and waits for its
            // release — but closes no documents
            w0.done = (g.d0 >= g.d1) || (g.nm < 1);
            w1.done = (g.d0 >= g.d1) || (g.nm < 2);
#pragma unroll 1
            for (int t = 0; t < g.ntiles; ++t) {
                const int pt = g.p0 + t * BN;
                if (g.nm_pair >= 1) drain_accumulator(c, s, w0, 0, t_lane, pt);
                if (g.nm_pair == 2) drain_accumulator(c, s, w1, 1, t_lane, pt);
            }
            while (!w0.done) close_doc(c, w0, 0);
            while (!w1.done) close_doc(c, w1, 1);
        }
    }

    tc_fence_before();
    cluster_sync_all();  // the peer may still be reading this CTA's shared memory / writing its TMEM through the pair MMA
    if (warp == 2) tmem_dealloc_2cta(tmem_base, TMEM_COLS);
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

// ---------------------------------------------------------------------------------------------- SHORT prep
__host__ __device__ __forceinline__ int short_padded_len(int L) {
    const int r = (L + SHORT_ALIGN - 1) & ~(SHORT_ALIGN - 1);
    return r < SHORT_MIN ? SHORT_MIN : r;
}

// pd_off[d] = exclusive prefix sum of the padded document lengths (one block; nd / 1024 sequential rounds)
__global__ void __launch_bounds__(1024)
short_offsets_kernel(const int32_t* __restrict__ d_off, int nd, int32_t* __restrict__ pd_off) {
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < nd; base += 1024) {
        const int d = base + t;
        const int len = d < nd ? short_padded_len(d_off[d + 1] - d_off[d]) : 0;
        int incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += u;
            }
            warp_tot[lane] = w;  // inclusive totals of the warps
        }
        __syncthreads();
        const int before = carry_s + (warp > 0 ? warp_tot[warp - 1] : 0) + incl - len;
        if (d < nd) pd_off[d] = before;
        __syncthreads();
        if (t == 1023) carry_s = before + len;
        __syncthreads();
    }
    if (t == 0) pd_off[nd] = carry_s;
}

// one warp per document: copy its rows into the padded matrix, fill the tail with copies of its first row (zeros
// for an empty document), and mark the group that ends it
__global__ void __launch_bounds__(256)
short_pad_docs_kernel(const uint2* __restrict__ d_tok, const int32_t* __restrict__ d_off,
                      const int32_t* __restrict__ pd_off, int nd, uint2* __restrict__ out, uint32_t* __restrict__ ends) {
    const int lane = threadIdx.x & 31;
    const int d = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (d >= nd) return;
    const long long src = d_off[d], L = d_off[d + 1] - src;
    const long long dst = pd_off[d], Lp = pd_off[d + 1] - dst;
    for (long long r = 0; r < Lp; ++r) {
        uint2 v = make_uint2(0u, 0u);
        if (L > 0) v = d_tok[(src + (r < L ? r : 0)) * 32 + lane];  // a row is 128 bf16 = 32 x 8 bytes
        out[(dst + r) * 32 + lane] = v;
    }
    if (lane == 0) {
        const long long g = ((dst + Lp) >> 2) - 1;
        atomicOr(ends + (g >> 5), 1u << (g & 31));
    }
}

}  // namespace maxsim
}  // namespace mmx

using namespace mmx;
using namespace mmx::maxsim;

namespace {
struct WsPlan {
    bool shortdocs;
    long long pad_rows;  // upper bound of the padded document matrix
    size_t off_tok2q, off_carry, off_pdoff, off_ends, off_pad, total;
};
WsPlan plan_ws(int64_t n_qtok, int64_t nd, int64_t n_dtok) {
    WsPlan w{};
    auto up = [](size_t x) { return (x + 1023) & ~size_t(1023); };
    const int64_t n_mtiles = (n_qtok + BM - 1) / BM;
    w.shortdocs = nd > 0 && n_dtok < SHORT_AVG_ROWS * nd;
    size_t o = 1024;
    w.off_tok2q = o; o += up((size_t)n_qtok * 4);
    w.off_carry = o; o += up((size_t)n_mtiles * (size_t)nd * 4);
    if (w.shortdocs) {
        w.pad_rows = n_dtok + (long long)SHORT_MIN * nd + BN;
        w.off_pdoff = o; o += up((size_t)(nd + 1) * 4);
        w.off_ends = o;  o += up((size_t)(w.pad_rows / 4 / 32 + 4) * 4);
        w.off_pad = o;   o += up((size_t)w.pad_rows * DIM * 2);
    }
    w.total = o;
    return w;
}
}  // namespace

extern "C" int mmx_maxsim_workspace_bytes(int64_t n_qtok, int64_t nd, int64_t n_dtok, size_t* bytes) {
    MMX_REQUIRE(bytes != nullptr && n_qtok >= 0 && nd >= 0 && n_dtok >= 0, "bad argument");
    *bytes = plan_ws(n_qtok, nd, n_dtok).total;
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
    const WsPlan wp = plan_ws(n_qtok, nd, n_dtok);
    MMX_REQUIRE(workspace != nullptr && workspace_bytes >= wp.total, "workspace too small: need %zu bytes, got %zu",
                wp.total, workspace_bytes);
    MMX_REQUIRE(!wp.shortdocs || wp.pad_rows < (1ll << 31) - 512, "padded document matrix exceeds int32 offsets");

    Params p;
    p.q_off = q_off;
    p.d_off = d_off;
    uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023));
    int32_t* tok2q = reinterpret_cast<int32_t*>(ws + wp.off_tok2q - 1024);
    p.tok2q = tok2q;
    p.carry = reinterpret_cast<float*>(ws + wp.off_carry - 1024);
    p.ends = nullptr;
    const void* d_mat = d_tok;
    long long d_rows = n_dtok;
    if (wp.shortdocs) {
        // re-lay the documents out with lengths rounded up to a multiple of 4 (>= 16) rows
        int32_t* pd_off = reinterpret_cast<int32_t*>(ws + wp.off_pdoff - 1024);
        uint32_t* ends = reinterpret_cast<uint32_t*>(ws + wp.off_ends - 1024);
        void* pad = ws + wp.off_pad - 1024;
        MMX_CHECK_CUDA(cudaMemsetAsync(ends, 0, (size_t)(wp.pad_rows / 4 / 32 + 4) * 4, stream));
        short_offsets_kernel<<<1, 1024, 0, stream>>>(d_off, (int)nd, pd_off);
        short_pad_docs_kernel<<<(unsigned)((nd + 7) / 8), 256, 0, stream>>>(
            reinterpret_cast<const uint2*>(d_tok), d_off, pd_off, (int)nd, reinterpret_cast<uint2*>(pad), ends);
        MMX_CHECK_CUDA(cudaGetLastError());
        p.d_off = pd_off;
        p.ends = ends;
        d_mat = pad;
        d_rows = wp.pad_rows;
    }
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
    long long chunk_bytes = L2_CHUNK_BYTES;
    if (const char* e = getenv("MMX_L2_CHUNK_MB")) {  // tuning knob (dev/tune_chunk.sh); unset in production
        const long long mb = atoll(e);
        if (mb >= 1 && mb <= 4096) chunk_bytes = mb << 20;
    }
    long long dpc = chunk_bytes / (avg_doc_bytes > 0 ? avg_doc_bytes : 1);
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
    if (make_tmap_bf16_sw128(&td, d_mat, (uint64_t)(d_rows > 0 ? d_rows : 1), DIM, DIM, BN) != 0) return -4;

    tok2q_kernel<<<(unsigned)((n_qtok + 255) / 256), 256, 0, stream>>>(q_off, (int)nq, (int)n_qtok, tok2q);
    MMX_CHECK_CUDA(cudaGetLastError());

    const size_t smem_bytes = sizeof(Smem) + 1024;
    static const char attr_set_tag = 0;
    if (first_use_on_device(&attr_set_tag)) {
        MMX_CHECK_CUDA(cudaFuncSetAttribute(maxsim_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        MMX_CHECK_CUDA(cudaFuncSetAttribute(maxsim_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    }
    const long long n_items = (long long)p.n_mpairs * p.n_chunks;
    const int grid = (int)(n_items < sms ? n_items : sms);
    // CTA-pair variant (long documents): two CTAs of a TPC share every B tile through tcgen05 cta_group::2
    static const int pair_mode = [] {  // on by default; MMX_MAXSIM_PAIR=0 selects the one-CTA kernel (dev/check_pair.py)
        const char* e = getenv("MMX_MAXSIM_PAIR");
        return e != nullptr ? atoi(e) : 1;
    }();
    const long long n_items_pair = (long long)((p.n_mpairs + 1) / 2) * p.n_chunks;
    if (wp.shortdocs) {
        maxsim_kernel<true><<<grid, NUM_THREADS, smem_bytes, stream>>>(tq, td, p);
    } else if (pair_mode != 0 && p.n_mtiles >= 4 && n_items_pair >= sms / 2) {
        CUtensorMap td_half;
        if (make_tmap_bf16_sw128(&td_half, d_mat, (uint64_t)(d_rows > 0 ? d_rows : 1), DIM, DIM, BP_ROWS / 2) != 0) return -4;
        static const char pair_attr_tag = 0;
        if (first_use_on_device(&pair_attr_tag))
            MMX_CHECK_CUDA(cudaFuncSetAttribute(maxsim_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        const int grid_pair = (int)std::min<long long>(2 * n_items_pair, sms & ~1);
        maxsim_pair_kernel<<<grid_pair, NUM_THREADS, smem_bytes, stream>>>(tq, td_half, p);
    } else {
        maxsim_kernel<false><<<grid, NUM_THREADS, smem_bytes, stream>>>(tq, td, p);
    }
    MMX_CHECK_CUDA(cudaGetLastError());

    if (p.n_mtiles > 1) {
        dim3 fg((unsigned)(p.n_mtiles - 1), (unsigned)((nd + 255) / 256));
        maxsim_fixup_kernel<<<fg, 256, 0, stream>>>(tok2q, q_off, p.carry, scores, ld_scores, (int)n_qtok,
                                                    p.n_mtiles, (int)nd);
        MMX_CHECK_CUDA(cudaGetLastError());
    }
    return 0;
}

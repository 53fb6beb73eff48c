// K1 — dense CLIP/SigLIP similarity as a tcgen05 GEMM whose epilogue ranks instead of storing.
//
//   S[q, g] = sum_k bf16(Q[q,k]) * bf16(G[g,k])   (fp32 accumulate in TMEM)
//
// Replaces torch.matmul(txt.float(), img.float().t()) at /root/reference/main.py:473 (v28 :390-391) together with
// the per-row torch.topk membership tests of main.py:331-378: the 5k x 25k score matrix is never written.
// Two passes over the same MMA code, so a ground-truth score is bit-identical in both and can never beat itself:
//   MODE_THRESHOLD  only tiles that hold a GT column; epilogue: thr[q] = max(thr[q], S[q, c]) for c in GT(q)
//   MODE_COUNT      every tile; epilogue: bits = (S[q, :] > thr[q]) packed 32 columns per word, cnt_above += popc,
//                   optionally the fp32 scores themselves (the drop-in compute_similarity_scores needs them).
// Top-k without the matrix (mmx_dense_topk) is two more uses of the same code: MODE_GROUPMAX stores the maximum of
// every 32-column group (256 for huge galleries) per query; the k-th largest group maximum t_k of a row is a lower
// bound of its k-th largest score (k groups hold a value >= t_k), so a COUNT pass with the bar at t_k meets only a
// handful of candidates per row (k plus the few extra values sharing a group) and appends them (score, column) to a
// short per-row list, from which a warp picks the final k.
//
// CTA = persistent worker over 256 x 256 output tiles = two 128-row halves that share every B block: a 64-wide K block
// is 2 x 16 KB of A + 32 KB of B for eight 128x256x16 MMAs (1024 clk), i.e. 64 B/clk/SM of L2->SM operand traffic
// (a 128 x 256 tile needs 96 B/clk/SM and was limited by exactly that, 63-68 % of the bf16 peak). K streams through a
// 3-stage TMA ring of 128-byte-swizzled blocks; the halves accumulate in TMEM columns [0,256) and [256,512), each
// with its own full/empty barrier: half 0 finishes 512 clk before half 1, its four epilogue warps drain it while the
// tensor core is still busy with half 1, and the next tile's half-0 MMAs start as soon as they are done.
// Warp roles: 0 = TMA, 1 = MMA issuer, 2 = TMEM allocator, 4..7 = epilogue of half 0, 8..11 = epilogue of half 1.
#include <algorithm>
#include <cfloat>
#include <cuda_bf16.h>

#include <cstdlib>

#include "../../include/mmx.h"
#include "mmx_host.cuh"
#include "sm100_ptx.cuh"
#include "topk.cuh"

#define MMX_MODE_GROUPMAX 2   // internal to mmx_dense_topk
#define MMX_MODE_LISTS 3      // internal to mmx_dense_rank_lists: count + above-GT column lists, no bitmask
// Optional outputs of the COUNT pass, compile-time (FEAT template parameter): with all of them behind run-time flags
// the unrolled epilogue was ~200 KB of SASS and the pass lost 7 % (5000 x 25 000) to 27 % (10^6 x 10^6) of its
// throughput to instruction-cache misses; every caller wants at most one of them.
#define MMX_FEAT_NONE 0    // counts (+ bitmask): the ranking pass of every dense model
#define MMX_FEAT_GRP 1     // + group maxima for the top-k bound (mmx_dense_rank_topk)
#define MMX_FEAT_CAND 2    // + (score, column) candidates above the bar (second pass of the matrix-free top-k)
#define MMX_FEAT_SCORES 3  // + the fp32 scores themselves (compute_similarity_scores drop-in, small-gallery top-k)

namespace mmx {
namespace dense {

constexpr int BM = 128, BN = 256, BK = 64;   // one MMA: 128 rows (TMEM lanes) x 256 columns
constexpr int TM = 2 * BM;                   // rows of an output tile (two accumulator halves)
constexpr int NSTAGE = 3;
constexpr int A_BYTES = BM * 128, B_BYTES = BN * 128, STAGE_BYTES = 2 * A_BYTES + B_BYTES;   // 64 KB
constexpr int NUM_THREADS = 384;
constexpr int TMEM_COLS = 512;
constexpr long long BAND_BYTES = 32ll << 20;  // rows of Q kept hot in L2 while a band sweeps every column tile

struct Params {
    const int32_t* gt_off;
    const int32_t* gt_idx;
    long long g_base;
    float* thr;
    int32_t* cnt_above;
    uint32_t* above_bits;
    long long words_per_row;
    float* scores;
    long long ld_scores;
    const int32_t* tile_list;   // MODE_THRESHOLD: compacted ids of tiles holding a GT column
    const int32_t* n_tiles_dev; // MODE_THRESHOLD: their count (device scalar)
    float* grp_max;             // MODE_GROUPMAX: [nq, n_grp] maxima of 32-column groups (grp_tile: of 256-column tiles)
    long long n_grp;
    int grp_tile;
    float* cand_s;              // MODE_COUNT, optional: scores above thr are appended to a per-row candidate list
    int32_t* cand_i;            //   [nq, cand_cap] global column ids; cand_n [nq] counts every hit (may exceed the cap)
    int32_t* cand_n;
    int cand_cap;
    long long nq, ng;
    int kblocks, n_mblk, n_ntile, n_tiles, mode;
    int band;                   // 256-row tiles per raster band: concurrent CTAs share A and B tiles through L2
    int n_mpair, band_pair, n_tiles_pair;  // CTA-pair kernels: the same raster over 512-row tile pairs
};

// CTA-pair kernels (tcgen05 cta_group::2): a stage holds this CTA's two A halves and 128 of the 256 gallery rows
constexpr int PSTAGE = 4;
constexpr int BP_BYTES = (BN / 2) * 128;                  // 16 KB
constexpr int PSTAGE_BYTES = 2 * A_BYTES + BP_BYTES;      // 48 KB
static_assert(PSTAGE * PSTAGE_BYTES <= NSTAGE * STAGE_BYTES, "the pair kernels reuse the ring's shared memory");

struct __align__(1024) Smem {
    uint8_t stage[NSTAGE][STAGE_BYTES];
    uint64_t pfull[PSTAGE], pempty[PSTAGE];   // pair kernels
    uint64_t full[NSTAGE], empty[NSTAGE];
    uint64_t t_full[2], t_empty[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ void tile_coords(const Params& p, int t, int& m_blk, int& n_tile) {
    const int per_band = p.band * p.n_ntile;
    const int band = t / per_band;
    const int rem = t - band * per_band;
    const int rows = min(p.band, p.n_mblk - band * p.band);
    n_tile = rem / rows;
    m_blk = band * p.band + rem % rows;
}
// the same raster over 512-row tile PAIRS (pair kernels): m_pair = index of the pair of 256-row tiles
__device__ __forceinline__ void tile_coords_pair(const Params& p, int t, int& m_pair, int& n_tile) {
    const int per_band = p.band_pair * p.n_ntile;
    const int band = t / per_band;
    const int rem = t - band * per_band;
    const int rows = min(p.band_pair, p.n_mpair - band * p.band_pair);
    n_tile = rem / rows;
    m_pair = band * p.band_pair + rem % rows;
}
__device__ __forceinline__ int tile_id(const Params& p, int i) {
    return p.tile_list != nullptr ? __ldg(p.tile_list + i) : i;
}

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    // by sign BIT: -0.0 compares >= 0 but its pattern is INT_MIN, which a signed max would lose against -inf
    if (__float_as_int(v) >= 0)
        atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else
        atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// MODE_LISTS: write the columns of the recorded hit words into the list slots [base, base + hits) that were reserved
// for them (slots beyond the capacity are dropped) and clear the words.
__device__ __forceinline__ void flush_pending(const Params& p, uint32_t (&bits)[BN / 32], int base, long long q,
                                              long long n0) {
    int slot = base;
#pragma unroll
    for (int c = 0; c < BN / 32; ++c) {
        uint32_t w = bits[c];
        bits[c] = 0u;
        while (w != 0u && slot < p.cand_cap) {
            const int j = __ffs(w) - 1;
            w &= w - 1u;
            p.cand_i[q * p.cand_cap + slot] = (int32_t)(p.g_base + n0 + c * 32 + j);
            ++slot;
        }
        slot += __popc(w);  // hits that found the list full still own their slots
    }
}

// MODE is a template parameter: each pass gets an epilogue without the other passes' code (the threshold pass
// indexes the accumulator registers dynamically, which would put the COUNT pass's registers on the stack too).
// PAIR = true: launched as clusters of two CTAs (one TPC). The pair owns a 512 x 256 output tile — 256 rows per CTA —
// and every 256-row gallery block is split: each CTA's TMA fetches 128 of its rows (completing on the leader's barrier)
// and one tcgen05.mma.cta_group::2 (M = 256: 128 rows of each CTA, N = 256) issued by the leader fills the same
// accumulator half in both CTAs' TMEM, so every CTA pulls 48 KB instead of 64 KB per K block from L2. Epilogues are per
// CTA and unchanged; their releases arrive on the leader's t_empty (the peer's through the cluster), the MMA commits are
// multicast to both CTAs. Not used for the threshold pass (a compacted list of single tiles).
template <int MODE, int FEAT = MMX_FEAT_NONE, bool PAIR = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
dense_rank_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_g,
                  const __grid_constant__ Params p) {
    extern __shared__ uint8_t smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    // work distribution: CTAs (or CTA pairs) stride over the tile list
    const int w_first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int w_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_g);
    }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < NSTAGE; ++i) {
            mbar_init(&s.full[i], 1);
            mbar_init(&s.empty[i], 1);
        }
        for (int i = 0; i < PSTAGE; ++i) {
            mbar_init(&s.pfull[i], 1);
            mbar_init(&s.pempty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s.t_full[i], 1);
            mbar_init(&s.t_empty[i], PAIR ? 8 : 4);   // pair: the four epilogue warps of this half in BOTH CTAs
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        if (PAIR) {
            tmem_alloc_2cta(&s.tmem_base, TMEM_COLS);
            tmem_relinquish_2cta();
        } else {
            tmem_alloc(&s.tmem_base, TMEM_COLS);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all();
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s.tmem_base;
    const int n_work = PAIR ? p.n_tiles_pair : (p.n_tiles_dev != nullptr ? __ldg(p.n_tiles_dev) : p.n_tiles);
    uint8_t* const pring = &s.stage[0][0];   // pair kernels: [PSTAGE][A0 | A1 | B half]

    if (warp == 0 && PAIR) {
        if (elect_one()) {
            uint32_t kc = 0;
            for (int i = w_first; i < n_work; i += w_step) {
                int m_pair, n_tile;
                tile_coords_pair(p, i, m_pair, n_tile);
                const int m_blk = 2 * m_pair + (int)rank;
                for (int kb = 0; kb < p.kblocks; ++kb, ++kc) {
                    const uint32_t st = kc % PSTAGE, ph = (kc / PSTAGE) & 1;
                    mbar_wait(&s.pempty[st], ph ^ 1);
                    if (leader) mbar_arrive_expect_tx(&s.pfull[st], 2 * PSTAGE_BYTES);
                    const uint32_t full_leader = mapa_shared(smem_u32(&s.pfull[st]), 0);
                    uint8_t* dst = pring + st * PSTAGE_BYTES;
                    tma_load_2d_2cta(dst, &tmap_q, full_leader, kb * BK, m_blk * TM);
                    tma_load_2d_2cta(dst + A_BYTES, &tmap_q, full_leader, kb * BK, m_blk * TM + BM);
                    tma_load_2d_2cta(dst + 2 * A_BYTES, &tmap_g, full_leader, kb * BK, n_tile * BN + (int)rank * (BN / 2));
                }
            }
        }
    } else if (warp == 1 && PAIR) {
        if (leader && elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN);
            uint32_t kc = 0, tc = 0;
            for (int i = w_first; i < n_work; i += w_step, ++tc) {
                for (int kb = 0; kb < p.kblocks; ++kb, ++kc) {
                    const uint32_t st = kc % PSTAGE, ph = (kc / PSTAGE) & 1;
                    mbar_wait(&s.pfull[st], ph);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(pring + st * PSTAGE_BYTES);
                    const uint32_t b_addr = a_addr + 2 * A_BYTES;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (kb == 0) {  // the epilogue warps of this half (both CTAs) have drained the previous tile
                            mbar_wait(&s.t_empty[h], (tc & 1) ^ 1);
                            tc_fence_after();
                        }
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            umma_bf16_ss_2cta(tmem_base + h * BN, umma_desc_kmajor_sw128(a_addr + h * A_BYTES + k * 32),
                                              umma_desc_kmajor_sw128(b_addr + k * 32), idesc, (kb | k) != 0);
                        if (kb == p.kblocks - 1) umma_commit_2cta(&s.t_full[h]);
                    }
                    umma_commit_2cta(&s.pempty[st]);
                }
            }
        }
    } else if (warp == 0) {
        if (elect_one()) {
            uint32_t kc = 0;
            for (int i = blockIdx.x; i < n_work; i += gridDim.x) {
                int m_blk, n_tile;
                tile_coords(p, tile_id(p, i), m_blk, n_tile);
                for (int kb = 0; kb < p.kblocks; ++kb, ++kc) {
                    const uint32_t st = kc % NSTAGE, ph = (kc / NSTAGE) & 1;
                    mbar_wait(&s.empty[st], ph ^ 1);
                    mbar_arrive_expect_tx(&s.full[st], STAGE_BYTES);
                    tma_load_2d(s.stage[st], &tmap_q, &s.full[st], kb * BK, m_blk * TM);
                    tma_load_2d(s.stage[st] + A_BYTES, &tmap_q, &s.full[st], kb * BK, m_blk * TM + BM);
                    tma_load_2d(s.stage[st] + 2 * A_BYTES, &tmap_g, &s.full[st], kb * BK, n_tile * BN);
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
            uint32_t kc = 0, tc = 0;
            for (int i = blockIdx.x; i < n_work; i += gridDim.x, ++tc) {
                for (int kb = 0; kb < p.kblocks; ++kb, ++kc) {
                    const uint32_t st = kc % NSTAGE, ph = (kc / NSTAGE) & 1;
                    mbar_wait(&s.full[st], ph);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(s.stage[st]);
                    const uint32_t b_addr = a_addr + 2 * A_BYTES;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (kb == 0) {  // the epilogue warps of this half have drained the previous tile
                            mbar_wait(&s.t_empty[h], (tc & 1) ^ 1);
                            tc_fence_after();
                        }
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            umma_bf16_ss(tmem_base + h * BN, umma_desc_kmajor_sw128(a_addr + h * A_BYTES + k * 32),
                                         umma_desc_kmajor_sw128(b_addr + k * 32), idesc, (kb | k) != 0);
                        if (kb == p.kblocks - 1) umma_commit(&s.t_full[h]);
                    }
                    umma_commit(&s.empty[st]);
                }
            }
        }
    } else if (warp >= 4) {
        const int wq = warp & 3;
        const uint32_t acc = (warp - 4) >> 2;  // which 128-row half of every tile this warp drains
        const int row = acc * BM + wq * 32 + lane;
        uint32_t tc = 0;
        // MODE_LISTS: hit words of the last drained tile, their row / first column and the reserved list slots
        uint32_t pend_bits[BN / 32];
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) pend_bits[c] = 0u;
        int pend_base = 0, pend_m = 0, pend_nt = 0;  // reserved slots, row tile and column tile of those hits
        const uint32_t t_empty_leader = PAIR ? mapa_shared(smem_u32(&s.t_empty[acc]), 0) : 0u;
        for (int i = w_first; i < n_work; i += w_step, ++tc) {
            int m_blk, n_tile;
            if (PAIR) {
                tile_coords_pair(p, i, m_blk, n_tile);
                m_blk = 2 * m_blk + (int)rank;
            } else {
                tile_coords(p, tile_id(p, i), m_blk, n_tile);
            }
            const uint32_t use = tc;
            const long long q = (long long)m_blk * TM + row;
            const long long n0 = (long long)n_tile * BN;
            const bool valid_row = q < p.nq;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * BN;
            // per-row state, loaded while the MMAs of this tile are still running
            float thr = INFINITY;
            int e0 = 0, e1 = 0;
            // candidate / above-GT list emission: a row whose counter had already passed the capacity when this tile
            // started appends nothing more (rows far below their threshold would otherwise pay one atomic per hit);
            // the counter still ends above the capacity, which is how the top-k path recognises an overflowed row
            bool emit = false;
            if (MODE == MMX_MODE_COUNT) {
                if (valid_row) thr = __ldg(p.thr + q);
                if (FEAT == MMX_FEAT_CAND) emit = valid_row && __ldcg(p.cand_n + q) <= p.cand_cap;
            } else if (MODE == MMX_MODE_LISTS) {
                if (valid_row) thr = __ldg(p.thr + q);
                emit = valid_row && __ldcg(p.cand_n + q) <= p.cand_cap;
                // the hits of the PREVIOUS tile go to the slots reserved at its end: that atomic has long returned, and
                // the tensor core is still busy with this tile, so the list stores cost the pipeline nothing
                flush_pending(p, pend_bits, pend_base, (long long)pend_m * TM + row, (long long)pend_nt * BN);
            } else if (MODE == MMX_MODE_THRESHOLD && valid_row) {
                e0 = __ldg(p.gt_off + q);
                e1 = __ldg(p.gt_off + q + 1);
            }
            mbar_wait(&s.t_full[acc], use & 1);
            tc_fence_after();
            int cnt = 0;
            float best = -INFINITY;
            uint32_t vbuf[2][32];
            tmem_ld_32x32(taddr, vbuf[0]);
            // eight 32-column chunks as four pairs (the two register buffers alternate inside a pair). The candidate
            // variant keeps the pair loop rolled: its per-column append code is 8x smaller that way (117 -> ~30 KB)
            constexpr int PAIR_UNROLL = (MODE == MMX_MODE_COUNT && FEAT == MMX_FEAT_CAND) ? 1 : BN / 64;
#pragma unroll PAIR_UNROLL
            for (int cp = 0; cp < BN / 64; ++cp)
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                const int ch = cp * 2 + h2;
                uint32_t (&v)[32] = vbuf[h2];
                tmem_ld_wait();
                if (ch + 1 < BN / 32) {
                    tmem_ld_32x32(taddr + (ch + 1) * 32, vbuf[h2 ^ 1]);  // in flight while this chunk is ranked
                } else {
                    // the whole accumulator half is in registers: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (PAIR) mbar_arrive_cluster(t_empty_leader);
                        else mbar_arrive(&s.t_empty[acc]);
                    }
                }
                const long long c0 = n0 + ch * 32;
                if (MODE == MMX_MODE_GROUPMAX) {
                    const long long nvalid = p.ng - c0;
                    float m = -INFINITY;
#pragma unroll
                    for (int j = 0; j < 32; ++j) m = fmaxf(m, j < nvalid ? __uint_as_float(v[j]) : -INFINITY);
                    if (p.grp_tile)
                        best = fmaxf(best, m);
                    else if (valid_row && nvalid > 0)
                        p.grp_max[q * p.n_grp + (c0 >> 5)] = m;
                } else if (MODE == MMX_MODE_LISTS) {
                    const long long nvalid = p.ng - c0;
                    const uint32_t cmask = nvalid >= 32 ? 0xffffffffu : (nvalid <= 0 ? 0u : ((1u << (int)nvalid) - 1u));
                    uint32_t bits = 0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) bits |= (__uint_as_float(v[j]) > thr ? 1u : 0u) << j;
                    bits &= cmask;
                    cnt += __popc(bits);
                    pend_bits[ch] = emit ? bits : 0u;
                } else if (MODE == MMX_MODE_COUNT) {
                    const long long nvalid = p.ng - c0;
                    const uint32_t cmask = nvalid >= 32 ? 0xffffffffu : (nvalid <= 0 ? 0u : ((1u << (int)nvalid) - 1u));
                    uint32_t bits = 0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) bits |= (__uint_as_float(v[j]) > thr ? 1u : 0u) << j;
                    bits &= cmask;
                    cnt += __popc(bits);
                    if (FEAT == MMX_FEAT_GRP) {  // fused group maxima for the top-k bound (mmx_dense_rank_topk)
                        float m = -INFINITY;
#pragma unroll
                        for (int j = 0; j < 32; ++j) m = fmaxf(m, j < nvalid ? __uint_as_float(v[j]) : -INFINITY);
                        if (p.grp_tile)
                            best = fmaxf(best, m);
                        else if (valid_row && nvalid > 0)
                            p.grp_max[q * p.n_grp + (c0 >> 5)] = m;
                    }
                    const long long w = (c0 >> 5);
                    if (p.above_bits != nullptr && valid_row && w < p.words_per_row)
                        p.above_bits[q * p.words_per_row + w] = bits;
                    if (FEAT == MMX_FEAT_CAND && emit && bits != 0u) {   // a handful of hits per row (bar just below the k-th group maximum)
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if ((bits >> j) & 1u) {
                                const int slot = atomicAdd(p.cand_n + q, 1);
                                if (slot < p.cand_cap) {
                                    p.cand_s[q * p.cand_cap + slot] = __uint_as_float(v[j]);
                                    p.cand_i[q * p.cand_cap + slot] = (int32_t)(p.g_base + c0 + j);
                                }
                            }
                        }
                    }
                    if (FEAT == MMX_FEAT_SCORES && valid_row) {
                        float* dst = p.scores + q * p.ld_scores + c0;
                        if (nvalid >= 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                *reinterpret_cast<float4*>(dst + j) =
                                    make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j < nvalid) dst[j] = __uint_as_float(v[j]);
                        }
                    }
                } else {
                    for (int e = e0; e < e1; ++e) {
                        const long long lc = (long long)__ldg(p.gt_idx + e) - p.g_base - c0;
                        if (lc >= 0 && lc < 32 && lc + c0 < p.ng) {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j == (int)lc) best = fmaxf(best, __uint_as_float(v[j]));
                        }
                    }
                }
            }
            if (MODE == MMX_MODE_LISTS) {
                if (valid_row && cnt > 0) atomicAdd(p.cnt_above + q, cnt);
                // reserve list slots for this tile's hits; the result is only used at the start of the next tile
                int tot = 0;
#pragma unroll
                for (int c = 0; c < BN / 32; ++c) tot += __popc(pend_bits[c]);
                if (tot > 0) pend_base = atomicAdd(p.cand_n + q, tot);
                pend_m = m_blk;
                pend_nt = n_tile;
            } else if (MODE == MMX_MODE_COUNT) {
                if (valid_row && cnt > 0) atomicAdd(p.cnt_above + q, cnt);
                if (FEAT == MMX_FEAT_GRP && p.grp_tile && valid_row) p.grp_max[q * p.n_grp + n_tile] = best;
            } else if (MODE == MMX_MODE_GROUPMAX) {
                if (p.grp_tile && valid_row) p.grp_max[q * p.n_grp + n_tile] = best;
            } else if (valid_row && best > -INFINITY) {
                atomic_max_float(p.thr + q, best);
            }
        }
        if (MODE == MMX_MODE_LISTS)
            flush_pending(p, pend_bits, pend_base, (long long)pend_m * TM + row, (long long)pend_nt * BN);
    }

    tc_fence_before();
    if (PAIR) {
        cluster_sync_all();  // the peer's tensor core may still read this CTA's shared memory / write its TMEM
        if (warp == 2) tmem_dealloc_2cta(tmem_base, TMEM_COLS);
    } else {
        __syncthreads();
        if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// MODE_THRESHOLD work list: ids of the (row tile, column tile) pairs that contain at least one GT column.
__global__ void gt_tiles_kernel(const int32_t* __restrict__ gt_off, const int32_t* __restrict__ gt_idx, long long nq,
                                long long ng, long long g_base, int n_mblk, int n_ntile, int band_rows,
                                int32_t* __restrict__ flags, int32_t* __restrict__ list, int32_t* __restrict__ count) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const int m_blk = (int)(q / TM);
    const int band = m_blk / band_rows;
    const int rows = min(band_rows, n_mblk - band * band_rows);
    for (int e = gt_off[q]; e < gt_off[q + 1]; ++e) {
        const long long c = (long long)gt_idx[e] - g_base;
        if (c < 0 || c >= ng) continue;
        const int n_tile = (int)(c / BN);
        const int t = band * band_rows * n_ntile + n_tile * rows + (m_blk - band * band_rows);
        if (atomicExch(&flags[t], 1) == 0) list[atomicAdd(count, 1)] = t;
    }
}

}  // namespace dense
}  // namespace mmx

using namespace mmx;
using namespace mmx::dense;

extern "C" int mmx_dense_workspace_bytes(int64_t nq, int64_t ng, size_t* bytes) {
    MMX_REQUIRE(bytes != nullptr, "bytes == NULL");
    const int64_t tiles = ((nq + TM - 1) / TM) * ((ng + BN - 1) / BN);
    *bytes = (size_t)tiles * 8 + 512;
    return 0;
}

// 0 = one-CTA kernels, 1 (default) = CTA pairs when the sweep has at least one wave of pair tiles, 2 = pairs whenever possible
static int dense_pair_mode() {
    static const int mode = [] {
        const char* e = getenv("MMX_DENSE_PAIR");
        return e != nullptr ? atoi(e) : 1;
    }();
    return mode;
}

template <int MODE, int FEAT>
static int launch_variant(const CUtensorMap& tq, const CUtensorMap& tg, const CUtensorMap& tg_half, const Params& p,
                          cudaStream_t stream) {
    const size_t smem_bytes = sizeof(Smem) + 1024;
    const int sms = sm_count();
    const int pm = dense_pair_mode();
    if (MODE != MMX_MODE_THRESHOLD && pm != 0 && p.n_tiles_pair >= (pm >= 2 ? 1 : sms / 2)) {
        static const char pair_attr_tag = 0;  // one per template instantiation
        if (first_use_on_device(&pair_attr_tag))
            MMX_CHECK_CUDA(cudaFuncSetAttribute(dense_rank_kernel<MODE, FEAT, true>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)std::min(2 * p.n_tiles_pair, sms & ~1));
        cfg.blockDim = dim3(NUM_THREADS);
        cfg.dynamicSmemBytes = smem_bytes;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        MMX_CHECK_CUDA(cudaLaunchKernelEx(&cfg, dense_rank_kernel<MODE, FEAT, true>, tq, tg_half, p));
        return 0;
    }
    static const char attr_set_tag = 0;  // one per template instantiation
    if (first_use_on_device(&attr_set_tag))
        MMX_CHECK_CUDA(cudaFuncSetAttribute(dense_rank_kernel<MODE, FEAT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem_bytes));
    const int grid = p.n_tiles < sms ? p.n_tiles : sms;
    dense_rank_kernel<MODE, FEAT><<<grid, NUM_THREADS, smem_bytes, stream>>>(tq, tg, p);
    MMX_CHECK_CUDA(cudaGetLastError());
    return 0;
}

static int launch_dense(const void* Q, int64_t nq, const void* G, int64_t ng, int d, const Params& p,
                        cudaStream_t stream) {
    CUtensorMap tq, tg, th;
    if (make_tmap_bf16_sw128(&tq, Q, (uint64_t)nq, (uint64_t)d, (uint64_t)d, BM) != 0) return -4;
    if (make_tmap_bf16_sw128(&tg, G, (uint64_t)ng, (uint64_t)d, (uint64_t)d, BN) != 0) return -4;
    if (make_tmap_bf16_sw128(&th, G, (uint64_t)ng, (uint64_t)d, (uint64_t)d, BN / 2) != 0) return -4;   // pair kernels
    if (p.mode == MMX_MODE_THRESHOLD) return launch_variant<MMX_MODE_THRESHOLD, MMX_FEAT_NONE>(tq, tg, th, p, stream);
    if (p.mode == MMX_MODE_GROUPMAX) return launch_variant<MMX_MODE_GROUPMAX, MMX_FEAT_NONE>(tq, tg, th, p, stream);
    if (p.mode == MMX_MODE_LISTS) return launch_variant<MMX_MODE_LISTS, MMX_FEAT_NONE>(tq, tg, th, p, stream);
    // MMX_MODE_COUNT: at most one optional output per pass
    const int n_opt = (p.grp_max != nullptr) + (p.cand_n != nullptr) + (p.scores != nullptr);
    MMX_REQUIRE(n_opt <= 1, "COUNT pass: group maxima, candidates and scores are mutually exclusive");
    if (p.grp_max != nullptr) return launch_variant<MMX_MODE_COUNT, MMX_FEAT_GRP>(tq, tg, th, p, stream);
    if (p.cand_n != nullptr) {
        MMX_REQUIRE(p.cand_s != nullptr && p.cand_i != nullptr, "candidate buffers missing");
        return launch_variant<MMX_MODE_COUNT, MMX_FEAT_CAND>(tq, tg, th, p, stream);
    }
    if (p.scores != nullptr) return launch_variant<MMX_MODE_COUNT, MMX_FEAT_SCORES>(tq, tg, th, p, stream);
    return launch_variant<MMX_MODE_COUNT, MMX_FEAT_NONE>(tq, tg, th, p, stream);
}

static int base_params(Params& p, int64_t nq, int64_t ng, int d) {
    const long long n_mblk = (nq + TM - 1) / TM, n_ntile = (ng + BN - 1) / BN;
    MMX_REQUIRE(n_mblk * n_ntile < (1ll << 31), "too many tiles");
    p = Params{};
    p.nq = nq;
    p.ng = ng;
    p.kblocks = (d + BK - 1) / BK;
    p.n_mblk = (int)n_mblk;
    p.n_ntile = (int)n_ntile;
    p.n_tiles = (int)(n_mblk * n_ntile);
    long long band = BAND_BYTES / ((long long)TM * d * 2);
    p.band = (int)(band < 8 ? 8 : (band > 64 ? 64 : band));
    p.n_mpair = (int)((n_mblk + 1) / 2);
    p.band_pair = std::max(4, p.band / 2);
    p.n_tiles_pair = (int)((long long)p.n_mpair * n_ntile);
    return 0;
}

extern "C" int mmx_dense_rank(const void* Q, int64_t nq, const void* G, int64_t ng, int d, const int32_t* gt_off,
                              const int32_t* gt_idx, int64_t g_base, int mode, float* thr, int32_t* cnt_above,
                              uint32_t* above_bits, int64_t words_per_row, float* scores, int64_t ld_scores,
                              void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    MMX_REQUIRE(nq >= 0 && ng >= 0 && d >= 8 && d % 8 == 0, "bad shape nq=%lld ng=%lld d=%d (d must be a multiple of 8)",
                (long long)nq, (long long)ng, d);
    MMX_REQUIRE(mode == MMX_MODE_THRESHOLD || mode == MMX_MODE_COUNT, "unknown mode %d", mode);
    MMX_REQUIRE(thr != nullptr, "thr == NULL");
    if (nq == 0 || ng == 0) return 0;
    if (mmx_device_check() != 0) return -3;
    Params p;
    if (base_params(p, nq, ng, d) != 0) return -2;
    p.gt_off = gt_off;
    p.gt_idx = gt_idx;
    p.g_base = g_base;
    p.thr = thr;
    p.cnt_above = cnt_above;
    p.above_bits = above_bits;
    p.words_per_row = words_per_row;
    p.scores = scores;
    p.ld_scores = ld_scores;
    p.mode = mode;
    if (mode == MMX_MODE_COUNT) {
        MMX_REQUIRE(cnt_above != nullptr, "cnt_above == NULL");
        MMX_REQUIRE(above_bits == nullptr || words_per_row >= (ng + 31) / 32, "words_per_row too small");
        MMX_REQUIRE(scores == nullptr || ld_scores >= ng, "ld_scores < ng");
    } else {
        MMX_REQUIRE(gt_off != nullptr && gt_idx != nullptr, "GT CSR missing");
        size_t need = 0;
        mmx_dense_workspace_bytes(nq, ng, &need);
        MMX_REQUIRE(workspace != nullptr && workspace_bytes >= need, "workspace too small: need %zu bytes", need);
        uintptr_t ws = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
        int32_t* count = reinterpret_cast<int32_t*>(ws);
        int32_t* flags = count + 64;
        int32_t* list = flags + p.n_tiles;
        MMX_CHECK_CUDA(cudaMemsetAsync(count, 0, (size_t)(64 + p.n_tiles) * 4, stream));
        gt_tiles_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, stream>>>(gt_off, gt_idx, nq, ng, g_base, p.n_mblk,
                                                                          p.n_ntile, p.band, flags, list, count);
        MMX_CHECK_CUDA(cudaGetLastError());
        p.tile_list = list;
        p.n_tiles_dev = count;
    }
    return launch_dense(Q, nq, G, ng, d, p, stream);
}

// ---------------------------------------------------------------------------------------------------------------
// mmx_dense_topk — top-k (score, column) per query without the score matrix
// ---------------------------------------------------------------------------------------------------------------
namespace mmx {
namespace dense {

constexpr int CAND_CAP = 64;           // candidate slots per row; rows that overflow (mass ties) are recomputed exactly
constexpr int64_t SMALL_GALLERY = 4096;  // below this the matrix is a few KB per row: materialise it in the workspace

// bar[q] = the largest float below the k-th largest group maximum, so that "> bar" means ">= t_k"
__global__ void __launch_bounds__(256)
kth_group_bar_kernel(const float* __restrict__ grp_max, long long nq, long long n_grp, int k, float* __restrict__ bar) {
    const int lane = threadIdx.x & 31;
    const long long q = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (q >= nq) return;
    topk::WarpList L;
    L.init();
    const float* row = grp_max + q * n_grp;
    for (long long c = 0; c < n_grp; c += 32) {
        const bool ok = c + lane < n_grp;
        topk::offer(L, ok ? __ldg(row + c + lane) : -INFINITY, (int)(c + lane), ok, k, lane);
    }
    if (lane == k - 1) {
        float t = L.v;  // -inf when the row has fewer than k groups
        if (t > -INFINITY) t = (t == 0.0f) ? -1.401298464e-45f : __uint_as_float(__float_as_uint(t) + (t > 0.0f ? -1 : 1));
        bar[q] = t;
    }
}

// exact recomputation of the rows whose candidate list overflowed: one warp per row, fp32 dot products
__global__ void __launch_bounds__(256)
overflow_rows_kernel(const __nv_bfloat16* __restrict__ Q, const __nv_bfloat16* __restrict__ G, long long nq,
                     long long ng, int d, long long g_base, const int32_t* __restrict__ cand_n, int k,
                     float* __restrict__ top_s, int32_t* __restrict__ top_i) {
    const int lane = threadIdx.x & 31;
    const long long q = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (q >= nq || __ldg(cand_n + q) <= CAND_CAP) return;
    topk::WarpList L;
    L.init();
    const __nv_bfloat16* qr = Q + q * d;
    for (long long c = 0; c < ng; c += 32) {
        const bool ok = c + lane < ng;
        float acc = 0.0f;
        if (ok) {
            const __nv_bfloat16* gr = G + (c + lane) * d;
            for (int i = 0; i < d; ++i) acc = fmaf(__bfloat162float(qr[i]), __bfloat162float(gr[i]), acc);
        }
        topk::offer(L, acc, (int)(c + lane), ok, k, lane);
    }
    if (lane < k) {
        const bool real = L.i != 0x7fffffff;
        top_s[q * k + lane] = real ? L.v : -INFINITY;
        top_i[q * k + lane] = real ? (int32_t)(g_base + L.i) : -1;
    }
}

struct TopkPlan;
}  // namespace dense
}  // namespace mmx
static int topk_from_group_maxima(const void* Q, int64_t nq, const void* G, int64_t ng, int d, int64_t g_base, int k,
                                  float* top_s, int32_t* top_i, uint8_t* ws, const mmx::dense::TopkPlan& t,
                                  cudaStream_t stream);
namespace mmx {
namespace dense {

struct TopkPlan {
    bool small;
    int grp_tile;
    long long n_grp;
    size_t off_bar, off_cnt, off_candn, off_cs, off_ci, off_grp, total;
};

static TopkPlan plan_topk(int64_t nq, int64_t ng) {
    TopkPlan t{};
    auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
    t.small = ng <= SMALL_GALLERY;
    size_t o = 256;
    t.off_bar = o;   o += up((size_t)nq * 4);
    t.off_cnt = o;   o += up((size_t)nq * 4);
    if (t.small) {
        t.off_grp = o; o += up((size_t)nq * (size_t)ng * 4);          // the score matrix itself
    } else {
        t.grp_tile = ng > 32768 ? 1 : 0;
        t.n_grp = t.grp_tile ? (ng + BN - 1) / BN : (ng + 31) / 32;
        t.off_candn = o; o += up((size_t)nq * 4);
        t.off_cs = o;    o += up((size_t)nq * CAND_CAP * 4);
        t.off_ci = o;    o += up((size_t)nq * CAND_CAP * 4);
        t.off_grp = o;   o += up((size_t)nq * (size_t)t.n_grp * 4);
    }
    t.total = o;
    return t;
}

}  // namespace dense
}  // namespace mmx

extern "C" int mmx_dense_topk_workspace_bytes(int64_t nq, int64_t ng, size_t* bytes) {
    MMX_REQUIRE(bytes != nullptr && nq >= 0 && ng >= 0, "bad argument");
    *bytes = plan_topk(nq, ng).total;
    return 0;
}

extern "C" int mmx_dense_topk(const void* Q, int64_t nq, const void* G, int64_t ng, int d, int64_t g_base, int k,
                              float* top_s, int32_t* top_i, void* workspace, size_t workspace_bytes,
                              cudaStream_t stream) {
    MMX_REQUIRE(nq >= 0 && ng >= 0 && d >= 8 && d % 8 == 0, "bad shape nq=%lld ng=%lld d=%d (d must be a multiple of 8)",
                (long long)nq, (long long)ng, d);
    MMX_REQUIRE(k >= 1 && k <= topk::MAX_K, "k must be in [1, %d]", topk::MAX_K);
    MMX_REQUIRE(top_s != nullptr && top_i != nullptr, "NULL output");
    MMX_REQUIRE(g_base + ng < (1ll << 31), "column index does not fit int32");
    if (nq == 0) return 0;
    if (mmx_device_check() != 0) return -3;
    const TopkPlan t = plan_topk(nq, ng);
    MMX_REQUIRE(workspace != nullptr && workspace_bytes >= t.total, "workspace too small: need %zu bytes", t.total);
    uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
    float* bar = reinterpret_cast<float*>(ws + t.off_bar);
    int32_t* cnt = reinterpret_cast<int32_t*>(ws + t.off_cnt);
    Params p;
    if (ng == 0) return mmx_topk_from_scores(nullptr, nq, 0, 0, g_base, k, top_s, top_i, stream);
    if (base_params(p, nq, ng, d) != 0) return -2;
    p.g_base = g_base;
    p.thr = bar;
    p.cnt_above = cnt;
    if (t.small) {
        // bar = +inf: nothing counts, the epilogue only stores the scores
        float* S = reinterpret_cast<float*>(ws + t.off_grp);
        MMX_CHECK_CUDA(cudaMemsetAsync(bar, 0x7f, (size_t)nq * 4, stream));   // 0x7f7f7f7f = 3.39e38 > any score
        p.mode = MMX_MODE_COUNT;
        p.scores = S;
        p.ld_scores = ng;
        if (int rc = launch_dense(Q, nq, G, ng, d, p, stream)) return rc;
        return mmx_topk_from_scores(S, nq, ng, ng, g_base, k, top_s, top_i, stream);
    }
    float* grp = reinterpret_cast<float*>(ws + t.off_grp);
    p.mode = MMX_MODE_GROUPMAX;
    p.grp_max = grp;
    p.n_grp = t.n_grp;
    p.grp_tile = t.grp_tile;
    if (int rc = launch_dense(Q, nq, G, ng, d, p, stream)) return rc;
    return topk_from_group_maxima(Q, nq, G, ng, d, g_base, k, top_s, top_i, ws, t, stream);
}

// second half of the matrix-free top-k: bar from the group maxima, candidate pass, selection, exact overflow rows
static int topk_from_group_maxima(const void* Q, int64_t nq, const void* G, int64_t ng, int d, int64_t g_base, int k,
                                  float* top_s, int32_t* top_i, uint8_t* ws, const TopkPlan& t, cudaStream_t stream) {
    float* bar = reinterpret_cast<float*>(ws + t.off_bar);
    int32_t* cnt = reinterpret_cast<int32_t*>(ws + t.off_cnt);
    float* grp = reinterpret_cast<float*>(ws + t.off_grp);
    int32_t* cand_n = reinterpret_cast<int32_t*>(ws + t.off_candn);
    float* cand_s = reinterpret_cast<float*>(ws + t.off_cs);
    int32_t* cand_i = reinterpret_cast<int32_t*>(ws + t.off_ci);
    Params p;
    if (base_params(p, nq, ng, d) != 0) return -2;
    p.g_base = g_base;
    p.thr = bar;
    p.cnt_above = cnt;
    kth_group_bar_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, stream>>>(grp, nq, t.n_grp, k, bar);
    MMX_CHECK_CUDA(cudaGetLastError());
    MMX_CHECK_CUDA(cudaMemsetAsync(cand_n, 0, (size_t)nq * 4, stream));
    MMX_CHECK_CUDA(cudaMemsetAsync(cnt, 0, (size_t)nq * 4, stream));
    p.mode = MMX_MODE_COUNT;
    p.grp_max = nullptr;
    p.cand_s = cand_s;
    p.cand_i = cand_i;
    p.cand_n = cand_n;
    p.cand_cap = CAND_CAP;
    if (int rc = launch_dense(Q, nq, G, ng, d, p, stream)) return rc;
    if (int rc = mmx_topk_select(cand_s, cand_i, cand_n, nq, CAND_CAP, k, top_s, top_i, stream)) return rc;
    overflow_rows_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(Q), reinterpret_cast<const __nv_bfloat16*>(G), nq, ng, d, g_base, cand_n,
        k, top_s, top_i);
    MMX_CHECK_CUDA(cudaGetLastError());
    return 0;
}


// K1 ranking COUNT pass and the top-k in one call: the COUNT pass (which every tile goes through anyway) also
// stores the group maxima, so the lists cost ONE extra GEMM pass (the candidate pass) instead of two.
// thr must hold the (shard-merged) thresholds, cnt_above must be zeroed by the caller, exactly as for mmx_dense_rank
// in MMX_MODE_COUNT. workspace: mmx_dense_topk_workspace_bytes(nq, ng).
extern "C" int mmx_dense_rank_topk(const void* Q, int64_t nq, const void* G, int64_t ng, int d, const int32_t* gt_off,
                                   const int32_t* gt_idx, int64_t g_base, float* thr, int32_t* cnt_above,
                                   uint32_t* above_bits, int64_t words_per_row, int k, float* top_s, int32_t* top_i,
                                   void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    MMX_REQUIRE(nq >= 0 && ng >= 0 && d >= 8 && d % 8 == 0, "bad shape nq=%lld ng=%lld d=%d (d must be a multiple of 8)",
                (long long)nq, (long long)ng, d);
    MMX_REQUIRE(k >= 1 && k <= topk::MAX_K, "k must be in [1, %d]", topk::MAX_K);
    MMX_REQUIRE(thr != nullptr && cnt_above != nullptr && top_s != nullptr && top_i != nullptr, "NULL argument");
    MMX_REQUIRE(above_bits == nullptr || words_per_row >= (ng + 31) / 32, "words_per_row too small");
    MMX_REQUIRE(g_base + ng < (1ll << 31), "column index does not fit int32");
    if (nq == 0) return 0;
    if (mmx_device_check() != 0) return -3;
    if (ng == 0) return mmx_topk_from_scores(nullptr, nq, 0, 0, g_base, k, top_s, top_i, stream);
    const TopkPlan t = plan_topk(nq, ng);
    MMX_REQUIRE(workspace != nullptr && workspace_bytes >= t.total, "workspace too small: need %zu bytes", t.total);
    uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
    Params p;
    if (base_params(p, nq, ng, d) != 0) return -2;
    p.gt_off = gt_off;
    p.gt_idx = gt_idx;
    p.g_base = g_base;
    p.thr = thr;
    p.cnt_above = cnt_above;
    p.above_bits = above_bits;
    p.words_per_row = words_per_row;
    p.mode = MMX_MODE_COUNT;
    if (t.small) {
        // few columns: the COUNT pass also writes the scores into the workspace and the lists are read off them
        float* S = reinterpret_cast<float*>(ws + t.off_grp);
        p.scores = S;
        p.ld_scores = ng;
        if (int rc = launch_dense(Q, nq, G, ng, d, p, stream)) return rc;
        return mmx_topk_from_scores(S, nq, ng, ng, g_base, k, top_s, top_i, stream);
    }
    p.grp_max = reinterpret_cast<float*>(ws + t.off_grp);
    p.n_grp = t.n_grp;
    p.grp_tile = t.grp_tile;
    if (int rc = launch_dense(Q, nq, G, ng, d, p, stream)) return rc;
    return topk_from_group_maxima(Q, nq, G, ng, d, g_base, k, top_s, top_i, ws, t, stream);
}


// K1 in LIST mode, for galleries whose [nq, ng] above-GT bitmask does not fit in memory (a 10^6 x 10^6
// problem would need 125 GB of bits): the COUNT pass appends the columns that score above the threshold to a per-query
// list of `cap` entries (global column ids; whichever hits arrive first, a full list takes no more) next to the exact
// count. The bootstrap only needs min(c_q, 10) with c_q = sum of the drawn multiplicities over the above-GT set, so
// a list that holds the WHOLE set (cnt_above <= cap) is exact, and a truncated one is exact whenever its own
// multiplicities already reach 10 (mmx_boot_counts_lists reports the other cases for an exact recount).
// thr = merged thresholds (input); cnt_above [nq] and list_n [nq] are zeroed here.
extern "C" int mmx_dense_rank_lists(const void* Q, int64_t nq, const void* G, int64_t ng, int d, int64_t g_base,
                                    const float* thr, int32_t* cnt_above, int32_t* list_i, int32_t* list_n, int cap,
                                    cudaStream_t stream) {
    MMX_REQUIRE(nq >= 0 && ng >= 0 && d >= 8 && d % 8 == 0, "bad shape nq=%lld ng=%lld d=%d (d must be a multiple of 8)",
                (long long)nq, (long long)ng, d);
    MMX_REQUIRE(thr != nullptr && cnt_above != nullptr && list_i != nullptr && list_n != nullptr, "NULL argument");
    MMX_REQUIRE(cap >= 1 && cap <= 1024, "cap out of range");
    MMX_REQUIRE(g_base + ng < (1ll << 31), "column index does not fit int32");
    if (nq == 0) return 0;
    if (mmx_device_check() != 0) return -3;
    MMX_CHECK_CUDA(cudaMemsetAsync(cnt_above, 0, (size_t)nq * 4, stream));
    MMX_CHECK_CUDA(cudaMemsetAsync(list_n, 0, (size_t)nq * 4, stream));
    if (ng == 0) return 0;
    Params p;
    if (base_params(p, nq, ng, d) != 0) return -2;
    p.g_base = g_base;
    p.thr = const_cast<float*>(thr);
    p.cnt_above = cnt_above;
    p.cand_i = list_i;
    p.cand_n = list_n;
    p.cand_cap = cap;
    p.mode = MMX_MODE_LISTS;
    return launch_dense(Q, nq, G, ng, d, p, stream);
}

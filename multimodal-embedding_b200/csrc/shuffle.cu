// K6s — numpy's cumulative in-place shuffles (the permutation test) replayed on the device.
//
// /root/reference/main.py:295-299 calls np.random.shuffle(combined) 10 000 times on the same array. The legacy
// RandomState.shuffle of a 1-D array is Fisher-Yates from the top (numpy/random/mtrand.pyx _shuffle_raw):
//     for i = n-1 .. 1:  j = random_interval(i);  swap(x[i], x[j])
//     random_interval(max): mask = smallest 2^k - 1 >= max;  do v = next_uint32() & mask; while (v > max)
// so every step consumes a data-dependent number of 32-bit MT19937 outputs and 2e7 steps form one serial chain
// (0.27-0.35 s on a host core — ten times the rest of a dense model's statistics on the GPU).
//
// The chain is a finite-state transducer: the state is the current `i` in [1, n-1] (n-1 states), the input one
// random word; a word is accepted iff (w & mask(i)) <= i, and an accept emits j and moves i -> i-1 (wrapping to n-1:
// next permutation). That parallelises exactly, with the raw word stream W produced by K6p:
//   B  tables   for every chunk of L words and EVERY start state, count the accepts. Runs that start in different
//               states never cross, and two runs that meet stay together, so the n-1 runs of a chunk coalesce quickly
//               (n = 2000: 840 distinct runs after 1024 words, 128 after 8192, 40 after 32768): one CTA per chunk keeps
//               the distinct runs as a sorted array (first start position, state, accepts) in shared memory, advances
//               them 512 words at a time and compacts equal neighbours in place — ~12x fewer steps than n-1 full runs;
//   C  chain    one thread walks the chunks: start state and global step offset of each (|W| / L table lookups);
//   D  emit     one thread per chunk replays it from its now known start state and writes the accepted j values in
//               step order; the thread that emits the last step reports how many words were consumed;
//   E  apply    one thread per permutation applies its n-1 swaps to an identity array (column layout [i][P]);
//   F  compose  idx_p = idx_{p-1} o pi_p in three short phases (blocks of permutations composed independently,
//               a scan over the block composites, then each block replayed from its prefix): [P, n] index rows,
//               which is what K5 (mmx_perm_stats) consumes.
// Bit-exact against numpy by construction; tests/test_gpu_parity.py compares the index rows and the generator state.
#include <algorithm>
#include <cmath>

#include "../../include/mmx.h"
#include "mmx_host.cuh"

namespace mmx {
namespace shuf {

constexpr int L = 32768;       // words per chunk (accept counts fit uint16)
constexpr int CT = 512;        // threads of a table CTA
constexpr int RW = 512;        // words per round (simulate, then compact)
constexpr int PB = 64;         // permutations per composition block
constexpr int MAX_N = 16384;   // B/C/D: work grows with n * |W| (~n^2 P); longer arrays take the LONG emit below
constexpr int MAX_N_LONG = 65536;   // LONG emit: masked values fit uint16
constexpr int LT = 1024;            // threads of the LONG emit CTA

constexpr int SERIAL_BELOW = 512;       // LONG emit: steps with i < 512 are replayed by one warp, 32 words at a time
constexpr int MAX_EPOCHS = 20;
// mask epochs of the LONG emit's bracket scheme: bit lengths of n - 1 down to that of SERIAL_BELOW
static int long_epochs(int n) {
    int e = 0;
    for (int i_hi = n - 1; i_hi >= SERIAL_BELOW && e < MAX_EPOCHS;) {
        int b = 0;
        while ((1ll << b) <= i_hi) ++b;
        i_hi = std::max(1 << (b - 1), SERIAL_BELOW) - 1;
        ++e;
    }
    return e;
}

struct Plan {
    long long n_chunks, steps;
    int n_blocks, pb;
    bool long_n, lists_in_smem, rows_in_smem;
    int apply_grid;
    int n_ep;
    size_t off_tab, off_m, off_s, off_j, off_arr, off_pi, off_c, off_d, off_g, off_l, off_rm, off_rb, off_eh, total;
};

constexpr size_t SMEM_MAX = 227 * 1024;

static Plan make_plan(int n, int P, long long n_words) {
    Plan pl{};
    auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
    pl.n_chunks = (n_words + L - 1) / L;
    pl.steps = (long long)P * (n - 1);
    pl.long_n = n > MAX_N;
    // long arrays: one wave of composition CTAs covers all permutations
    pl.pb = pl.long_n ? std::max(8, (P + sm_count() - 1) / sm_count()) : PB;
    pl.n_blocks = (P + pl.pb - 1) / pl.pb;
    const size_t nw = (size_t)(n + 2) / 2;
    pl.lists_in_smem = (((nw + 3) & ~size_t(3)) * 4 + (size_t)n * 2) <= SMEM_MAX - 1024;
    pl.rows_in_smem = (size_t)((n + 7) & ~7) * 4 <= SMEM_MAX - 1024;
    pl.apply_grid = std::min(P, sm_count());
    size_t o = 256;
    pl.off_tab = o; o += up(pl.long_n ? 0 : (size_t)pl.n_chunks * (n - 1) * 2);   // uint16 accepts per (chunk, start state)
    pl.off_m = o;   o += up((size_t)(pl.n_chunks + 1) * 4);       // start state of every chunk
    pl.off_s = o;   o += up((size_t)(pl.n_chunks + 1) * 8);       // global step offset of every chunk
    pl.off_j = o;   o += up((size_t)pl.steps * 4);                // j of every step
    pl.off_arr = o; o += up(pl.long_n ? 0 : (size_t)n * P * 4);   // [i][P] working arrays of stage E
    pl.off_pi = o;  o += up((size_t)n * P * 4);                   // [P][n] pi_p
    pl.off_c = o;   o += up((size_t)pl.n_blocks * n * 4);         // block composites
    pl.off_d = o;   o += up((size_t)(pl.n_blocks + 1) * n * 4);   // prefixes of the block composites
    pl.off_g = o;   o += up(pl.long_n && !pl.rows_in_smem ? (size_t)(pl.n_blocks + 1) * 2 * n * 4 : 0);  // rows in L2
    pl.off_l = o;   o += up(pl.long_n && !pl.lists_in_smem ? (size_t)pl.apply_grid * n * 2 : 0);        // step lists in L2
    // LONG emit records: per (permutation, epoch, thread) the accept mask and the exact base, per (permutation, epoch)
    // the window's first word and its geometry — the j values are extracted from them by all SMs afterwards
    pl.n_ep = pl.long_n ? long_epochs(n) : 0;
    const size_t recs = (size_t)P * pl.n_ep;
    pl.off_rm = o;  o += up(recs * LT * 8);
    pl.off_rb = o;  o += up(recs * LT * 2);
    pl.off_eh = o;  o += up(recs * 16);
    pl.total = o;
    return pl;
}

__device__ __forceinline__ bool step(uint32_t w, int& m, int n1, uint32_t& v) {
    const uint32_t mask = 0xffffffffu >> __clz(m);   // smallest 2^k - 1 >= m  (m >= 1)
    v = w & mask;
    const bool acc = v <= (uint32_t)m;
    if (acc) m = (m == 1) ? n1 : m - 1;              // the step with max 1 is a permutation's last
    return acc;
}

// B: tab[c][pos] = accepts of chunk c when it starts at position pos (state m = n1 - pos), for every pos in [0, n1).
// The distinct runs live in shared memory as parallel arrays sorted by start position: g = first start position of the
// run's group, m = current state, a = accepts so far of the group's first member; every member `pos` of the group has
// accepts (g + a) - pos because merged runs sit at the same unrolled position u = pos + accepts.
__global__ void __launch_bounds__(CT)
tables_kernel(const uint32_t* __restrict__ words, long long n_words, int n1, uint16_t* __restrict__ tab) {
    extern __shared__ uint32_t sm32[];
    uint32_t* sw = sm32;                                     // [RW] words of the round
    uint32_t* su = sm32 + RW;                                // [CT] unrolled positions of a compaction strip
    uint16_t* sm = reinterpret_cast<uint16_t*>(su + CT);     // [n1] state
    uint16_t* sa = sm + ((n1 + 1) & ~1);                     // [n1] accepts
    uint16_t* sg = sa + ((n1 + 1) & ~1);                     // [n1] first start position of the group
    __shared__ int warp_cnt[CT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long c = blockIdx.x;
    const long long base_w = c * L;
    const int len = (int)min((long long)L, n_words - base_w);
    for (int i = tid; i < n1; i += CT) {
        sm[i] = (uint16_t)(n1 - i);
        sa[i] = 0;
        sg[i] = (uint16_t)i;
    }
    int D = n1;
    for (int r0 = 0; r0 < len; r0 += RW) {
        const int tl = min(RW, len - r0);
        __syncthreads();
        for (int t = tid; t < tl; t += CT) sw[t] = words[base_w + r0 + t];
        __syncthreads();
        // ---- advance every distinct run through the round's words
        for (int i = tid; i < D; i += CT) {
            int m = sm[i];
            uint32_t acc = sa[i], v;
#pragma unroll 4
            for (int t = 0; t < tl; ++t) acc += step(sw[t], m, n1, v) ? 1u : 0u;
            sm[i] = (uint16_t)m;
            sa[i] = (uint16_t)acc;
        }
        __syncthreads();
        if (D == 1) continue;
        // ---- compact in place: neighbours at the same unrolled position have merged for good
        int out = 0;
        uint32_t prev_u = 0xffffffffu;
        for (int s0 = 0; s0 < D; s0 += CT) {
            const int i = s0 + tid;
            const bool valid = i < D;
            const uint32_t mm = valid ? sm[i] : 0u, aa = valid ? sa[i] : 0u, gg = valid ? sg[i] : 0u;
            const uint32_t u = gg + aa;
            su[tid] = u;
            __syncthreads();
            const uint32_t before = tid > 0 ? su[tid - 1] : prev_u;
            const uint32_t strip_last = su[min(CT, D - s0) - 1];
            const bool head = valid && u != before;
            const unsigned bal = __ballot_sync(0xffffffffu, head);
            if (lane == 0) warp_cnt[warp] = __popc(bal);
            __syncthreads();
            int off = __popc(bal & ((1u << lane) - 1u)), total = 0;
#pragma unroll
            for (int w = 0; w < CT / 32; ++w) {
                const int cw = warp_cnt[w];
                if (w < warp) off += cw;
                total += cw;
            }
            if (head) {  // new index <= old index, and every element of this strip has been read
                sm[out + off] = (uint16_t)mm;
                sa[out + off] = (uint16_t)aa;
                sg[out + off] = (uint16_t)gg;
            }
            out += total;
            prev_u = strip_last;
            __syncthreads();
        }
        D = out;
    }
    __syncthreads();
    // ---- expand: every start position takes its group's unrolled end position
    for (int pos = tid; pos < n1; pos += CT) {
        int lo = 0, hi = D;  // largest group with sg <= pos
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if ((int)sg[mid] <= pos) lo = mid; else hi = mid;
        }
        tab[c * n1 + pos] = (uint16_t)((int)sg[lo] + (int)sa[lo] - pos);
    }
}

// C: start state / step offset of every chunk; end_chunk = the chunk in which the last step falls (-1: too few words)
__global__ void chain_kernel(const uint16_t* __restrict__ tab, long long n_chunks, int n1, long long steps,
                             int32_t* __restrict__ start_m, long long* __restrict__ start_s,
                             long long* __restrict__ info) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int m = n1;
    long long s = 0, end_chunk = -1;
    for (long long c = 0; c < n_chunks; ++c) {
        start_m[c] = m;
        start_s[c] = s;
        const int a = tab[c * n1 + (n1 - m)];
        s += a;
        if (s >= steps) {
            end_chunk = c;
            break;
        }
        m = (int)(((long long)(m - 1) - a) % n1);
        if (m < 0) m += n1;
        m += 1;
    }
    info[0] = end_chunk;
    info[1] = -1;  // words consumed, written by the emit kernel
}

// D: replay every chunk from its start state and write the accepted values in step order. One warp per chunk: the
// run itself is one dependent chain (lane 0), the other lanes stage the chunk's words in shared memory tile by tile so
// that the chain never waits for global memory.
constexpr int ET = 1024;  // words per staged tile
__global__ void __launch_bounds__(32)
emit_kernel(const uint32_t* __restrict__ words, long long n_words, int n1, long long steps,
            const int32_t* __restrict__ start_m, const long long* __restrict__ start_s, long long* __restrict__ info,
            int32_t* __restrict__ jout) {
    __shared__ uint32_t sw[2][ET];
    const long long c = blockIdx.x;
    const int lane = threadIdx.x;
    const long long end_chunk = info[0];
    if (end_chunk < 0 || c > end_chunk) return;
    const long long base = c * L;
    const int len = (int)min((long long)L, n_words - base);
    int m = start_m[c];
    long long s = start_s[c];
    uint32_t v;
    bool done = false;
    for (int t = lane; t < min(ET, len); t += 32) sw[0][t] = words[base + t];
    __syncwarp();
    for (int t0 = 0, buf = 0; t0 < len && !done; t0 += ET, buf ^= 1) {
        const int tl = min(ET, len - t0);
        if (lane > 0) {  // 31 lanes fetch the next tile while lane 0 walks this one
            const int nl = min(ET, len - (t0 + ET));
            for (int t = lane - 1; t < nl; t += 31) sw[buf ^ 1][t] = words[base + t0 + ET + t];
        } else {
            for (int t = 0; t < tl; ++t) {
                if (step(sw[buf][t], m, n1, v)) {
                    jout[s] = (int32_t)v;
                    if (++s == steps) {
                        info[1] = base + t0 + t + 1;  // exactly this many words were consumed
                        done = true;
                        break;
                    }
                }
            }
        }
        done = __shfl_sync(0xffffffffu, done, 0);
        __syncwarp();
    }
}

// D-long: the accepted j values of P cumulative shuffles of an array too long for the table scheme (n > MAX_N), by ONE
// CTA that is parallel WITHIN a permutation (numpy emulation and proof of exactness: dev/emul_emit_long2.py).
// Inside a mask epoch (all steps i with the same mask M, i from i_hi down to i_lo; c = i_hi - i_lo + 1 steps) the k-th
// accepted word serves step i_hi - k, so word t is accepted iff  v_t + a(t) <= i_hi,  v_t = word_t & M,  a(t) = accepts
// before t. a(t) is the only serial quantity. Every thread owns S consecutive words of the epoch's window and keeps a
// bracket l <= a(t) <= u along them, updated word by word (Gauss-Seidel inside the segment):
//     l += [v + u <= i_hi]   (surely accepted)        u += [v + l <= i_hi]   (not surely rejected)
// starting from the bracket (Lb, Ub) of the segment's first word; one block-wide scan of the packed (l - Lb, u - Ub)
// counts per round gives the next round's bases (Jacobi across segments). Thread 0's base is exact (0, 0), an exact
// base makes the walk the serial simulation itself, and the brackets contract like tau^r / r! (tau = window / mask),
// so every thread below the epoch's end is exact after ~5 rounds (one barrier each); then one more walk emits.
// The raw words live in a shared-memory ring that cp.async refills a whole epoch ahead, the last 63 steps of a
// permutation (masks below 64: a handful of words each) are walked by one thread.
constexpr int RING = 57344;             // words of the ring (224 KB)
constexpr int CLASSIFY_MIN_S = 9;       // segments shorter than this are simply walked again every round
constexpr int RING_PAD = 64;            // copy of the ring's first words behind its end: a thread's segment never wraps
constexpr int WCAP = RING - 8;          // longest window of one epoch
constexpr int TAIL_W = 4096;            // words the warp tail may look at
#ifndef MMX_S_MIN
#define MMX_S_MIN 1
#endif
constexpr int S_MIN = MMX_S_MIN;        // fewest words per thread: small epochs run on few warps (cheaper rounds)

struct LongEpoch {
    int i_hi, c, W0;
    uint32_t M;
};
struct LongPlan {
    int n_epochs, i_tail;
    LongEpoch e[MAX_EPOCHS];
};

static LongPlan make_long_plan(int n) {
    LongPlan pl{};
    int i_hi = n - 1;
    while (i_hi >= SERIAL_BELOW && pl.n_epochs < MAX_EPOCHS) {
        int b = 0;
        while ((1ll << b) <= i_hi) ++b;                  // bit length of i_hi
        const uint32_t M = (uint32_t)((1ull << b) - 1);
        const int i_lo = std::max(1 << (b - 1), SERIAL_BELOW);
        double mean = 0.0, var = 0.0;
        for (int i = i_lo; i <= i_hi; ++i) {
            const double p = (i + 1.0) / ((double)M + 1.0);
            mean += 1.0 / p;
            var += (1.0 - p) / (p * p);
        }
        LongEpoch& e = pl.e[pl.n_epochs++];
        e.i_hi = i_hi;
        e.c = i_hi - i_lo + 1;
        e.M = M;
        e.W0 = (int)std::min<double>(WCAP, mean + 8.0 * std::sqrt(var) + 64.0);
        i_hi = i_lo - 1;
    }
    pl.i_tail = i_hi;
    return pl;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

#ifdef MMX_EMIT_TIMING
// dev/emit_timing.cu: cycles per epoch and phase, accumulated by thread 0  [epoch][stage, rounds, emit, #rounds]
__device__ long long g_emit_prof[MAX_EPOCHS + 1][4];
__device__ long long g_emit_round[MAX_EPOCHS + 1][4];   // thread 0, summed over rounds: walk, scan, barrier, base
#define EMIT_RADD(ep, k, val) \
    if (threadIdx.x == 0) g_emit_round[ep][k] += (val)
#define EMIT_CLOCK(v) const long long v = clock64()
#define EMIT_ADD(ep, k, val) \
    if (threadIdx.x == 0) g_emit_prof[ep][k] += (val)
#else
#define EMIT_CLOCK(v)
#define EMIT_ADD(ep, k, val)
#define EMIT_RADD(ep, k, val)
#endif

// barrier 1 over the first `nthreads` threads of the CTA (a multiple of 32) with an AND reduction of `pred`
__device__ __forceinline__ int bar_and_active(int pred, int nthreads) {
    int r;
    asm volatile(
        "{\n .reg .pred p, q;\n setp.ne.s32 q, %1, 0;\n barrier.cta.red.and.pred p, 1, %2, q;\n selp.s32 %0, 1, 0, p;\n}"
        : "=r"(r)
        : "r"(pred), "r"(nthreads)
        : "memory");
    return r;
}

// The walks below run over ONE thread's segment: `n` consecutive words at w (the ring keeps a copy of its first RING_PAD
// words behind its end, so a segment never wraps). Word k is accepted iff the accepts before it (a) are <= i_hi - (w[k] & M).
// two independent runs at once (the lower and the upper end of a bracket)
__device__ __forceinline__ void walk_count2(const uint32_t* __restrict__ w, int n, uint32_t M, int i_hi, int& a, int& b) {
#pragma unroll 4
    for (int k = 0; k < n; ++k) {
        const int d = i_hi - (int)(w[k] & M);
        a += (a <= d) ? 1 : 0;
        b += (b <= d) ? 1 : 0;
    }
}
// two runs, recording the accept masks of both (32 words per mask half)
__device__ __forceinline__ void walk_count2_masks(const uint32_t* __restrict__ w, int n, uint32_t M, int i_hi, int& a, int& b,
                                                  unsigned long long& ma, unsigned long long& mb) {
    uint32_t a0 = 0, b0 = 0, a1 = 0, b1 = 0;
    const int n0 = min(n, 32);
#pragma unroll 4
    for (int k = 0; k < n0; ++k) {
        const int d = i_hi - (int)(w[k] & M);
        const uint32_t acca = (a <= d) ? 1u : 0u, accb = (b <= d) ? 1u : 0u;
        a += acca;
        b += accb;
        a0 |= acca << k;
        b0 |= accb << k;
    }
#pragma unroll 4
    for (int k = 32; k < n; ++k) {
        const int d = i_hi - (int)(w[k] & M);
        const uint32_t acca = (a <= d) ? 1u : 0u, accb = (b <= d) ? 1u : 0u;
        a += acca;
        b += accb;
        a1 |= acca << (k - 32);
        b1 |= accb << (k - 32);
    }
    ma = (unsigned long long)a0 | ((unsigned long long)a1 << 32);
    mb = (unsigned long long)b0 | ((unsigned long long)b1 << 32);
}
// the exact run: also records which words were accepted
__device__ __forceinline__ int walk_mask(const uint32_t* __restrict__ w, int n, uint32_t M, int i_hi, int a,
                                         unsigned long long& m) {
    uint32_t m0 = 0, m1 = 0;
    const int n0 = min(n, 32);
#pragma unroll 4
    for (int k = 0; k < n0; ++k) {
        const int d = i_hi - (int)(w[k] & M);
        const uint32_t acc = (a <= d) ? 1u : 0u;
        a += acc;
        m0 |= acc << k;
    }
#pragma unroll 4
    for (int k = 32; k < n; ++k) {
        const int d = i_hi - (int)(w[k] & M);
        const uint32_t acc = (a <= d) ? 1u : 0u;
        a += acc;
        m1 |= acc << (k - 32);
    }
    m = (unsigned long long)m0 | ((unsigned long long)m1 << 32);
    return a;
}

__global__ void __launch_bounds__(LT, 1)
emit_long_kernel(const uint32_t* __restrict__ words, long long n_words, int n, int P,
                 const __grid_constant__ LongPlan plan, long long* __restrict__ info, int32_t* __restrict__ jout,
                 unsigned long long* __restrict__ rmask, uint16_t* __restrict__ rbase, long long* __restrict__ ehead) {
    extern __shared__ uint32_t ring[];  // [RING + RING_PAD] word g of the stream sits at ring[g % RING]
    __shared__ uint32_t wtot[2][LT / 32];
    __shared__ int s_end;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    n_words &= ~3ll;                    // 16-byte copies only
    long long pos = 0;                  // first word not consumed yet
    long long loaded = 0;               // words [.., loaded) are in the ring or on their way
    int pos_mod = 0, loaded_mod = 0;    // pos % RING, loaded % RING
    int err = 0;

    // refill the ring up to pos + RING (the slots of consumed words), wait until [pos, pos + W) has landed
    auto stage = [&](int W) {
        __syncthreads();  // everybody is done reading the consumed words (and has seen s_end)
        if (tid == 0) s_end = -1;
        const long long until_prev = loaded;
        const long long until = min(n_words, (pos + RING) & ~3ll);
        if (until > loaded) {
            const int cnt = (int)(until - loaded);
            int r = loaded_mod + 4 * tid;
            if (r >= RING) r -= RING;
            const uint32_t* src = words + loaded + 4 * tid;
            for (int g = 4 * tid; g < cnt; g += 4 * LT) {
                cp_async16(ring + r, src);
                if (r < RING_PAD) cp_async16(ring + RING + r, src);
                src += 4 * LT;
                r += 4 * LT;
                if (r >= RING) r -= RING;
            }
            loaded_mod = (loaded_mod + cnt) % RING;
            loaded = until;
        }
        cp_async_commit();
        if (pos + W <= until_prev) cp_async_wait<1>();  // this epoch's refill stays in flight
        else cp_async_wait<0>();
        __syncthreads();
    };

    for (int p = 0; p < P && !err; ++p) {
        int32_t* const jp = jout + (long long)p * (n - 1);
        for (int e = 0; e < plan.n_epochs && !err; ++e) {
            const int i_hi = plan.e[e].i_hi, c = plan.e[e].c;
            const uint32_t M = plan.e[e].M;
            int W = plan.e[e].W0;
            for (;;) {
                W = (int)min((long long)W, n_words - pos);
                if (W < c) {  // the stream is too short: the caller retries with more words
                    err = -1;
                    break;
                }
                EMIT_CLOCK(tc0);
                stage(W);
                EMIT_CLOCK(tc1);
                EMIT_ADD(e, 0, tc1 - tc0);
                int S = max((W + LT - 1) / LT, S_MIN);
                S |= 1;  // odd: the threads of a warp walk distinct banks
                const int n_act = (((W + S - 1) / S + 31) >> 5) << 5;  // threads that own words, whole warps
                if (tid < n_act) {
                    const int t0 = min(W, tid * S), len = min(W, t0 + S) - t0;
                    int r0 = pos_mod + t0;
                    if (r0 >= RING) r0 -= RING;
                    const uint32_t* seg = ring + r0;  // contiguous thanks to the ring's padding
                    int Lb = 0, Ub = t0, cnt = 0;
                    bool have_cnt = false, have_mask = false;  // the segment's accept count / accept mask is final
                    unsigned long long acc_mask = 0;
                    // classified segment: every word is surely accepted (mask sure_m, sure_n of them), surely rejected, or
                    // one of at most UNC uncertain words e[j] = theta | k << 16 (word k accepted iff start + uncertain
                    // accepts before it <= theta)
                    constexpr int UNC = 4;
                    bool cls = false;
                    unsigned long long sure_m = 0;
                    int sure_n = 0, un = 0;
                    uint32_t e0 = 0, e1 = 0, e2 = 0, e3 = 0;
                    for (int rnd = 0;; ++rnd) {
                        EMIT_CLOCK(ra);
                        uint32_t pk;  // low half: a lower bound of this segment's accepts, high half: an upper bound
                        if (Lb >= c) {
                            pk = (uint32_t)len << 16;  // behind the epoch's end: nothing depends on it
                        } else if (have_cnt) {
                            pk = (uint32_t)cnt | ((uint32_t)cnt << 16);
                        } else if (cls) {
                            // two runs over the uncertain words only
                            int xa = Ub, xb = Lb;
                            const uint32_t ee[UNC] = {e0, e1, e2, e3};
#pragma unroll
                            for (int j = 0; j < UNC; ++j)
                                if (j < un) {
                                    const int th = (int)(ee[j] & 0xffffu);
                                    xa += (xa <= th) ? 1 : 0;
                                    xb += (xb <= th) ? 1 : 0;
                                }
                            const int cl = sure_n + xa - Ub, cu = sure_n + xb - Lb;
                            pk = (uint32_t)cl | ((uint32_t)cu << 16);
                            if (cl == cu) {  // the same uncertain words are accepted from both ends: final
                                acc_mask = sure_m;
                                int x = Lb;
#pragma unroll
                                for (int j = 0; j < UNC; ++j)
                                    if (j < un && x <= (int)(ee[j] & 0xffffu)) {
                                        acc_mask |= 1ull << (ee[j] >> 16);
                                        ++x;
                                    }
                                cnt = cl;
                                have_cnt = have_mask = true;
                            }
                        } else if (rnd == 0 || S < CLASSIFY_MIN_S) {
                            // accepts are non-increasing in the start value: a run from Ub bounds them from below, a
                            // run from Lb from above. (Every lane of the warp takes the same kind of walk in a round:
                            // a warp that mixes walk kinds pays for each of them.)
                            int lo = Ub, hi = Lb;
                            walk_count2(seg, len, M, i_hi, lo, hi);
                            pk = (uint32_t)(lo - Ub) | ((uint32_t)(hi - Lb) << 16);
                            if (rnd > 0 && lo - Ub == hi - Lb) {  // short segments: frozen as soon as both runs agree
                                cnt = hi - Lb;
                                have_cnt = true;
                            }
                        } else {
                            // the brackets are narrow from the second round on. A run from anywhere inside the bracket
                            // stays between the run from Lb and the run from Ub: words both accept are accepted from
                            // everywhere in the bracket (and every narrower one: the brackets are nested), words both
                            // reject are rejected, the FEW others are kept with their thresholds — no full walk again
                            int lo = Ub, hi = Lb;
                            unsigned long long mU = 0, mL = 0;
                            walk_count2_masks(seg, len, M, i_hi, lo, hi, mU, mL);
                            pk = (uint32_t)(lo - Ub) | ((uint32_t)(hi - Lb) << 16);
                            unsigned long long unc = mL & ~mU;
                            if (unc == 0ull) {
                                acc_mask = mU;
                                cnt = hi - Lb;
                                have_cnt = have_mask = true;
                            } else if (__popcll(unc) <= UNC) {
                                cls = true;
                                sure_m = mU;
                                sure_n = lo - Ub;
                                un = 0;
                                while (unc != 0ull) {
                                    const int k = __ffsll((long long)unc) - 1;
                                    unc &= unc - 1ull;
                                    const int d = i_hi - (int)(seg[k] & M);
                                    const uint32_t ent =
                                        (uint32_t)(d - __popcll(mU & ((1ull << k) - 1ull))) | ((uint32_t)k << 16);
                                    if (un == 0) e0 = ent;
                                    else if (un == 1) e1 = ent;
                                    else if (un == 2) e2 = ent;
                                    else e3 = ent;
                                    ++un;
                                }
                            }
                        }
                        EMIT_CLOCK(rb);
                        uint32_t incl = pk;  // inclusive scan of both counts at once (each stays below 2^16)
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                            if (lane >= o) incl += t;
                        }
                        if (lane == 31) wtot[rnd & 1][warp] = incl;
                        const int settled = (Lb >= c) || (Lb == Ub) || (len == 0);
                        EMIT_CLOCK(rc);
                        const int all_settled = bar_and_active(settled, n_act);
                        EMIT_CLOCK(rd);
                        EMIT_RADD(e, 0, rb - ra);
                        EMIT_RADD(e, 1, rc - rb);
                        EMIT_RADD(e, 2, rd - rc);
                        if (all_settled) break;
                        const uint32_t base = __reduce_add_sync(0xffffffffu, lane < warp ? wtot[rnd & 1][lane] : 0u)
                                              + incl - pk;
                        Lb = max(Lb, (int)(base & 0xffffu));  // every round's bounds are valid: keep the brackets nested
                        Ub = min(Ub, (int)(base >> 16));
                        EMIT_ADD(e, 3, 1);
                        EMIT_CLOCK(re);
                        EMIT_RADD(e, 3, re - rd);
                    }
                    EMIT_CLOCK(tc2);
                    EMIT_ADD(e, 1, tc2 - tc1);
                    if (Lb < c && !have_mask) {  // frozen before the masks were kept: one exact walk
                        cnt = walk_mask(seg, len, M, i_hi, Lb, acc_mask) - Lb;
                    }
                    // ---- record: the j values themselves (the k-th accepted word of the epoch is the j of step
                    // i_hi - k) are pulled out of the word stream by all SMs afterwards (extract_long_kernel) from this
                    // thread's accept mask and exact base — 10 coalesced bytes instead of ~20 scattered stores
                    {
                        const size_t rec = ((size_t)p * plan.n_epochs + e) * LT + tid;
                        rmask[rec] = acc_mask;
                        rbase[rec] = (uint16_t)min(Lb, 0xffff);
                        if (tid == 0) {
                            ehead[2 * ((size_t)p * plan.n_epochs + e)] = pos;
                            ehead[2 * ((size_t)p * plan.n_epochs + e) + 1] = (long long)S | ((long long)n_act << 32);
                        }
                    }
                    if (Lb < c && Lb + cnt >= c) {  // this segment holds the epoch's last accept
                        int need = c - Lb, k = 0;
                        for (;; ++k)
                            if ((acc_mask >> k) & 1ull)
                                if (--need == 0) break;
                        s_end = t0 + k + 1;
                    }
                }
                __syncthreads();
                EMIT_CLOCK(tc3);
                EMIT_ADD(e, 2, tc3 - tc1);
                const int end = s_end;
                if (end >= 0) {
                    pos += end;
                    pos_mod += end;
                    if (pos_mod >= RING) pos_mod -= RING;
                    break;
                }
                // the c-th accept is not inside the window: stage a longer one
                if (W >= WCAP || pos + W >= n_words) {
                    err = (W >= WCAP) ? -3 : -1;
                    break;
                }
                W = min(WCAP, 2 * W);
            }
        }
        if (err) break;
        // ---- warp tail: the steps i < SERIAL_BELOW (about 1.45 words each), 32 words at a time by warp 0. Lane k's word
        // serves step i - (accepts among the lanes before it): the accept bits are the fixed point of
        //     A = ballot( (w & mask(i - popc(A & lanes_before))) <= i - popc(A & lanes_before) ),
        // reached from any start because lane k only depends on lanes < k (a few rounds; the masks change with i, so
        // this also walks across the mask epochs below 512 without any staging)
        if (plan.i_tail >= 1) {
            const int W = (int)min((long long)TAIL_W, n_words - pos);
            EMIT_CLOCK(tt0);
            stage(W);
            EMIT_CLOCK(tt1);
            EMIT_ADD(MAX_EPOCHS, 0, tt1 - tt0);
            if (warp == 0) {
                const uint32_t lanes_before = (1u << lane) - 1u;
                int i = plan.i_tail, k0 = 0, end = -1;
                while (i >= 1 && k0 + 32 <= W) {
                    const int r = pos_mod + k0 + lane;
                    const uint32_t w = ring[r >= RING ? r - RING : r];
                    uint32_t A = 0xffffffffu, prevA;
                    uint32_t v = 0;
                    int il, acc;
                    do {
                        prevA = A;
                        il = i - __popc(A & lanes_before);
                        acc = 0;
                        if (il >= 1) {
                            v = w & (0xffffffffu >> __clz(il));
                            acc = v <= (uint32_t)il;
                        }
                        A = __ballot_sync(0xffffffffu, acc);
                    } while (A != prevA);
                    if (acc) jp[n - 1 - il] = (int32_t)v;
                    const int tot = __popc(A);
                    if (tot >= i) {  // the permutation's last step (i = 1) was served by the highest accepting lane
                        end = k0 + (31 - __clz(A)) + 1;
                        i = 0;
                    } else {
                        i -= tot;
                        k0 += 32;
                    }
                }
                if (lane == 0) s_end = end;
            }
            __syncthreads();
            EMIT_CLOCK(tt2);
            EMIT_ADD(MAX_EPOCHS, 2, tt2 - tt1);
            const int end = s_end;
            if (end < 0) {
                err = (W >= TAIL_W) ? -3 : -1;
                break;
            }
            pos += end;
            pos_mod += end;
            if (pos_mod >= RING) pos_mod -= RING;
        }
    }
    if (tid == 0) {
        info[0] = err ? -1 : 0;
        info[1] = err ? err : pos;  // exactly this many words were consumed
    }
}

// D-long, second half: the j values of the bracket epochs from the records of emit_long_kernel. One CTA per
// (permutation, epoch); thread t re-reads its segment's accepted words (bit k of its mask = word t S + k of the window)
// and writes them from its exact base on. All SMs, HBM/L2-bound.
__global__ void __launch_bounds__(LT)
extract_long_kernel(const uint32_t* __restrict__ words, int n, const __grid_constant__ LongPlan plan,
                    const long long* __restrict__ info, const unsigned long long* __restrict__ rmask,
                    const uint16_t* __restrict__ rbase, const long long* __restrict__ ehead, int32_t* __restrict__ jout) {
    if (info[1] < 0) return;  // too few words: the records are incomplete, the caller retries
    const int e = blockIdx.x % plan.n_epochs;
    const long long p = blockIdx.x / plan.n_epochs;
    const long long pos = ehead[2 * (size_t)blockIdx.x];
    const long long geo = ehead[2 * (size_t)blockIdx.x + 1];
    const int S = (int)(geo & 0xffffffffll), n_act = (int)(geo >> 32);
    const int tid = threadIdx.x;
    if (tid >= n_act) return;
    const int c = plan.e[e].c, i_hi = plan.e[e].i_hi;
    const uint32_t M = plan.e[e].M;
    const int Lb = rbase[(size_t)blockIdx.x * LT + tid];
    if (Lb >= c) return;
    unsigned long long m = rmask[(size_t)blockIdx.x * LT + tid];
    const uint32_t* w = words + pos + (long long)tid * S;
    int32_t* out = jout + p * (n - 1) + (n - 1 - i_hi) + Lb;
    int room = c - Lb;  // accepts beyond the epoch's last step belong to nobody
    while (m != 0ull && room > 0) {
        const int k = __ffsll((long long)m) - 1;
        m &= m - 1ull;
        *out++ = (int32_t)(w[k] & M);
        --room;
    }
}

// E-long: permutation p (the result of its n-1 swaps on an identity array) without replaying the swaps one by one.
// Step i (i = n-1 .. 1) exchanges positions i and j_i <= i, after which position i is final. The value a position q
// holds just before step i was put there by the most recent earlier step that targeted q — the smallest i' > i with
// j_i' = q — and is whatever position i' held just before step i'; without such a step it is q itself. So
//     pi[i] = root(first(j_i, i)),  first(q, i) = min{ i' > i : j_i' = q },  root(x) = follow first(x, x) to the end,
// a pointer chase of expected depth O(log n) over the steps grouped by target (CSR built with packed 16-bit shared
// atomics). One CTA per permutation; every position is resolved independently.
constexpr int AT = 1024;
__global__ void __launch_bounds__(AT, 1)
apply_long_kernel(const int32_t* __restrict__ jout, int n, int P, const long long* __restrict__ info,
                  int32_t* __restrict__ pi, uint16_t* __restrict__ glist) {
    extern __shared__ uint32_t asm32[];
    // cur[q] (uint16, packed two per word): while filling, the cursor of target q; afterwards the END of q's list
    uint32_t* cur = asm32;                                    // [(n + 2) / 2] words
    __shared__ uint32_t wsum[AT / 32];
    if (info[1] < 0) return;  // too few words: jout is incomplete, the caller retries
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nw = (n + 2) / 2;
    uint16_t* lst = glist != nullptr ? glist + (size_t)blockIdx.x * n
                                     : reinterpret_cast<uint16_t*>(asm32 + ((nw + 3) & ~3));
    const uint16_t* cur16 = reinterpret_cast<const uint16_t*>(cur);
    for (int p = blockIdx.x; p < P; p += gridDim.x) {
        const int32_t* jp = jout + (long long)p * (n - 1);  // jp[n - 1 - i] = j_i
        __syncthreads();
        for (int w = tid; w < nw; w += AT) cur[w] = 0;
        __syncthreads();
        for (int i = 1 + tid; i < n; i += AT) {
            const int q = jp[n - 1 - i];
            atomicAdd(&cur[q >> 1], 1u << ((q & 1) * 16));
        }
        __syncthreads();
        // exclusive scan of the n 16-bit counts, in place (every thread owns a contiguous run of words)
        {
            const int per = (nw + AT - 1) / AT;
            const int w0 = min(nw, tid * per), w1 = min(nw, w0 + per);
            uint32_t s = 0;
            for (int w = w0; w < w1; ++w) {
                const uint32_t x = cur[w];
                s += (x & 0xffffu) + (x >> 16);
            }
            uint32_t incl = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) wsum[warp] = incl;
            __syncthreads();
            uint32_t ws = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, ws, o);
                if (lane >= o) ws += t;
            }
            uint32_t run = __shfl_sync(0xffffffffu, ws, (warp + 31) & 31);
            if (warp == 0) run = 0;
            run += incl - s;
            for (int w = w0; w < w1; ++w) {
                const uint32_t x = cur[w];
                const uint32_t lo = run, hi = run + (x & 0xffffu);
                run = hi + (x >> 16);
                cur[w] = (lo & 0xffffu) | (hi << 16);  // offsets stay below n - 1 <= 65535
            }
        }
        __syncthreads();
        for (int i = 1 + tid; i < n; i += AT) {
            const int q = jp[n - 1 - i];
            const uint32_t old = atomicAdd(&cur[q >> 1], 1u << ((q & 1) * 16));
            lst[(old >> ((q & 1) * 16)) & 0xffffu] = (uint16_t)i;
        }
        __syncthreads();
        // list of target q = lst[begin(q) .. cur16[q]), begin(q) = cur16[q - 1] (0 for q = 0)
        int32_t* row = pi + (long long)p * n;
        for (int x = tid; x < n; x += AT) {
            int q = x >= 1 ? jp[n - 1 - x] : 0;
            int above = x;  // look for the smallest step > above that targeted q
            for (;;) {
                const int b = q > 0 ? cur16[q - 1] : 0, e = cur16[q];
                int best = 0x7fffffff;
                for (int k = b; k < e; ++k) {
                    const int s = lst[k];
                    if (s > above && s < best) best = s;
                }
                if (best == 0x7fffffff) break;
                q = best;
                above = best;
            }
            row[x] = q;
        }
    }
}

// F-long: the same three composition phases as below with the two working rows in shared memory as uint16
// (n <= 56 000: 2 x 2 n bytes), PB_long permutations per block so that one wave of CTAs covers all of them.
__global__ void __launch_bounds__(1024, 1)
compose_block16_kernel(const int32_t* __restrict__ pi, int n, int P, int pb, const int32_t* __restrict__ start,
                       int32_t* __restrict__ comp_out, int32_t* __restrict__ rows_out) {
    extern __shared__ uint16_t sm16[];
    const int b = blockIdx.x;
    const int np = (n + 7) & ~7;
    uint16_t* cur = sm16;
    uint16_t* nxt = sm16 + np;
    for (int k = threadIdx.x; k < n; k += blockDim.x)
        cur[k] = (uint16_t)(start != nullptr ? start[(long long)b * n + k] : k);
    __syncthreads();
    const int p_end = min(P, (b + 1) * pb);
    for (int p = b * pb; p < p_end; ++p) {
        const int32_t* row = pi + (long long)p * n;
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            const uint16_t x = cur[row[k]];
            nxt[k] = x;
            if (rows_out != nullptr) rows_out[(long long)p * n + k] = x;
        }
        __syncthreads();
        uint16_t* t = cur;
        cur = nxt;
        nxt = t;
    }
    if (comp_out != nullptr)
        for (int k = threadIdx.x; k < n; k += blockDim.x) comp_out[(long long)b * n + k] = cur[k];
}

__global__ void __launch_bounds__(1024, 1)
compose_scan16_kernel(const int32_t* __restrict__ comp, int n, int n_blocks, int32_t* __restrict__ prefix) {
    extern __shared__ uint16_t sm16[];
    const int np = (n + 7) & ~7;
    uint16_t* cur = sm16;
    uint16_t* nxt = sm16 + np;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        cur[k] = (uint16_t)k;
        prefix[k] = k;
    }
    __syncthreads();
    for (int b = 0; b < n_blocks; ++b) {
        const int32_t* c = comp + (long long)b * n;
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            const uint16_t x = cur[c[k]];
            nxt[k] = x;
            prefix[(long long)(b + 1) * n + k] = x;
        }
        __syncthreads();
        uint16_t* t = cur;
        cur = nxt;
        nxt = t;
    }
}

// E: thread p applies the swaps of permutation p to an identity array held as column p of arr[i][P]
__global__ void __launch_bounds__(128)
apply_kernel(const int32_t* __restrict__ jout, int n, int P, const long long* __restrict__ info,
             int32_t* __restrict__ arr) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P || info[1] < 0) return;  // too few words: jout is incomplete, the caller retries
    for (int i = 0; i < n; ++i) arr[(long long)i * P + p] = i;
    const int32_t* jp = jout + (long long)p * (n - 1);
    for (int i = n - 1; i >= 1; --i) {
        const int j = jp[n - 1 - i];
        const int32_t a = arr[(long long)i * P + p], b = arr[(long long)j * P + p];
        arr[(long long)i * P + p] = b;
        arr[(long long)j * P + p] = a;
    }
}

// [n][P] -> [P][n] through a shared-memory tile
__global__ void __launch_bounds__(256)
transpose_kernel(const int32_t* __restrict__ src, int n, int P, int32_t* __restrict__ dst) {
    __shared__ int32_t tile[32][33];
    const int i0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8)
        if (i0 + r < n && p0 + tx < P) tile[r][tx] = src[(long long)(i0 + r) * P + p0 + tx];
    __syncthreads();
    for (int r = ty; r < 32; r += 8)
        if (p0 + r < P && i0 + tx < n) dst[(long long)(p0 + r) * n + i0 + tx] = tile[tx][r];
}

// F1 / F3: CTA b composes the permutations of its block in order, starting from `start` (nullptr = identity);
// write_rows = 1 stores every intermediate row (the result), 0 stores only the block composite.
__global__ void __launch_bounds__(512)
compose_block_kernel(const int32_t* __restrict__ pi, int n, int P, int pb, const int32_t* __restrict__ start,
                     int32_t* __restrict__ comp_out, int32_t* __restrict__ rows_out, int32_t* gbuf) {
    extern __shared__ int32_t sm[];
    const int b = blockIdx.x;
    // long arrays (n > MAX_N) keep the two working rows in L2 (gbuf: 2 n words per block) instead of shared memory
    int32_t* cur = gbuf != nullptr ? gbuf + (size_t)b * 2 * n : sm;
    int32_t* nxt = cur + n;
    for (int k = threadIdx.x; k < n; k += blockDim.x) cur[k] = start != nullptr ? start[(long long)b * n + k] : k;
    __syncthreads();
    const int p_end = min(P, (b + 1) * pb);
    for (int p = b * pb; p < p_end; ++p) {
        const int32_t* row = pi + (long long)p * n;
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            const int32_t x = cur[row[k]];
            nxt[k] = x;
            if (rows_out != nullptr) rows_out[(long long)p * n + k] = x;
        }
        __syncthreads();
        int32_t* t = cur;
        cur = nxt;
        nxt = t;
    }
    if (comp_out != nullptr)
        for (int k = threadIdx.x; k < n; k += blockDim.x) comp_out[(long long)b * n + k] = cur[k];
}

// F2: prefix[b] = state before block b (prefix[0] = identity, prefix[b+1] = prefix[b] o comp[b])
__global__ void __launch_bounds__(512)
compose_scan_kernel(const int32_t* __restrict__ comp, int n, int n_blocks, int32_t* __restrict__ prefix,
                    int32_t* gbuf) {
    extern __shared__ int32_t sm[];
    int32_t* cur = gbuf != nullptr ? gbuf : sm;
    int32_t* nxt = cur + n;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        cur[k] = k;
        prefix[k] = k;
    }
    __syncthreads();
    for (int b = 0; b < n_blocks; ++b) {
        const int32_t* c = comp + (long long)b * n;
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            const int32_t x = cur[c[k]];
            nxt[k] = x;
            prefix[(long long)(b + 1) * n + k] = x;
        }
        __syncthreads();
        int32_t* t = cur;
        cur = nxt;
        nxt = t;
    }
}

}  // namespace shuf
}  // namespace mmx

using namespace mmx;
using namespace mmx::shuf;

extern "C" int mmx_perm_shuffle_workspace_bytes(int n, int P, int64_t n_words, size_t* bytes) {
    MMX_REQUIRE(bytes != nullptr && n >= 2 && n <= MAX_N_LONG && P >= 1 && n_words >= 1, "bad argument (2 <= n <= %d)",
                MAX_N_LONG);
    *bytes = make_plan(n, P, n_words).total;
    return 0;
}

extern "C" int mmx_perm_shuffle(const uint32_t* words, int64_t n_words, int n, int P, int32_t* perm, int64_t* info,
                                void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    MMX_REQUIRE(words != nullptr && perm != nullptr && info != nullptr && workspace != nullptr, "NULL argument");
    MMX_REQUIRE(n >= 2 && n <= MAX_N_LONG && P >= 1 && n_words >= 1, "bad argument (2 <= n <= %d)", MAX_N_LONG);
    const Plan pl = make_plan(n, P, n_words);
    MMX_REQUIRE(workspace_bytes >= pl.total, "workspace too small: need %zu bytes", pl.total);
    MMX_REQUIRE(pl.n_chunks < (1ll << 31), "too many chunks");
    uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
    uint16_t* tab = reinterpret_cast<uint16_t*>(ws + pl.off_tab - 256);
    int32_t* start_m = reinterpret_cast<int32_t*>(ws + pl.off_m - 256);
    long long* start_s = reinterpret_cast<long long*>(ws + pl.off_s - 256);
    int32_t* jout = reinterpret_cast<int32_t*>(ws + pl.off_j - 256);
    int32_t* arr = reinterpret_cast<int32_t*>(ws + pl.off_arr - 256);
    int32_t* pi = reinterpret_cast<int32_t*>(ws + pl.off_pi - 256);
    int32_t* comp = reinterpret_cast<int32_t*>(ws + pl.off_c - 256);
    int32_t* prefix = reinterpret_cast<int32_t*>(ws + pl.off_d - 256);
    int32_t* gbuf = pl.long_n ? reinterpret_cast<int32_t*>(ws + pl.off_g - 256) : nullptr;
    const int n1 = n - 1;
    if (pl.long_n) {
        // n > MAX_N: the emit is one CTA that is parallel within a permutation (bracketing prefix sums per mask epoch),
        // the swaps are resolved per position by a pointer chase, the composition keeps its rows in shared memory
        const LongPlan lp = make_long_plan(n);
        const size_t smem_long = (size_t)(RING + RING_PAD) * 4;
        static const char long_attr_tag = 0;
        if (first_use_on_device(&long_attr_tag)) {
            MMX_CHECK_CUDA(cudaFuncSetAttribute(emit_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)smem_long));
            MMX_CHECK_CUDA(cudaFuncSetAttribute(apply_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)(SMEM_MAX - 1024)));
            MMX_CHECK_CUDA(cudaFuncSetAttribute(compose_block16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)(SMEM_MAX - 1024)));
            MMX_CHECK_CUDA(cudaFuncSetAttribute(compose_scan16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)(SMEM_MAX - 1024)));
        }
        unsigned long long* rmask = reinterpret_cast<unsigned long long*>(ws + pl.off_rm - 256);
        uint16_t* rbase = reinterpret_cast<uint16_t*>(ws + pl.off_rb - 256);
        long long* ehead = reinterpret_cast<long long*>(ws + pl.off_eh - 256);
        MMX_REQUIRE(lp.n_epochs == pl.n_ep, "epoch plan mismatch");
        emit_long_kernel<<<1, LT, smem_long, stream>>>(words, n_words, n, P, lp, reinterpret_cast<long long*>(info), jout,
                                                       rmask, rbase, ehead);
        MMX_CHECK_CUDA(cudaGetLastError());
        if (lp.n_epochs > 0) {
            extract_long_kernel<<<(unsigned)((size_t)P * lp.n_epochs), LT, 0, stream>>>(
                words, n, lp, reinterpret_cast<const long long*>(info), rmask, rbase, ehead, jout);
            MMX_CHECK_CUDA(cudaGetLastError());
        }
        const size_t nw = (size_t)(n + 2) / 2;
        const size_t smem_apply = ((nw + 3) & ~size_t(3)) * 4 + (pl.lists_in_smem ? (size_t)n * 2 : 0);
        uint16_t* glist = pl.lists_in_smem ? nullptr : reinterpret_cast<uint16_t*>(ws + pl.off_l - 256);
        apply_long_kernel<<<pl.apply_grid, AT, smem_apply, stream>>>(jout, n, P, reinterpret_cast<const long long*>(info),
                                                                     pi, glist);
        MMX_CHECK_CUDA(cudaGetLastError());
        if (pl.rows_in_smem) {
            const size_t smem16 = (size_t)((n + 7) & ~7) * 4;
            compose_block16_kernel<<<pl.n_blocks, 1024, smem16, stream>>>(pi, n, P, pl.pb, nullptr, comp, nullptr);
            compose_scan16_kernel<<<1, 1024, smem16, stream>>>(comp, n, pl.n_blocks, prefix);
            compose_block16_kernel<<<pl.n_blocks, 1024, smem16, stream>>>(pi, n, P, pl.pb, prefix, nullptr, perm);
        } else {
            int32_t* gscan = gbuf + (size_t)pl.n_blocks * 2 * n;
            compose_block_kernel<<<pl.n_blocks, 512, 0, stream>>>(pi, n, P, pl.pb, nullptr, comp, nullptr, gbuf);
            compose_scan_kernel<<<1, 512, 0, stream>>>(comp, n, pl.n_blocks, prefix, gscan);
            compose_block_kernel<<<pl.n_blocks, 512, 0, stream>>>(pi, n, P, pl.pb, prefix, nullptr, perm, gbuf);
        }
        MMX_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    {
        const size_t smem_tab = (size_t)(RW + CT) * 4 + 3 * (size_t)((n1 + 1) & ~1) * 2;
        static const char tab_attr_tag = 0;
        if (first_use_on_device(&tab_attr_tag)) {
            MMX_CHECK_CUDA(cudaFuncSetAttribute(tables_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (RW + CT) * 4 + 3 * MAX_N * 2));
        }
        tables_kernel<<<(unsigned)pl.n_chunks, CT, smem_tab, stream>>>(words, n_words, n1, tab);
        MMX_CHECK_CUDA(cudaGetLastError());
        chain_kernel<<<1, 32, 0, stream>>>(tab, pl.n_chunks, n1, pl.steps, start_m, start_s,
                                           reinterpret_cast<long long*>(info));
        emit_kernel<<<(unsigned)pl.n_chunks, 32, 0, stream>>>(words, n_words, n1, pl.steps, start_m, start_s,
                                                              reinterpret_cast<long long*>(info), jout);
        MMX_CHECK_CUDA(cudaGetLastError());
    }
    // when the words did not suffice (info[1] = -1) the swaps are skipped and the stages below shuffle unspecified
    // but in-range index rows; the caller checks info[1] before using `perm` and retries with a longer stream
    MMX_CHECK_CUDA(cudaMemsetAsync(arr, 0, (size_t)n * P * 4, stream));
    apply_kernel<<<(P + 127) / 128, 128, 0, stream>>>(jout, n, P, reinterpret_cast<const long long*>(info), arr);
    dim3 gt((unsigned)((P + 31) / 32), (unsigned)((n + 31) / 32));
    transpose_kernel<<<gt, 256, 0, stream>>>(arr, n, P, pi);
    const size_t smem = (size_t)n * 8;
    static const char attr_set_tag = 0;
    if (first_use_on_device(&attr_set_tag)) {
        MMX_CHECK_CUDA(cudaFuncSetAttribute(compose_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_N * 8));
        MMX_CHECK_CUDA(cudaFuncSetAttribute(compose_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_N * 8));
    }
    compose_block_kernel<<<pl.n_blocks, 512, smem, stream>>>(pi, n, P, pl.pb, nullptr, comp, nullptr, nullptr);
    compose_scan_kernel<<<1, 512, smem, stream>>>(comp, n, pl.n_blocks, prefix, nullptr);
    compose_block_kernel<<<pl.n_blocks, 512, smem, stream>>>(pi, n, P, pl.pb, prefix, nullptr, perm, nullptr);
    MMX_CHECK_CUDA(cudaGetLastError());
    return 0;
}

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
constexpr int LW_MAX = 65536;       // words of one mask epoch's window (uint16 values + one flag byte each: 192 KB)

struct Plan {
    long long n_chunks, steps;
    int n_blocks;
    bool long_n;
    size_t off_tab, off_m, off_s, off_j, off_arr, off_pi, off_c, off_d, off_g, total;
};

static Plan make_plan(int n, int P, long long n_words) {
    Plan pl{};
    auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
    pl.n_chunks = (n_words + L - 1) / L;
    pl.steps = (long long)P * (n - 1);
    pl.n_blocks = (P + PB - 1) / PB;
    pl.long_n = n > MAX_N;
    size_t o = 256;
    pl.off_tab = o; o += up(pl.long_n ? 0 : (size_t)pl.n_chunks * (n - 1) * 2);   // uint16 accepts per (chunk, start state)
    pl.off_m = o;   o += up((size_t)(pl.n_chunks + 1) * 4);       // start state of every chunk
    pl.off_s = o;   o += up((size_t)(pl.n_chunks + 1) * 8);       // global step offset of every chunk
    pl.off_j = o;   o += up((size_t)pl.steps * 4);                // j of every step
    pl.off_arr = o; o += up((size_t)n * P * 4);                   // [i][P] working arrays of stage E
    pl.off_pi = o;  o += up((size_t)n * P * 4);                   // [P][n] pi_p
    pl.off_c = o;   o += up((size_t)pl.n_blocks * n * 4);         // block composites
    pl.off_d = o;   o += up((size_t)(pl.n_blocks + 1) * n * 4);   // prefixes of the block composites
    pl.off_g = o;   o += up(pl.long_n ? (size_t)(pl.n_blocks + 1) * 2 * n * 4 : 0);   // LONG: composition buffers in L2
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
// CTA that is parallel WITHIN a permutation (prototype and proof of exactness against numpy: dev/proto_long_shuffle.py).
// Inside a mask epoch (all steps i with the same mask M, i from i_hi down to i_lo) the k-th accepted word serves step
// i_hi - k, so word t is accepted iff  v_t <= i_hi - a(t),  v_t = word_t & M,  a(t) = accepts before t. a(t) is the
// only serial quantity; it is bracketed, L(t) <= a(t) <= U(t), and the brackets are tightened round by round:
//     surely accepted  v_t <= i_hi - U(t)        surely rejected  v_t > i_hi - L(t)
//     L'(t) = surely accepted words before t     U'(t) = words before t that are not surely rejected
// — two block-wide prefix sums per round, every round exact, 4-8 rounds until L = U up to the epoch's last accept.
// Every thread owns a contiguous segment of the epoch's window (values and one flag byte per word in shared memory).
struct LongScan {
    int baseA, baseU, totA, totU;
};
__device__ __forceinline__ LongScan long_block_scan(int a, int u, int* wA, int* wU, int tid) {
    const int lane = tid & 31, warp = tid >> 5;
    int ia = a, iu = u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int ta = __shfl_up_sync(0xffffffffu, ia, o), tu = __shfl_up_sync(0xffffffffu, iu, o);
        if (lane >= o) {
            ia += ta;
            iu += tu;
        }
    }
    __syncthreads();  // the previous use of wA / wU is over
    if (lane == 31) {
        wA[warp] = ia;
        wU[warp] = iu;
    }
    __syncthreads();
    int pa = 0, pu = 0, sa = 0, su = 0;
#pragma unroll
    for (int w = 0; w < LT / 32; ++w) {
        const int ca = wA[w], cu = wU[w];
        if (w < warp) {
            pa += ca;
            pu += cu;
        }
        sa += ca;
        su += cu;
    }
    LongScan r;
    r.baseA = pa + ia - a;
    r.baseU = pu + iu - u;
    r.totA = sa;
    r.totU = su;
    return r;
}

__global__ void __launch_bounds__(LT)
emit_long_kernel(const uint32_t* __restrict__ words, long long n_words, int n, int P, long long* __restrict__ info,
                 int32_t* __restrict__ jout) {
    extern __shared__ uint8_t lsm[];
    uint16_t* sv = reinterpret_cast<uint16_t*>(lsm);  // [LW_MAX] masked values of the window
    uint8_t* sf = lsm + (size_t)LW_MAX * 2;           // [LW_MAX] bit 0 = surely accepted, bit 1 = surely rejected
    __shared__ int wA[LT / 32], wU[LT / 32];
    __shared__ int s_done_at, s_conv;
    const int tid = threadIdx.x;
    long long pos = 0, step_base = 0;
    for (int p = 0; p < P; ++p, step_base += n - 1) {
        int i_hi = n - 1;
        while (i_hi >= 1) {
            const int b = 32 - __clz(i_hi);
            const uint32_t M = (b >= 32) ? 0xffffffffu : ((1u << b) - 1u);
            const int i_lo = max(1 << (b - 1), 1);
            const int c = i_hi - i_lo + 1;
            int W = min(LW_MAX, (int)(1.40f * (float)c) + 1024);
            bool epoch_done = false;
            while (!epoch_done) {
                if (pos + W > n_words) W = (int)(n_words - pos);
                if (W < c) {  // the stream is too short: the caller retries with more words
                    if (tid == 0) {
                        info[0] = -1;
                        info[1] = -1;
                    }
                    return;
                }
                __syncthreads();  // everybody is done with the previous window
                for (int t = tid; t < W; t += LT) {
                    const uint32_t v = words[pos + t] & M;
                    sv[t] = (uint16_t)v;
                    sf[t] = (v <= (uint32_t)i_hi) ? 0 : 2;  // a(t) >= 0: a word above i_hi is never accepted
                }
                __syncthreads();
                const int S = (W + LT - 1) / LT;
                const int t0 = min(W, tid * S), t1 = min(W, t0 + S);
                int prevA = -1, prevU = -1;
                for (int round = 0;; ++round) {
                    // ---- counts of the current flags, their block-wide exclusive prefixes
                    int a = 0, u = 0;
                    for (int t = t0; t < t1; ++t) {
                        const int f = sf[t];
                        a += f & 1;
                        u += (f >> 1) ^ 1;
                    }
                    const LongScan sc = long_block_scan(a, u, wA, wU, tid);
                    if (tid == 0) s_done_at = -1;
                    __syncthreads();
                    // ---- the thread whose segment holds the c-th sure accept: is everything before it settled?
                    if (sc.baseA < c && sc.baseA + a >= c) {
                        int ra = sc.baseA, ru = sc.baseU;
                        for (int t = t0; t < t1; ++t) {
                            const int f = sf[t];
                            ra += f & 1;
                            ru += (f >> 1) ^ 1;
                            if (ra == c) {
                                s_done_at = t;
                                s_conv = (ru == c) ? 1 : 0;  // U = L up to and including the last accept
                                break;
                            }
                        }
                    }
                    __syncthreads();
                    const int done_at = s_done_at;
                    if (done_at >= 0 && s_conv) {
                        // ---- emit: the k-th accepted word of the epoch is the j of step i_hi - k
                        int ra = sc.baseA;
                        int32_t* out = jout + step_base + (n - 1 - i_hi);
                        for (int t = t0; t < t1 && ra < c; ++t)
                            if (sf[t] & 1) out[ra++] = (int32_t)sv[t];
                        pos += done_at + 1;
                        epoch_done = true;
                        break;
                    }
                    if (sc.totA == prevA && sc.totU == prevU) {
                        // no flag changed in the last round and the c-th accept is not settled: the window is too short
                        if (W >= LW_MAX || pos + W >= n_words) {
                            if (tid == 0) {
                                info[0] = -1;
                                info[1] = (W >= LW_MAX) ? -3 : -1;
                            }
                            return;
                        }
                        W = min(LW_MAX, 2 * W);
                        break;  // restage a longer window of the same epoch
                    }
                    prevA = sc.totA;
                    prevU = sc.totU;
                    // ---- tighten: new flags from the brackets L(t) = ra, U(t) = ru of the OLD flags
                    {
                        int ra = sc.baseA, ru = sc.baseU;
                        for (int t = t0; t < t1; ++t) {
                            const int f = sf[t];
                            const int v = sv[t];
                            const int nf = (v <= i_hi - ru ? 1 : 0) | (v > i_hi - ra ? 2 : 0);
                            ra += f & 1;
                            ru += (f >> 1) ^ 1;
                            sf[t] = (uint8_t)nf;
                        }
                    }
                    // each thread rewrote only its own segment; the next round's scan synchronises before any
                    // cross-thread read
                }
            }
            i_hi = i_lo - 1;
        }
    }
    if (tid == 0) {
        info[0] = 0;
        info[1] = pos;  // exactly this many words were consumed
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
compose_block_kernel(const int32_t* __restrict__ pi, int n, int P, const int32_t* __restrict__ start,
                     int32_t* __restrict__ comp_out, int32_t* __restrict__ rows_out, int32_t* gbuf) {
    extern __shared__ int32_t sm[];
    const int b = blockIdx.x;
    // long arrays (n > MAX_N) keep the two working rows in L2 (gbuf: 2 n words per block) instead of shared memory
    int32_t* cur = gbuf != nullptr ? gbuf + (size_t)b * 2 * n : sm;
    int32_t* nxt = cur + n;
    for (int k = threadIdx.x; k < n; k += blockDim.x) cur[k] = start != nullptr ? start[(long long)b * n + k] : k;
    __syncthreads();
    const int p_end = min(P, (b + 1) * PB);
    for (int p = b * PB; p < p_end; ++p) {
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
        // n > MAX_N: one CTA, parallel within a permutation (bracketing prefix sums per mask epoch)
        const size_t smem_long = (size_t)LW_MAX * 3;
        static const char long_attr_tag = 0;
        if (first_use_on_device(&long_attr_tag))
            MMX_CHECK_CUDA(cudaFuncSetAttribute(emit_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)smem_long));
        emit_long_kernel<<<1, LT, smem_long, stream>>>(words, n_words, n, P, reinterpret_cast<long long*>(info), jout);
        MMX_CHECK_CUDA(cudaGetLastError());
    } else {
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
    const size_t smem = pl.long_n ? 0 : (size_t)n * 8;
    static const char attr_set_tag = 0;
    if (first_use_on_device(&attr_set_tag)) {
        MMX_CHECK_CUDA(cudaFuncSetAttribute(compose_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_N * 8));
        MMX_CHECK_CUDA(cudaFuncSetAttribute(compose_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_N * 8));
    }
    int32_t* gscan = pl.long_n ? gbuf + (size_t)pl.n_blocks * 2 * n : nullptr;
    compose_block_kernel<<<pl.n_blocks, 512, smem, stream>>>(pi, n, P, nullptr, comp, nullptr, gbuf);
    compose_scan_kernel<<<1, 512, smem, stream>>>(comp, n, pl.n_blocks, prefix, gscan);
    compose_block_kernel<<<pl.n_blocks, 512, smem, stream>>>(pi, n, P, prefix, nullptr, perm, gbuf);
    MMX_CHECK_CUDA(cudaGetLastError());
    return 0;
}

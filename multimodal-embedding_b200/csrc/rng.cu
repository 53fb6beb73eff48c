// K6 — numpy's legacy MT19937 stream, continued on the device.
//
// The reference draws every resample index from numpy's global RandomState: np.random.choice(n, n) per bootstrap
// iteration (/root/reference/main.py:588) and np.random.choice(data, n) x 10 000 per confidence interval
// (main.py:269) — ~9.5e7 draws per model, ~1.1 s of host time that nothing on the GPU can hide once scoring takes
// milliseconds. Both calls reduce to RandomState.randint(0, n, size), i.e. numpy's legacy
// "masked rejection" on 32-bit MT19937 outputs (numpy/random/src/distributions/distributions.c,
// random_bounded_uint64_fill -> buffered_bounded_masked_uint32):
//       do { v = next_uint32() & mask; } while (v > rng);        mask = 2^ceil(log2(rng+1)) - 1
// This kernel takes the generator state (624-word key + position), produces exactly the values numpy would produce
// next, in order, and leaves the state where numpy would leave it, so host and device draws can be interleaved
// and np.random.set_state() keeps the reference's stream consumption bit-identical.
//
// MT19937 is a serial recurrence over 624-word blocks, but inside a block x[k] depends on x[k-227] and old words
// only: three parallel phases (227 + 227 + 170 words) regenerate a block. One small CTA (128 threads, <= 32
// registers, 5 KB of shared memory — it fits beside a resident scoring CTA, so it overlaps with K1/K2 on a side
// stream): regenerate, temper, test, and compact the accepted values in order with warp ballots.
#include "../../include/mmx.h"
#include "mmx_host.cuh"

namespace mmx {
namespace rng {

constexpr int N = 624, M = 397;
constexpr uint32_t MATRIX_A = 0x9908b0dfu, UPPER = 0x80000000u, LOWER = 0x7fffffffu;
constexpr int THREADS = 128;
constexpr int WARPS = THREADS / 32;
constexpr int ROUNDS = (N + 31) / 32;                  // 20 rounds of 32 consecutive words
constexpr int RPW = (ROUNDS + WARPS - 1) / WARPS;      // rounds per warp

__device__ __forceinline__ uint32_t twist(uint32_t a, uint32_t b) {
    const uint32_t y = (a & UPPER) | (b & LOWER);
    return (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
}
__device__ __forceinline__ uint32_t temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

__global__ void __launch_bounds__(THREADS, 16)
mt19937_randint_kernel(uint32_t* __restrict__ state, uint32_t rng, uint32_t mask, long long count,
                       int32_t* __restrict__ out, const unsigned long long* __restrict__ seg_cnt, int n_seg,
                       long long win_lo, long long win_hi) {   // only values [win_lo, win_hi) of the request are stored
    __shared__ uint32_t key[2][N];
    __shared__ int round_cnt[32];
    __shared__ int new_pos_s;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    for (int i = t; i < N; i += THREADS) key[0][i] = state[i];
    if (t < 32) round_cnt[t] = 0;
    int pos = (int)state[N];
    int cur = 0;
    long long produced = 0;
    // tail of the parallel path: the segments already emitted sum(seg_cnt) values; nothing to do unless the
    // budgeted stream ran out early (then `state` is the end of the last segment and the loop simply continues)
    for (int j = 0; j < n_seg; ++j) produced += (long long)seg_cnt[j];
    if (produced >= count) return;
    __syncthreads();
    while (produced < count) {
        if (pos >= N) {
            // ---- regenerate the block: old = key[cur], new = key[cur^1]
            const uint32_t* o = key[cur];
            uint32_t* n = key[cur ^ 1];
#pragma unroll
            for (int i = t; i < 227; i += THREADS) n[i] = o[i + M] ^ twist(o[i], o[i + 1]);
            __syncthreads();
#pragma unroll
            for (int i = t; i < 227; i += THREADS) n[i + 227] = n[i] ^ twist(o[i + 227], o[i + 228]);
            __syncthreads();
#pragma unroll
            for (int i = t; i < 169; i += THREADS) n[i + 454] = n[i + 227] ^ twist(o[i + 454], o[i + 455]);
            if (t == THREADS - 1) n[623] = n[396] ^ twist(o[623], n[0]);
            __syncthreads();
            cur ^= 1;
            pos = 0;
        }
        // ---- temper + masked rejection on words [pos, 624): warp w owns rounds w, w+4, ... (32 words each)
        const uint32_t* k = key[cur];
        uint32_t v[RPW], m[RPW];
#pragma unroll
        for (int j = 0; j < RPW; ++j) {
            const int e = (warp + j * WARPS) * 32 + lane;
            bool ok = false;
            v[j] = 0;
            if (e < N && e >= pos) {
                v[j] = temper(k[e]) & mask;
                ok = v[j] <= rng;
            }
            m[j] = __ballot_sync(0xffffffffu, ok);
            if (lane == 0) round_cnt[warp + j * WARPS] = __popc(m[j]);
        }
        __syncthreads();
        // exclusive prefix over the 20 round counts (every warp computes it redundantly with shuffles)
        const int c = round_cnt[lane];
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        const long long need = count - produced;
        const bool last_block = (long long)total >= need;
#pragma unroll
        for (int j = 0; j < RPW; ++j) {
            const int r = warp + j * WARPS;
            const int base = __shfl_sync(0xffffffffu, incl - c, r & 31);
            if (r < ROUNDS && ((m[j] >> lane) & 1u)) {
                const int off = base + __popc(m[j] & lt);
                if ((long long)off < need) {
                    const long long idx = produced + off;
                    if (out != nullptr && idx >= win_lo && idx < win_hi) out[idx] = (int32_t)v[j];
                    if (last_block && (long long)off == need - 1) new_pos_s = r * 32 + lane + 1;
                }
            }
        }
        __syncthreads();
        if (last_block) {
            pos = new_pos_s;
            produced = count;
        } else {
            pos = N;
            produced += total;
        }
    }
    for (int i = t; i < N; i += THREADS) state[i] = key[cur][i];
    if (t == 0) state[N] = (uint32_t)pos;
}

}  // namespace rng
}  // namespace mmx

using namespace mmx;

extern "C" int mmx_mt19937_randint(uint32_t* state, uint32_t high, int64_t count, int32_t* out, cudaStream_t stream) {
    MMX_REQUIRE(state != nullptr && (out != nullptr || count == 0), "NULL argument");
    MMX_REQUIRE(high >= 1 && high <= 0x7fffffffu, "high must be in [1, 2^31)");
    MMX_REQUIRE(count >= 0, "negative count");
    if (count == 0) return 0;
    const uint32_t r = high - 1;  // inclusive maximum
    if (r == 0) {                 // numpy draws nothing for a single-value range
        MMX_CHECK_CUDA(cudaMemsetAsync(out, 0, (size_t)count * sizeof(int32_t), stream));
        return 0;
    }
    uint32_t mask = r;
    mask |= mask >> 1;
    mask |= mask >> 2;
    mask |= mask >> 4;
    mask |= mask >> 8;
    mask |= mask >> 16;
    rng::mt19937_randint_kernel<<<1, rng::THREADS, 0, stream>>>(state, r, mask, (long long)count, out, nullptr, 0, 0,
                                                                (long long)count);
    MMX_CHECK_CUDA(cudaGetLastError());
    return 0;
}


// ================================================================================================================
// K6p — the same stream, produced by every SM at once (MT19937 jump-ahead).
//
// The stream is cut into segments of `seg_blocks` numpy blocks (624 raw words each); segment k starts from the block
// W_k = F^(k * seg_blocks * 624)(key), obtained WITHOUT stepping: MT19937 is a linear recurrence over GF(2) with a
// degree-19937 characteristic polynomial phi, so with g_k(x) = x^(k*seg_words) mod phi(x) (host table,
// multimodal-embedding_b200/mtjump.py) every raw word obeys x[n + k*seg_words] = XOR_{i: g_k[i]=1} x[n + i]:
// W_k is a GF(2) convolution of the first 33 blocks of the stream (expanded once by one CTA) with the bits of g_k.
// Then every CTA regenerates its own segment. Two passes keep numpy's order under rejection sampling: pass A counts
// the accepted values per segment, pass B recomputes them and writes at the prefix-summed offsets; the CTA that
// emits the last requested value stores the generator state (block key + position) exactly where numpy would be.
// The segment budget carries 2 % + 64 blocks of slack over the expected rejection rate; should a request still
// outrun it, the sequential kernel above finishes the remainder in the same stream order (no host round trip).
// ================================================================================================================
namespace mmx {
namespace rng {

constexpr int PTHREADS = 256;
constexpr int PWARPS = PTHREADS / 32;
constexpr int XBLOCKS = 33;  // 33 * 624 = 20592 >= 19937 + 623 raw words feed the convolution
constexpr int XPAD = 768;    // zero words behind the stream: the target of padded bit positions (index XBLOCKS * 624)

struct ParParams {
    const uint32_t* state;        // [625] key + pos (read only)
    const uint32_t* polys;        // [n_poly][624] jump polynomials, polys[k-1] = x^(k*seg_words) mod phi
    const uint16_t* bitpos;       // nullable [n_poly][bitpos_ld]: positions of the set bits of every polynomial, padded
    const int32_t* bitcnt;        //   with a position whose x word is zero (see mt_par_expand_kernel); bitcnt[k] = count
    int bitpos_ld;
    uint32_t* xstream;            // [XBLOCKS * 624] raw words that follow the current key (key itself = block 0)
    uint32_t* seg_states;         // [n_seg][624]
    unsigned long long* seg_cnt;  // [n_seg]
    uint32_t* state_out;          // [625] key + pos where the request ends
    int32_t* out;                 // nullable (only advance the state); indexed by the value's position in this launch
    long long count;
    long long win_lo, win_hi;     // only values [win_lo, win_hi) of this launch are stored (a rank's block of the stream)
    int n_seg, seg_blocks;
    uint32_t rng, mask;
};

// n = the block after o (both in shared memory), 256 threads
__device__ __forceinline__ void gen_block256(const uint32_t* o, uint32_t* n, int t) {
    if (t < 227) n[t] = o[t + M] ^ twist(o[t], o[t + 1]);
    __syncthreads();
    if (t < 227) n[t + 227] = n[t] ^ twist(o[t + 227], o[t + 228]);
    __syncthreads();
    if (t < 169) n[t + 454] = n[t + 227] ^ twist(o[t + 454], o[t + 455]);
    if (t == 255) n[623] = n[396] ^ twist(o[623], n[0]);
    __syncthreads();
}

__global__ void __launch_bounds__(PTHREADS)
mt_par_expand_kernel(const ParParams p) {
    __shared__ uint32_t key[2][N];
    const int t = threadIdx.x;
    for (int i = t; i < N; i += PTHREADS) {
        key[0][i] = p.state[i];
        p.xstream[i] = key[0][i];
    }
    __syncthreads();
    int cur = 0;
    for (int b = 1; b < XBLOCKS; ++b) {
        gen_block256(key[cur], key[cur ^ 1], t);
        cur ^= 1;
        for (int i = t; i < N; i += PTHREADS) p.xstream[b * N + i] = key[cur][i];
    }
}

__global__ void __launch_bounds__(PTHREADS)
mt_par_count_kernel(const ParParams p) {
    extern __shared__ uint32_t sm[];
    uint32_t* x = sm;                   // [XBLOCKS * N] raw stream from the base key (segments k > 0)
    uint32_t* key0 = sm + XBLOCKS * N + XPAD;  // two block buffers for the segment itself
    uint32_t* key1 = key0 + N;
    __shared__ unsigned long long wcnt[PWARPS];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int k = blockIdx.x;
    const int pos0 = (int)p.state[N];
    if (k == 0) {
        for (int i = t; i < N; i += PTHREADS) key0[i] = p.state[i];
    } else {
        for (int i = t; i < XBLOCKS * N; i += PTHREADS) x[i] = p.xstream[i];
        for (int i = t; i < XPAD; i += PTHREADS) x[XBLOCKS * N + i] = 0u;
        __syncthreads();
        uint32_t a0 = 0, a1 = 0, a2 = 0;
        const bool third = t + 512 < N;
        const uint32_t* xt = x + t;
        if (p.bitpos != nullptr) {
            // precomputed set-bit positions: four independent gathers in flight, no bit scanning in the loop
            const uint16_t* bp = p.bitpos + (size_t)(k - 1) * p.bitpos_ld;
            const int nb = (__ldg(p.bitcnt + (k - 1)) + 3) & ~3;   // the padding positions index zero words
            for (int j = 0; j < nb; j += 4) {
                const uint2 q = __ldg(reinterpret_cast<const uint2*>(bp + j));
                const int i0 = q.x & 0xffff, i1 = q.x >> 16, i2 = q.y & 0xffff, i3 = q.y >> 16;
                a0 ^= xt[i0] ^ xt[i1] ^ xt[i2] ^ xt[i3];
                a1 ^= xt[i0 + 256] ^ xt[i1 + 256] ^ xt[i2 + 256] ^ xt[i3 + 256];
                if (third) a2 ^= xt[i0 + 512] ^ xt[i1 + 512] ^ xt[i2 + 512] ^ xt[i3 + 512];
            }
        } else {
            // no position table registered: scan the bits of the polynomial word by word
            const uint32_t* g = p.polys + (size_t)(k - 1) * N;
            uint32_t gnext = __ldg(g);
            for (int w = 0; w < N; ++w) {
                uint32_t gw = gnext;
                if (w + 1 < N) gnext = __ldg(g + w + 1);
                const uint32_t* xw = xt + w * 32;
                while (gw) {
                    const int i = __ffs(gw) - 1;
                    gw &= gw - 1;
                    a0 ^= xw[i];
                    a1 ^= xw[i + 256];
                    if (third) a2 ^= xw[i + 512];
                }
            }
        }
        key0[t] = a0;
        key0[t + 256] = a1;
        if (third) key0[t + 512] = a2;
    }
    __syncthreads();
    for (int i = t; i < N; i += PTHREADS) p.seg_states[(size_t)k * N + i] = key0[i];
    // ---- count the accepted values of this segment (segment 0: plus the unread tail of the current block)
    unsigned long long cnt = 0;
    if (k == 0) {
        for (int e = t; e < N; e += PTHREADS)
            if (e >= pos0 && (temper(key0[e]) & p.mask) <= p.rng) ++cnt;
    }
    uint32_t* o = key0;
    uint32_t* n = key1;
    for (int b = 0; b < p.seg_blocks; ++b) {
        gen_block256(o, n, t);
        for (int e = t; e < N; e += PTHREADS)
            if ((temper(n[e]) & p.mask) <= p.rng) ++cnt;
        uint32_t* tmp = o;
        o = n;
        n = tmp;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, s);
    if (lane == 0) wcnt[warp] = cnt;
    __syncthreads();
    if (t == 0) {
        unsigned long long tot = 0;
        for (int w = 0; w < PWARPS; ++w) tot += wcnt[w];
        p.seg_cnt[k] = tot;
    }
}

// pass B: ordered compaction of the accepted words of the segment at the prefix-summed offset
__global__ void __launch_bounds__(PTHREADS)
mt_par_write_kernel(const ParParams p) {
    __shared__ uint32_t key[2][N];
    __shared__ int round_cnt[32];
    __shared__ int stop_pos_s;
    __shared__ unsigned long long base_s, total_s;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const int k = blockIdx.x;
    if (t < 32) {
        round_cnt[t] = 0;
        unsigned long long b = 0, tot = 0;
        for (int j = lane; j < p.n_seg; j += 32) {
            const unsigned long long c = p.seg_cnt[j];
            if (j < k) b += c;
            tot += c;
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            b += __shfl_xor_sync(0xffffffffu, b, s);
            tot += __shfl_xor_sync(0xffffffffu, tot, s);
        }
        if (lane == 0) {
            base_s = b;
            total_s = tot;
        }
    }
    for (int i = t; i < N; i += PTHREADS) key[0][i] = p.seg_states[(size_t)k * N + i];
    __syncthreads();
    const long long base = (long long)base_s;
    const bool shortfall = (long long)total_s < p.count;
    if (base >= p.count) return;  // the request ends in an earlier segment
    {
        // nothing to store from this segment (no output, or outside the window) and the request does not end in it:
        // the count pass already told everybody where the later segments start
        const long long seg_end = base + (long long)p.seg_cnt[k];
        const bool terminal = seg_end >= p.count || (shortfall && k == p.n_seg - 1);
        if (!terminal && (p.out == nullptr || seg_end <= p.win_lo || base >= p.win_hi)) return;
    }
    long long produced = base;    // global index of the next value this CTA emits
    int cur = 0;
    int pos = (k == 0) ? (int)p.state[N] : N;
    int blocks_left = p.seg_blocks;
    bool finished = false;
    while (!finished) {
        if (pos >= N) {
            if (blocks_left == 0) break;
            gen_block256(key[cur], key[cur ^ 1], t);
            cur ^= 1;
            pos = 0;
            --blocks_left;
        }
        const uint32_t* kk = key[cur];
        constexpr int RPW8 = (ROUNDS + PWARPS - 1) / PWARPS;  // 3
        uint32_t v[RPW8], m[RPW8];
#pragma unroll
        for (int j = 0; j < RPW8; ++j) {
            const int r = warp + j * PWARPS;
            const int e = r * 32 + lane;
            bool ok = false;
            v[j] = 0;
            if (r < ROUNDS && e < N && e >= pos) {
                v[j] = temper(kk[e]) & p.mask;
                ok = v[j] <= p.rng;
            }
            m[j] = __ballot_sync(0xffffffffu, ok);
            if (lane == 0 && r < ROUNDS) round_cnt[r] = __popc(m[j]);
        }
        __syncthreads();
        const int c = round_cnt[lane];
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        const long long need = p.count - produced;
        const bool last_block = (long long)total >= need;
#pragma unroll
        for (int j = 0; j < RPW8; ++j) {
            const int r = warp + j * PWARPS;
            const int bs = __shfl_sync(0xffffffffu, incl - c, r & 31);
            if (r < ROUNDS && ((m[j] >> lane) & 1u)) {
                const int off = bs + __popc(m[j] & lt);
                if ((long long)off < need) {
                    const long long idx = produced + off;
                    if (p.out != nullptr && idx >= p.win_lo && idx < p.win_hi) p.out[idx] = (int32_t)v[j];
                    if (last_block && (long long)off == need - 1) stop_pos_s = r * 32 + lane + 1;
                }
            }
        }
        __syncthreads();
        if (last_block) {
            pos = stop_pos_s;
            finished = true;
        } else {
            pos = N;
            produced += total;
        }
    }
    if (finished || (shortfall && k == p.n_seg - 1)) {
        // this CTA emitted the last requested value (or the budgeted stream ended): numpy's state is here
        for (int i = t; i < N; i += PTHREADS) p.state_out[i] = key[cur][i];
        if (t == 0) p.state_out[N] = (uint32_t)pos;
    }
}

static uint32_t mask_of(uint32_t r) {
    uint32_t mask = r;
    mask |= mask >> 1;
    mask |= mask >> 2;
    mask |= mask >> 4;
    mask |= mask >> 8;
    mask |= mask >> 16;
    return mask;
}

// raw words a request for `count` accepted values is budgeted with (mmx_mt19937_par_set_slack: test hook that
// lets the parity tests starve the budget and exercise the sequential finish)
static double g_slack_rel = 1.02, g_slack_blocks = 64.0;
static double budget_words(uint32_t r, double count) {
    const double p = ((double)r + 1.0) / ((double)mask_of(r) + 1.0);
    return count / p * g_slack_rel + g_slack_blocks * N;
}

}  // namespace rng
}  // namespace mmx

extern "C" int mmx_mt19937_par_set_slack(double rel, double blocks) {
    MMX_REQUIRE(rel > 0.0 && blocks >= 0.0, "bad slack");
    mmx::rng::g_slack_rel = rel;
    mmx::rng::g_slack_blocks = blocks;
    return 0;
}

extern "C" int mmx_mt19937_par_scratch_bytes(int n_poly, size_t* bytes) {
    MMX_REQUIRE(bytes != nullptr && n_poly >= 1, "bad argument");
    using namespace mmx::rng;
    *bytes = (size_t)XBLOCKS * N * 4 + (size_t)(n_poly + 1) * N * 4 + (size_t)(n_poly + 1) * 8 + 625 * 4 + 1024;
    return 0;
}

// Optional accelerator of the jump convolution: the set-bit positions of every polynomial of the table the caller
// passes as `polys` (uint16 [n_poly][ld], rows padded to a multiple of 4 with the value 33 * 624 = 20592) and their
// counts. One registration per table (device pointer); results do not depend on it.
struct BitPosReg {
    const uint32_t* polys;
    const uint16_t* bitpos;
    const int32_t* bitcnt;
    int ld;
};
static BitPosReg g_bitpos_reg[16];  // one entry per (device, table): keyed by the table's device pointer
static int g_n_bitpos_reg = 0;
extern "C" int mmx_mt19937_par_set_bit_positions(const uint32_t* polys, const uint16_t* bitpos, const int32_t* bitcnt,
                                                 int ld) {
    MMX_REQUIRE(polys != nullptr, "polys == NULL");
    MMX_REQUIRE((bitpos == nullptr) == (bitcnt == nullptr) && (bitpos == nullptr || (ld >= 4 && ld % 4 == 0)),
                "bad bit-position table");
    int slot = -1;
    for (int i = 0; i < g_n_bitpos_reg; ++i)
        if (g_bitpos_reg[i].polys == polys) slot = i;
    if (slot < 0) {
        MMX_REQUIRE(g_n_bitpos_reg < 16, "too many registered jump tables");
        slot = g_n_bitpos_reg++;
    }
    g_bitpos_reg[slot] = BitPosReg{polys, bitpos, bitcnt, ld};
    return 0;
}

// rng_incl = inclusive maximum of the masked-rejection draw (0xffffffff: every raw word is a value)
// out holds the values [win_begin, win_begin + win_count) of the request (the generator advances by all `count`)
static int par_generate(uint32_t* state, uint32_t rng_incl, int64_t count, int32_t* out, int64_t win_begin,
                        int64_t win_count, const uint32_t* polys, int n_poly, int seg_blocks, void* scratch,
                        size_t scratch_bytes, cudaStream_t stream) {
    using namespace mmx::rng;
    size_t need = 0;
    mmx_mt19937_par_scratch_bytes(n_poly, &need);
    MMX_REQUIRE(scratch_bytes >= need, "scratch too small: need %zu bytes", need);
    ParParams p;
    uintptr_t ws = (reinterpret_cast<uintptr_t>(scratch) + 255) & ~uintptr_t(255);
    p.xstream = reinterpret_cast<uint32_t*>(ws);
    ws += (size_t)XBLOCKS * N * 4;
    p.seg_states = reinterpret_cast<uint32_t*>(ws);
    ws += (size_t)(n_poly + 1) * N * 4;
    p.seg_cnt = reinterpret_cast<unsigned long long*>(ws);
    ws += (size_t)(n_poly + 1) * 8;
    p.state_out = reinterpret_cast<uint32_t*>(ws);
    p.state = state;
    p.polys = polys;
    p.seg_blocks = seg_blocks;
    p.bitpos = nullptr;
    p.bitcnt = nullptr;
    p.bitpos_ld = 0;
    for (int i = 0; i < g_n_bitpos_reg; ++i)
        if (g_bitpos_reg[i].polys == polys && g_bitpos_reg[i].bitpos != nullptr) {
            p.bitpos = g_bitpos_reg[i].bitpos;
            p.bitcnt = g_bitpos_reg[i].bitcnt;
            p.bitpos_ld = g_bitpos_reg[i].ld;
        }
    p.rng = rng_incl;
    p.mask = mask_of(p.rng);
    const size_t smem = ((size_t)(XBLOCKS + 2) * N + XPAD) * 4;
    MMX_CHECK_CUDA(cudaFuncSetAttribute(mt_par_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const double per_seg = (double)seg_blocks * N;
    // values one launch (n_poly + 1 segments) is budgeted to deliver; larger requests are chained launches, each
    // continuing from the state the previous one left (stream order)
    const double cap_words = per_seg * (n_poly + 1);
    MMX_REQUIRE(cap_words > 128.0 * N, "jump table too small");
    const double accept = ((double)p.rng + 1.0) / ((double)p.mask + 1.0);
    const int64_t cap_count = (int64_t)((cap_words - 64.0 * N) / 1.02 * accept);  // default slack: chunking is fixed
    int64_t done = 0;
    while (done < count) {
        const int64_t c = (count - done < cap_count) ? (count - done) : cap_count;
        int n_seg = (int)(budget_words(p.rng, (double)c) / per_seg) + 1;
        if (n_seg > n_poly + 1) n_seg = n_poly + 1;
        p.win_lo = std::max<int64_t>(win_begin - done, 0);
        p.win_hi = std::min<int64_t>(win_begin + win_count - done, c);
        // p.out[i] = value i of this launch: value done + i of the request = out[done + i - win_begin]
        p.out = (out != nullptr && p.win_hi > p.win_lo) ? out + (done - win_begin) : nullptr;
        p.count = c;
        p.n_seg = n_seg;
        mt_par_expand_kernel<<<1, PTHREADS, 0, stream>>>(p);
        mt_par_count_kernel<<<n_seg, PTHREADS, smem, stream>>>(p);
        mt_par_write_kernel<<<n_seg, PTHREADS, 0, stream>>>(p);
        MMX_CHECK_CUDA(cudaGetLastError());
        MMX_CHECK_CUDA(cudaMemcpyAsync(state, p.state_out, 625 * 4, cudaMemcpyDeviceToDevice, stream));
        mt19937_randint_kernel<<<1, THREADS, 0, stream>>>(state, p.rng, p.mask, (long long)c, p.out, p.seg_cnt, n_seg,
                                                          p.win_lo, p.win_hi);
        MMX_CHECK_CUDA(cudaGetLastError());
        done += c;
    }
    return 0;
}

extern "C" int mmx_mt19937_randint_par(uint32_t* state, uint32_t high, int64_t count, int32_t* out,
                                       const uint32_t* polys, int n_poly, int seg_blocks, void* scratch,
                                       size_t scratch_bytes, cudaStream_t stream) {
    MMX_REQUIRE(state != nullptr && (out != nullptr || count == 0) && polys != nullptr && scratch != nullptr,
                "NULL argument");
    MMX_REQUIRE(high >= 1 && high <= 0x7fffffffu, "high must be in [1, 2^31)");
    MMX_REQUIRE(count >= 0 && n_poly >= 1 && seg_blocks >= 1, "bad argument");
    if (count == 0 || high == 1) return mmx_mt19937_randint(state, high, count, out, stream);
    return par_generate(state, high - 1, count, out, 0, count, polys, n_poly, seg_blocks, scratch, scratch_bytes, stream);
}

extern "C" int mmx_mt19937_randint_par_window(uint32_t* state, uint32_t high, int64_t count, int32_t* out,
                                              int64_t win_begin, int64_t win_count, const uint32_t* polys, int n_poly,
                                              int seg_blocks, void* scratch, size_t scratch_bytes, cudaStream_t stream) {
    MMX_REQUIRE(state != nullptr && polys != nullptr && scratch != nullptr, "NULL argument");
    MMX_REQUIRE(high >= 1 && high <= 0x7fffffffu, "high must be in [1, 2^31)");
    MMX_REQUIRE(count >= 0 && n_poly >= 1 && seg_blocks >= 1, "bad argument");
    MMX_REQUIRE(win_begin >= 0 && win_count >= 0 && win_begin + win_count <= count, "window outside the request");
    MMX_REQUIRE(out != nullptr || win_count == 0, "out == NULL with a non-empty window");
    if (count == 0) return 0;
    if (high == 1) {  // numpy draws nothing for a single-value range
        if (win_count > 0) MMX_CHECK_CUDA(cudaMemsetAsync(out, 0, (size_t)win_count * sizeof(int32_t), stream));
        return 0;
    }
    return par_generate(state, high - 1, count, win_count > 0 ? out : nullptr, win_begin, win_count, polys, n_poly,
                        seg_blocks, scratch, scratch_bytes, stream);
}

extern "C" int mmx_mt19937_raw_par(uint32_t* state, int64_t count, uint32_t* out, const uint32_t* polys, int n_poly,
                                   int seg_blocks, void* scratch, size_t scratch_bytes, cudaStream_t stream) {
    MMX_REQUIRE(state != nullptr && polys != nullptr && scratch != nullptr, "NULL argument");
    MMX_REQUIRE(count >= 0 && n_poly >= 1 && seg_blocks >= 1, "bad argument");
    if (count == 0) return 0;
    return par_generate(state, 0xffffffffu, count, reinterpret_cast<int32_t*>(out), 0, count, polys, n_poly, seg_blocks,
                        scratch, scratch_bytes, stream);
}

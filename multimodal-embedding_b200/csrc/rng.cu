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
                       int32_t* __restrict__ out) {
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
                    out[produced + off] = (int32_t)v[j];
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
    rng::mt19937_randint_kernel<<<1, rng::THREADS, 0, stream>>>(state, r, mask, (long long)count, out);
    MMX_CHECK_CUDA(cudaGetLastError());
    return 0;
}

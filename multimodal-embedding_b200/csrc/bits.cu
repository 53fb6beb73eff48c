// Shard merge of the above-GT bitmask: every gallery/document shard produces the bit columns of ITS columns
// (bit j of word w = local column 32 w + j); ONE all-gather moves them, and this kernel shifts each shard's words to
// their global bit offset (a shard may start in the middle of a 32-column word), ORs them together and counts the set
// bits of every row — the strict-greater count the reference's `gt in torch.topk(row, k)` tests need
// (/root/reference/main.py:331-378) — so no second collective is needed for the counts. Integer/byte streaming work,
// HBM-bound: reads n_shards * nq * wmax words, writes nq * wpr words.
#include "../../include/mmx.h"
#include "mmx_host.cuh"

namespace mmx {
namespace bits {

constexpr int MAX_SHARDS = 64;
struct Bounds {
    long long begin[MAX_SHARDS + 1];  // begin[s] .. begin[s + 1] = global columns of shard s
};

// one warp per query row
__global__ void __launch_bounds__(256)
place_bit_columns_kernel(const uint32_t* __restrict__ gathered, int n_shards, long long nq, long long wmax,
                         const __grid_constant__ Bounds bd, long long wpr, uint32_t* __restrict__ out,
                         int32_t* __restrict__ cnt) {
    const int lane = threadIdx.x & 31;
    const long long q = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    int total = 0;
    for (long long w = lane; w < wpr; w += 32) {
        const long long c0 = w << 5;  // first global column of this word
        uint32_t word = 0;
        for (int s = 0; s < n_shards; ++s) {
            const long long b = bd.begin[s], e = bd.begin[s + 1];
            if (e <= c0 || b >= c0 + 32 || e == b) continue;
            const uint32_t* src = gathered + ((long long)s * nq + q) * wmax;
            const long long width = e - b;
            const long long lo = c0 - b;  // local column of this word's bit 0 (negative: the shard starts inside the word)
            uint32_t v;
            if (lo >= 0) {
                const long long lw = lo >> 5;
                const int sh = (int)(lo & 31);
                const uint32_t a = src[lw];
                const uint32_t nx = (sh != 0 && ((lw + 1) << 5) < width) ? src[lw + 1] : 0u;
                v = sh ? ((a >> sh) | (nx << (32 - sh))) : a;
                const long long valid = width - lo;  // local columns left from bit 0 of v on
                if (valid < 32) v &= (1u << valid) - 1u;
            } else {
                const int sh = (int)(-lo);  // 1..31
                v = src[0] << sh;
                if (width < 32 - sh) v &= ((1u << width) - 1u) << sh;
            }
            word |= v;
        }
        out[q * wpr + w] = word;
        total += __popc(word);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    if (lane == 0 && cnt != nullptr) cnt[q] = total;
}

}  // namespace bits
}  // namespace mmx

using namespace mmx;
using namespace mmx::bits;

extern "C" int mmx_place_bit_columns(const uint32_t* gathered, int n_shards, int64_t nq, int64_t wmax,
                                     const int64_t* shard_begin_h, uint32_t* above_bits, int64_t words_per_row,
                                     int32_t* cnt_above, cudaStream_t stream) {
    MMX_REQUIRE(gathered && shard_begin_h && above_bits, "NULL argument");
    MMX_REQUIRE(n_shards >= 1 && n_shards <= MAX_SHARDS, "1 <= n_shards <= %d", MAX_SHARDS);
    Bounds bd{};
    for (int s = 0; s <= n_shards; ++s) bd.begin[s] = shard_begin_h[s];
    for (int s = 0; s < n_shards; ++s) {
        MMX_REQUIRE(bd.begin[s + 1] >= bd.begin[s], "shard bounds must not decrease");
        MMX_REQUIRE((bd.begin[s + 1] - bd.begin[s] + 31) / 32 <= wmax, "shard %d is wider than wmax words", s);
    }
    MMX_REQUIRE(words_per_row == (bd.begin[n_shards] + 31) / 32, "words_per_row does not match the column count");
    if (nq == 0) return 0;
    place_bit_columns_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, stream>>>(gathered, n_shards, nq, wmax, bd, words_per_row,
                                                                          above_bits, cnt_above);
    MMX_CHECK_CUDA(cudaGetLastError());
    return 0;
}

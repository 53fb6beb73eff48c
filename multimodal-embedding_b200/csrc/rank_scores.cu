// KR — ground-truth thresholds, strict-greater counts and above-GT bitmasks from a materialised fp32 score matrix.
//
// Replaces the 45 000 per-row torch.topk + membership tests of /root/reference/main.py:331-378
// (compute_metrics_detailed) and v28 :203-229 (compute_metrics): query i is correct at k iff
// #{j : S[i,j] > max_{c in GT(i)} S[i,c]} < k. HBM-bound: one pass over S, one warp per row,
// 128-byte coalesced loads, __ballot_sync packs 32 comparisons into the bitmask word directly.
#include <cfloat>

#include "../../include/mmx.h"
#include "mmx_host.cuh"

namespace mmx {
namespace rank {

constexpr int WARPS = 8;

// thr[q] = max(thr[q], max over GT entries that fall inside this shard's columns [g_base, g_base + ng))
__global__ void gt_threshold_kernel(const float* __restrict__ S, long long nq, long long ng, long long ld,
                                    const int32_t* __restrict__ gt_off, const int32_t* __restrict__ gt_idx,
                                    long long g_base, float* __restrict__ thr) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    float t = thr[q];
    for (int e = gt_off[q]; e < gt_off[q + 1]; ++e) {
        const long long c = (long long)gt_idx[e] - g_base;
        if (c >= 0 && c < ng) t = fmaxf(t, S[q * ld + c]);
    }
    thr[q] = t;
}

__global__ void __launch_bounds__(WARPS * 32)
count_above_kernel(const float* __restrict__ S, long long nq, long long ng, long long ld, const float* __restrict__ thr,
                   int32_t* __restrict__ cnt_above, uint32_t* __restrict__ above_bits, long long words_per_row) {
    const int lane = threadIdx.x & 31;
    const long long q = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (q >= nq) return;
    const float t = thr[q];
    const float* row = S + q * ld;
    const long long nwords = (ng + 31) >> 5;
    int cnt = 0;
    long long w = 0;
    for (; w + 4 <= nwords; w += 4) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long c = ((w + u) << 5) + lane;
            v[u] = c < ng ? __ldg(row + c) : -INFINITY;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t bits = __ballot_sync(0xffffffffu, v[u] > t);
            cnt += __popc(bits);
            if (above_bits != nullptr && lane == u) above_bits[q * words_per_row + w + u] = bits;
        }
    }
    for (; w < nwords; ++w) {
        const long long c = (w << 5) + lane;
        const float v = c < ng ? __ldg(row + c) : -INFINITY;
        const uint32_t bits = __ballot_sync(0xffffffffu, v > t);
        cnt += __popc(bits);
        if (above_bits != nullptr && lane == 0) above_bits[q * words_per_row + w] = bits;
    }
    if (above_bits != nullptr)
        for (long long z = nwords + lane; z < words_per_row; z += 32) above_bits[q * words_per_row + z] = 0u;
    if (lane == 0) cnt_above[q] += cnt;
}

}  // namespace rank
}  // namespace mmx

using namespace mmx;

extern "C" int mmx_rank_from_scores(const float* scores, int64_t nq, int64_t ng, int64_t ld, const int32_t* gt_off,
                                    const int32_t* gt_idx, int64_t g_base, int mode, float* thr, int32_t* cnt_above,
                                    uint32_t* above_bits, int64_t words_per_row, cudaStream_t stream) {
    MMX_REQUIRE(nq >= 0 && ng >= 0 && ld >= ng, "bad shape nq=%lld ng=%lld ld=%lld", (long long)nq, (long long)ng,
                (long long)ld);
    MMX_REQUIRE(thr != nullptr, "thr == NULL");
    if (nq == 0) return 0;
    if (mode == MMX_MODE_THRESHOLD) {
        MMX_REQUIRE(gt_off != nullptr && gt_idx != nullptr, "GT CSR missing");
        rank::gt_threshold_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, stream>>>(scores, nq, ng, ld, gt_off, gt_idx,
                                                                                    g_base, thr);
    } else if (mode == MMX_MODE_COUNT) {
        MMX_REQUIRE(cnt_above != nullptr, "cnt_above == NULL");
        MMX_REQUIRE(above_bits == nullptr || words_per_row >= (ng + 31) / 32, "words_per_row too small");
        rank::count_above_kernel<<<(unsigned)((nq + rank::WARPS - 1) / rank::WARPS), rank::WARPS * 32, 0, stream>>>(
            scores, nq, ng, ld, thr, cnt_above, above_bits, words_per_row);
    } else {
        MMX_REQUIRE(false, "unknown mode %d", mode);
    }
    MMX_CHECK_CUDA(cudaGetLastError());
    return 0;
}

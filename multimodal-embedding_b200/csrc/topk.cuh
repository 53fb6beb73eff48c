// Warp-level top-k primitives shared by KT (topk.cu) and the fused dense top-k (dense_rank.cu).
//
// Order everywhere: higher score first, equal scores by lower column index — a total order, so the result does not
// depend on how the candidates were produced (torch.topk, which the reference calls at main.py:333,362,372, leaves
// the order of tied scores unspecified: tie class T2).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mmx {
namespace topk {

constexpr int MAX_K = 32;

__device__ __forceinline__ bool beats(float v, int i, float w, int j) { return v > w || (v == w && i < j); }

// Sorted list of the best (at most 32) entries seen so far, one entry per lane (lane 0 = best).
struct WarpList {
    float v;
    int i;
    __device__ __forceinline__ void init() {
        v = -INFINITY;
        i = 0x7fffffff;
    }
    // every lane passes the same candidate; it is inserted at its rank, the tail shifts down one lane
    __device__ __forceinline__ void insert(float cv, int ci, int lane) {
        const unsigned better = __ballot_sync(0xffffffffu, beats(v, i, cv, ci));
        const int pos = __popc(better);  // entries that stay ahead of the candidate
        const float uv = __shfl_up_sync(0xffffffffu, v, 1);
        const int ui = __shfl_up_sync(0xffffffffu, i, 1);
        if (lane == pos) {
            v = cv;
            i = ci;
        } else if (lane > pos) {
            v = uv;
            i = ui;
        }
    }
    __device__ __forceinline__ float kth_value(int k) const { return __shfl_sync(0xffffffffu, v, k - 1); }
    __device__ __forceinline__ int kth_index(int k) const { return __shfl_sync(0xffffffffu, i, k - 1); }
};

// Offer one value per lane (column index ci, ascending across calls and lanes); `ok` masks lanes without a value.
__device__ __forceinline__ void offer(WarpList& L, float cv, int ci, bool ok, int k, int lane) {
    float tv = L.kth_value(k);
    int ti = L.kth_index(k);
    unsigned m = __ballot_sync(0xffffffffu, ok && beats(cv, ci, tv, ti));
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const float sv = __shfl_sync(0xffffffffu, cv, src);
        const int si = __shfl_sync(0xffffffffu, ci, src);
        if (beats(sv, si, tv, ti)) {  // the bar may have risen since the ballot
            L.insert(sv, si, lane);
            tv = L.kth_value(k);
            ti = L.kth_index(k);
        }
    }
}

}  // namespace topk
}  // namespace mmx

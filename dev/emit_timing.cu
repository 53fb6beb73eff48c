// Cycle breakdown of emit_long_kernel (K6s long path) per mask epoch: staging, bracket rounds, emit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DMMX_EMIT_TIMING -o dev/emit_timing dev/emit_timing.cu \
//        multimodal-embedding_b200/csrc/mmx_host.cu -lcuda ;  ./dev/emit_timing [n] [P]
// The words are a xorshift stream (uniform 32-bit words are all the replay needs; exactness against numpy is the
// test-suite's business).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../multimodal-embedding_b200/csrc/shuffle.cu"

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 50000;
    const int P = argc > 2 ? atoi(argv[2]) : 200;
    const long long n_words = (long long)(1.6 * n * P) + 200000;
    std::vector<uint32_t> h(n_words);
    uint64_t x = 88172645463325252ull;
    for (auto& w : h) {
        x ^= x << 13;
        x ^= x >> 7;
        x ^= x << 17;
        w = (uint32_t)(x >> 16);
    }
    uint32_t* words;
    int32_t* jout;
    long long* info;
    cudaMalloc(&words, n_words * 4);
    cudaMalloc(&jout, (size_t)P * (n - 1) * 4);
    cudaMalloc(&info, 16);
    const LongPlan lp = make_long_plan(n);
    unsigned long long* rmask;
    uint16_t* rbase;
    long long* ehead;
    const size_t recs = (size_t)P * lp.n_epochs;
    cudaMalloc(&rmask, recs * LT * 8 + 8);
    cudaMalloc(&rbase, recs * LT * 2 + 8);
    cudaMalloc(&ehead, recs * 16 + 16);
    cudaMemcpy(words, h.data(), n_words * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(emit_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (RING + RING_PAD) * 4);
    cudaEvent_t a, b, c2;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventCreate(&c2);
    for (int rep = 0; rep < 2; ++rep) {
        long long zero[MAX_EPOCHS + 1][4] = {};
        cudaMemcpyToSymbol(g_emit_prof, zero, sizeof(zero));
        cudaMemcpyToSymbol(g_emit_round, zero, sizeof(zero));
        cudaEventRecord(a);
        emit_long_kernel<<<1, LT, (RING + RING_PAD) * 4>>>(words, n_words, n, P, lp, info, jout, rmask, rbase, ehead);
        cudaEventRecord(b);
        if (lp.n_epochs > 0) extract_long_kernel<<<(unsigned)recs, LT>>>(words, n, lp, info, rmask, rbase, ehead, jout);
        cudaEventRecord(c2);
        cudaDeviceSynchronize();
    }
    float ms, ms_x;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventElapsedTime(&ms_x, b, c2);
    printf("extract kernel: %.3f ms\n", ms_x);
    long long hi[2];
    cudaMemcpy(hi, info, 16, cudaMemcpyDeviceToHost);
    long long prof[MAX_EPOCHS + 1][4];
    cudaMemcpyFromSymbol(prof, g_emit_prof, sizeof(prof));
    printf("n=%d P=%d: %.3f ms = %.1f us per permutation; info = %lld %lld (%s)\n", n, P, ms, 1e3 * ms / P, hi[0], hi[1],
           cudaGetErrorString(cudaGetLastError()));
    printf("epoch   i_hi      c     W0 | cycles per permutation: stage  rounds   emit+bar | rounds\n");
    double tot = 0;
    for (int e = 0; e <= MAX_EPOCHS; ++e) {
        if (e < MAX_EPOCHS && e >= lp.n_epochs) continue;
        const double st = (double)prof[e][0] / P, rd = (double)prof[e][1] / P, em = (double)prof[e][2] / P;
        if (e < MAX_EPOCHS)
            printf("%5d %6d %6d %6d | %8.0f %8.0f %8.0f | %.2f\n", e, lp.e[e].i_hi, lp.e[e].c, lp.e[e].W0, st, rd, em - rd,
                   (double)prof[e][3] / P);
        else
            printf(" tail %6d                | %8.0f %8s %8.0f |\n", lp.i_tail, st, "-", em);
        tot += st + em;
    }
    printf("sum %.0f cycles per permutation\n", tot);
    long long rp[MAX_EPOCHS + 1][4];
    cudaMemcpyFromSymbol(rp, g_emit_round, sizeof(rp));
    printf("thread 0, cycles per ROUND: walk   scan+sts  barrier  base\n");
    for (int e = 0; e < lp.n_epochs; ++e) {
        const double nr = (double)prof[e][3] / P + 1.0;
        printf("%5d                    %8.0f %8.0f %8.0f %8.0f\n", e, rp[e][0] / (nr * P), rp[e][1] / (nr * P),
               rp[e][2] / (nr * P), rp[e][3] / (nr * P));
    }
    return 0;
}

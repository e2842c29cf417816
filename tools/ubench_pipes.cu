// micro-benchmark: which pipe do VIMNMX3.U16x2 / HMNMX2 / IMAD run on (do they overlap)?  tools only.
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(uint32_t* out, uint32_t seed, int iters) {
    uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 ^ 0x1234, b1 = a1 ^ 0x777, b2 = a2 ^ 0x999, b3 = a3 ^ 0xabc;
    __half2 h0 = __halves2half2(__ushort_as_half(0x6400 | (a0 & 255)), __ushort_as_half(0x6400 | (a1 & 255))), h1 = h0, h2 = h0, h3 = h0;
    __half2 g0 = __halves2half2(__ushort_as_half(0x6400 | (a2 & 255)), __ushort_as_half(0x6400 | (a3 & 255))), g1 = g0, g2 = g0, g3 = g0;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            if (MODE == 0 || MODE == 2) {
                a0 = __vimax3_u16x2(a0, b0, a1); a1 = __vimin3_u16x2(a1, b1, a2); a2 = __vimax3_u16x2(a2, b2, a3); a3 = __vimin3_u16x2(a3, b3, a0);
            }
            if (MODE == 1 || MODE == 2) {
                h0 = __hmax2(h0, g0); h1 = __hmin2(h1, g1); h2 = __hmax2(h2, g2); h3 = __hmin2(h3, g3);
                g0 = __hmin2(g0, h1); g1 = __hmax2(g1, h2); g2 = __hmin2(g2, h3); g3 = __hmax2(g3, h0);
            }
            if (MODE == 3 || MODE == 4) {   // IMAD chain (fma pipe)
                b0 = b0 * 3u + a0; b1 = b1 * 5u + a1; b2 = b2 * 7u + a2; b3 = b3 * 9u + a3;
            }
            if (MODE == 4) {
                a0 = __vimax3_u16x2(a0, b0, a1); a1 = __vimin3_u16x2(a1, b1, a2); a2 = __vimax3_u16x2(a2, b2, a3); a3 = __vimin3_u16x2(a3, b3, a0);
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ b0 ^ b1 ^ b2 ^ b3 ^ (uint32_t)__half_as_ushort(__low2half(h0 + h1 + h2 + h3 + g0 + g1 + g2 + g3));
}
template <int MODE> float run(uint32_t* d, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(d, 1, 10);
    cudaEventRecord(e0); k<MODE><<<148 * 8, 256>>>(d, 1, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    uint32_t* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    const int it = 2000;
    const double warps = 148.0 * 8 * 8;
    float t0 = run<0>(d, it), t1 = run<1>(d, it), t2 = run<2>(d, it), t3 = run<3>(d, it), t4 = run<4>(d, it);
    auto rate = [&](float ms, double instr_per_iter) { return warps * it * 16 * instr_per_iter / (ms * 1e-3) / 1e12; };
    printf("VIMNMX3.U16x2 only : %.3f ms  %.3f Twarp-inst/s\n", t0, rate(t0, 4));
    printf("HMNMX2 only        : %.3f ms  %.3f Twarp-inst/s\n", t1, rate(t1, 8));
    printf("both interleaved   : %.3f ms  (sum of separate %.3f) %.3f Twarp-inst/s\n", t2, t0 + t1, rate(t2, 12));
    printf("IMAD only          : %.3f ms  %.3f Twarp-inst/s\n", t3, rate(t3, 4));
    printf("IMAD + VIMNMX3     : %.3f ms  (sum of separate %.3f)\n", t4, t0 + t3);
    return 0;
}

// Does a global store (or RED) to an L1-resident line keep the line in L1?  Dependent-load chain, one thread.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(uint32_t *p, int mode, int iters, long long *out, uint32_t *sink) {
    // p: 32 words (one line), all zero.  chain: idx = p[idx] (always 0..31 via value&31)
    uint32_t idx = threadIdx.x & 31;
    uint32_t acc = 0;
    // warm
    for (int i = 0; i < 4; ++i) { uint32_t v = p[idx]; idx = (idx + v + 1) & 31; }
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint32_t v = p[idx];                       // dependent load
        acc += v;
        if (mode == 1) p[32 + ((idx + 7) & 31)] = 0;          // store to ANOTHER line
        if (mode == 2) p[(idx + 7) & 31] = 0;                 // store to the SAME line, other word
        if (mode == 3) p[idx] = 0;                            // store to the same word
        if (mode == 4) atomicOr(&p[(idx + 7) & 31], 0u);      // RED to the same line
        if (mode == 5) atomicOr(&p[64 + ((idx + 7) & 31)], 0u);   // RED to another line
        if (mode == 6) asm volatile("st.global.wb.u32 [%0], %1;" ::"l"(p + ((idx + 7) & 31)), "r"(0u));
        idx = (idx + v + 1) & 31;
    }
    long long t1 = clock64();
    out[0] = t1 - t0;
    *sink = acc + idx;
}
int main() {
    uint32_t *p; long long *out; uint32_t *sink;
    cudaMalloc(&p, 4096); cudaMemset(p, 0, 4096); cudaMalloc(&out, 8); cudaMalloc(&sink, 4);
    const char *names[] = {"load chain only", "+store other line", "+store same line", "+store same word", "+RED same line", "+RED other line", "+st.wb same line"};
    for (int mode = 0; mode < 7; ++mode) {
        k<<<1, 1>>>(p, mode, 2000, out, sink);
        cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
        printf("%-22s %.1f cycles/iter\n", names[mode], h / 2000.0);
    }
    return 0;
}

// the pheromone kernel's round (compaction + ordered fold) in isolation: cycles per round / per hit for several forms of the fold
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o /tmp/fold_round tools/ubench/fold_round.cu && /tmp/fold_round
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
struct __align__(16) Entry { double d; uint32_t w; uint32_t pad; };
template <int MODE>
__global__ void __launch_bounds__(256) rounds(const uint32_t *__restrict__ words, const double *__restrict__ dep, int k, double *out, long long *cyc) {
    __shared__ Entry s_buf[8][36];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    Entry *sb = s_buf[wid];
    const uint32_t lt = (1u << lane) - 1u;
    double t = 0.1;
    long long c_fold = 0, c_all = clock64();
    int hits = 0;
    for (int base = 0; base < k; base += 32) {
        const uint32_t word = words[(base + lane) * 8 + wid];
        const double d = dep[base + lane];
        const uint32_t nz = __ballot_sync(0xffffffffu, word != 0u);
        if (!nz) continue;
        const int n = __popc(nz);
        if (MODE >= 5) {
            uint32_t m = word;
            if (MODE != 6) {
#pragma unroll
                for (int j = 16, mk = 0x0000FFFF; j; j >>= 1, mk ^= mk << j) {
                    const uint32_t y = __shfl_xor_sync(0xffffffffu, m, j);
                    m = (lane & j) ? ((m & ~(uint32_t)mk) | ((y >> j) & (uint32_t)mk)) : ((m & (uint32_t)mk) | ((y << j) & ~(uint32_t)mk));
                }
            }
            double *sd = (double *)sb;
            sd[lane] = d;
            __syncwarp();
            hits += n;
            const long long c0 = clock64();
            if (MODE == 7) { t += (double)m; }
            else {
                const double2 *sd2 = (const double2 *)sd;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    if (!(nz & (0xFFu << (8 * g)))) continue;
                    const double2 v0 = sd2[4 * g], v1 = sd2[4 * g + 1], v2 = sd2[4 * g + 2], v3 = sd2[4 * g + 3];
                    const uint32_t mg = m >> (8 * g);
                    const double a0 = (mg & 1u) ? v0.x : 0.0, a1 = (mg & 2u) ? v0.y : 0.0, a2 = (mg & 4u) ? v1.x : 0.0;
                    const double a3 = (mg & 8u) ? v1.y : 0.0, a4 = (mg & 16u) ? v2.x : 0.0, a5 = (mg & 32u) ? v2.y : 0.0;
                    const double a6 = (mg & 64u) ? v3.x : 0.0, a7 = (mg & 128u) ? v3.y : 0.0;
                    t += a0; t += a1; t += a2; t += a3; t += a4; t += a5; t += a6; t += a7;
                }
            }
            c_fold += clock64() - c0;
            __syncwarp();
            continue;
        }
        if (word != 0u) { Entry e; e.d = d; e.w = word; e.pad = 0u; sb[__popc(nz & lt)] = e; }
        if (lane < 4) { Entry z; z.d = 0.0; z.w = 0u; z.pad = 0u; sb[n + lane] = z; }
        __syncwarp();
        hits += n;
        const long long c0 = clock64();
        if (MODE == 0) {            // ping-pong groups of four, struct loads
            Entry x[4], y[4];
            auto load4 = [&](Entry (&X)[4], int q) {
#pragma unroll
                for (int j = 0; j < 4; ++j) X[j] = sb[q + j];
            };
            auto add4 = [&](const Entry (&X)[4]) {
                double a[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) a[j] = ((X[j].w >> lane) & 1u) ? X[j].d : 0.0;
                t += a[0]; t += a[1]; t += a[2]; t += a[3];
            };
            load4(x, 0);
            for (int q = 4;; q += 8) {
                if (q < n) load4(y, q);
                add4(x);
                if (q >= n) break;
                if (q + 4 < n) load4(x, q + 4);
                add4(y);
                if (q + 4 >= n) break;
            }
        } else if (MODE == 1) {     // plain loop
            for (int q = 0; q < n; ++q) { const Entry x = sb[q]; t += ((x.w >> lane) & 1u) ? x.d : 0.0; }
        } else if (MODE == 2) {     // always 32 entries, fully unrolled (entries >= n: stale but masked by q < n)
#pragma unroll
            for (int q = 0; q < 32; ++q) { const Entry x = sb[q]; t += (q < n && ((x.w >> lane) & 1u)) ? x.d : 0.0; }
        } else if (MODE == 3) {     // no select: add everything (not the kernel's arithmetic; isolates the select)
            for (int q = 0; q < n; ++q) { t += sb[q].d; }
        } else if (MODE == 4) {     // uint4 loads
            const uint4 *s4 = (const uint4 *)sb;
            for (int q = 0; q < n; q += 4) {
                const uint4 e0 = s4[q], e1 = s4[q + 1], e2 = s4[q + 2], e3 = s4[q + 3];
                const double d0 = __hiloint2double(e0.y, e0.x), d1 = __hiloint2double(e1.y, e1.x), d2 = __hiloint2double(e2.y, e2.x), d3 = __hiloint2double(e3.y, e3.x);
                t += ((e0.z >> lane) & 1u) ? d0 : 0.0; t += ((e1.z >> lane) & 1u) ? d1 : 0.0;
                t += ((e2.z >> lane) & 1u) ? d2 : 0.0; t += ((e3.z >> lane) & 1u) ? d3 : 0.0;
            }
        }
        c_fold += clock64() - c0;
        __syncwarp();
    }
    c_all = clock64() - c_all;
    out[blockIdx.x * 256 + threadIdx.x] = t;
    if (lane == 0) { cyc[(blockIdx.x * 8 + wid) * 3] = c_all; cyc[(blockIdx.x * 8 + wid) * 3 + 1] = c_fold; cyc[(blockIdx.x * 8 + wid) * 3 + 2] = hits; }
}
template <int MODE> void run(const char *name, const uint32_t *w, const double *d, int k, double *out, long long *cyc, int ctas, int threads = 256) {
    for (int r = 0; r < 2; ++r) rounds<MODE><<<ctas, threads>>>(w, d, k, out, cyc);
    cudaDeviceSynchronize();
    long long h[24];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    if (threads < 256) { printf("%-44s ctas %4d threads %3d: warp0 %6lld cycles/round, fold %6.0f cycles/round\n", name, ctas, threads, h[0] / (k / 32), (double)h[1] / (k / 32)); return; }
    printf("%-44s ctas %4d: warp0 %6lld cycles/round, fold %6.0f cycles/round = %5.1f cycles/hit; warp7 fold %5.1f cycles/hit\n", name, ctas,
           h[0] / (k / 32), (double)h[1] / (k / 32), (double)h[1] / h[2], (double)h[22] / h[23]);
}
int main() {
    const int k = 2272;
    uint32_t *hw = new uint32_t[k * 8]; double *hd = new double[k];
    uint32_t s = 12345u;
    for (int i = 0; i < k * 8; ++i) { s = s * 1664525u + 1013904223u; const uint32_t r = s >> 8; hw[i] = (r % 100 < 95) ? (1u << (r % 32)) | (1u << ((r >> 5) % 32)) : 0u; }
    for (int i = 0; i < k; ++i) hd[i] = 2.5 / (700.0 + i % 500);
    uint32_t *w; double *d, *out; long long *cyc;
    cudaMalloc(&w, k * 8 * 4); cudaMalloc(&d, k * 8); cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 1 << 20);
    cudaMemcpy(w, hw, k * 8 * 4, cudaMemcpyHostToDevice); cudaMemcpy(d, hd, k * 8, cudaMemcpyHostToDevice);
    for (int th : {32, 64, 128, 256}) {
        run<5>("transpose + unrolled predicated fold", w, d, k, out, cyc, 1, th);
        run<6>("unrolled predicated fold, no transpose", w, d, k, out, cyc, 1, th);
        run<7>("transpose only", w, d, k, out, cyc, 1, th);
        run<1>("plain loop", w, d, k, out, cyc, 1, th);
        run<3>("no select", w, d, k, out, cyc, 1, th);
    }
    for (int ctas : {1, 148}) {
        run<0>("ping-pong groups of 4 (kernel)", w, d, k, out, cyc, ctas);
        run<1>("plain loop", w, d, k, out, cyc, ctas);
        run<2>("32 entries unrolled", w, d, k, out, cyc, ctas);
        run<3>("no select", w, d, k, out, cyc, ctas);
        run<4>("uint4 loads, groups of 4", w, d, k, out, cyc, ctas);
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

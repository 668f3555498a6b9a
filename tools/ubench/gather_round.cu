// What does a round's LOAD cost in the pheromone kernel?  (a) each row warp gathers its row word from 32 ants' slabs (32 lines
// per load instruction) + the 32 deposits, D rounds in flight; (b) the CTA's 8 warps load the 8-row sectors of the round's 32
// ants cooperatively (4 ants x 32 B per warp = 4 sectors per instruction) into shared memory.  No other work: cycles per round.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/gather_round tools/ubench/gather_round.cu && /tmp/gather_round
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int D>
__global__ void __launch_bounds__(256) gatherA(const uint32_t *__restrict__ slabs, const double *__restrict__ dep, const int *__restrict__ list, int k,
                                               unsigned *out, long long *cyc, int n_ants) {
    __shared__ int s_list[4096];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < k; i += 256) s_list[i] = list[i];
    __syncthreads();
    const uint32_t *slab_t = slabs + (size_t)blockIdx.x * n_ants * 32 + wid;   // one "tile" per CTA
    uint32_t w_n[D]; double d_n[D];
    unsigned acc = 0; double dacc = 0.0;
    const long long c0 = clock64();
#pragma unroll
    for (int u = 0; u < D; ++u) { const int a = s_list[u * 32 + lane]; w_n[u] = slab_t[(size_t)a * 32]; d_n[u] = dep[a]; }
    for (int base = 0; base < k; base += 32 * D) {
#pragma unroll
        for (int u = 0; u < D; ++u) {
            const uint32_t w = w_n[u]; const double d = d_n[u];
            const int i = base + 32 * D + u * 32 + lane;
            if (i < k) { const int a = s_list[i]; w_n[u] = slab_t[(size_t)a * 32]; d_n[u] = dep[a]; }
            acc += __popc(__ballot_sync(0xffffffffu, w != 0u)); dacc += d;
        }
    }
    const long long c1 = clock64();
    out[blockIdx.x * 256 + threadIdx.x] = acc + (unsigned)dacc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}
// cooperative: warp w loads ants 4w..4w+3 of the round: lane = (ant j = lane >> 3, row word = lane & 7); smem tile [2][8 rows][32 ants]
__global__ void __launch_bounds__(256) gatherB(const uint32_t *__restrict__ slabs, const double *__restrict__ dep, const int *__restrict__ list, int k,
                                               unsigned *out, long long *cyc, int n_ants, int rg) {
    __shared__ int s_list[4096];
    __shared__ uint32_t s_w[2][8][32];
    __shared__ double s_d[2][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < k; i += 256) s_list[i] = list[i];
    __syncthreads();
    const uint32_t *slab_t = slabs + (size_t)blockIdx.x * n_ants * 32 + rg * 8;
    unsigned acc = 0; double dacc = 0.0;
    const int j = wid * 4 + (lane >> 3), rw = lane & 7;
    const long long c0 = clock64();
    uint32_t wv = slab_t[(size_t)s_list[j] * 32 + rw];
    double dv = (threadIdx.x < 32) ? dep[s_list[lane]] : 0.0;
    for (int base = 0, par = 0; base < k; base += 32, par ^= 1) {
        s_w[par][rw][j] = wv;
        if (threadIdx.x < 32) s_d[par][lane] = dv;
        const int nb = base + 32;
        if (nb < k) { wv = slab_t[(size_t)s_list[nb + j] * 32 + rw]; if (threadIdx.x < 32) dv = dep[s_list[nb + lane]]; }
        __syncthreads();
        const uint32_t w = s_w[par][wid][lane]; const double d = s_d[par][lane];
        acc += __popc(__ballot_sync(0xffffffffu, w != 0u)); dacc += d;
    }
    const long long c1 = clock64();
    out[blockIdx.x * 256 + threadIdx.x] = acc + (unsigned)dacc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}
int main() {
    const int n_ants = 4096, k = 4096, tiles = 1036;
    uint32_t *slabs; double *dep; int *list; unsigned *out; long long *cyc;
    cudaMalloc(&slabs, (size_t)tiles * n_ants * 128); cudaMemset(slabs, 1, (size_t)tiles * n_ants * 128);
    cudaMalloc(&dep, n_ants * 8); cudaMemset(dep, 0, n_ants * 8);
    cudaMalloc(&list, k * 4); cudaMalloc(&out, tiles * 256 * 4); cudaMalloc(&cyc, tiles * 8);
    int *hl = new int[k]; for (int i = 0; i < k; ++i) hl[i] = i; cudaMemcpy(list, hl, k * 4, cudaMemcpyHostToDevice);
    long long h[4];
    for (int ctas : {1, 4, 148, 1036}) {
#define RUNA(D) for (int r = 0; r < 3; ++r) gatherA<D><<<ctas, 256>>>(slabs, dep, list, k, out, cyc, n_ants); cudaDeviceSynchronize(); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); \
        printf("gather per row warp, %d round(s) in flight, %4d CTAs: %6.0f cycles per round\n", D, ctas, (double)h[0] / (k / 32));
        RUNA(1) RUNA(2) RUNA(4) RUNA(8)
        for (int r = 0; r < 3; ++r) gatherB<<<ctas, 256>>>(slabs, dep, list, k, out, cyc, n_ants, 1);
        cudaDeviceSynchronize(); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("cooperative sector loads + smem,      %4d CTAs: %6.0f cycles per round\n", ctas, (double)h[0] / (k / 32));
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
}

// dependent-chain latency of fp64 adds on sm_100a, alone and with 1..16 warps per SM sharing the FP64 pipe
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o /tmp/dadd_chain tools/ubench/dadd_chain.cu && /tmp/dadd_chain
#include <cstdio>
#include <cuda_runtime.h>
__global__ void chain(double *out, long long *cyc, const double *in, int n) {
    double t = in[0];
    const double d = in[1 + (threadIdx.x & 3)];
    long long c0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) t += d;
    long long c1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
    if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}
// select + add, the fold's form
__global__ void chain_sel(double *out, long long *cyc, const double *in, const unsigned *w, int n) {
    double t = in[0];
    __shared__ double sd[64]; __shared__ unsigned sw[64];
    if (threadIdx.x < 64) { sd[threadIdx.x] = in[1 + (threadIdx.x & 3)]; sw[threadIdx.x] = w[threadIdx.x]; }
    __syncthreads();
    const unsigned lane = threadIdx.x & 31;
    long long c0 = clock64();
    for (int i = 0; i < n; i += 64) {
#pragma unroll
        for (int q = 0; q < 64; ++q) t += ((sw[q] >> lane) & 1u) ? sd[q] : 0.0;
    }
    long long c1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
    if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}
__global__ void chain_f32(float *out, long long *cyc, const float *in, int n) {
    float t = in[0]; const float d = in[1];
    long long c0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) t += d;
    long long c1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
    if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}
int main() {
    double *out, *in; long long *cyc; unsigned *w; float *fo, *fi;
    cudaMalloc(&out, 1 << 22); cudaMalloc(&in, 64); cudaMalloc(&cyc, 8192); cudaMalloc(&w, 256); cudaMalloc(&fo, 1 << 22); cudaMalloc(&fi, 64);
    double hin[5] = {0.1, 1e-3, 2e-3, 3e-3, 4e-3}; float hfi[2] = {0.1f, 1e-3f}; unsigned hw[64];
    for (int i = 0; i < 64; ++i) hw[i] = 0x9e3779b9u * (i + 1);
    cudaMemcpy(in, hin, sizeof hin, cudaMemcpyHostToDevice); cudaMemcpy(w, hw, sizeof hw, cudaMemcpyHostToDevice); cudaMemcpy(fi, hfi, sizeof hfi, cudaMemcpyHostToDevice);
    const int n = 1 << 16;
    long long h[1024];
    for (int threads : {32, 64, 128, 256, 512, 1024}) {
        for (int rep = 0; rep < 2; ++rep) chain<<<148, threads>>>(out, cyc, in, n);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
        printf("DADD chain, %4d threads/SM (1 CTA/SM): %.2f cycles per add per warp\n", threads, (double)h[0] / n);
    }
    for (int threads : {32, 256}) {
        for (int rep = 0; rep < 2; ++rep) chain_sel<<<148, threads>>>(out, cyc, in, w, n);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
        printf("LDS + select + DADD chain, %4d threads/SM: %.2f cycles per add\n", threads, (double)h[0] / n);
    }
    for (int rep = 0; rep < 2; ++rep) chain_f32<<<148, 32>>>(fo, cyc, fi, n);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("FADD chain, 32 threads: %.2f cycles per add\n", (double)h[0] / n);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

// Micro-benchmarks behind the LocalBA solver design: fp64 FMA latency / throughput, fp64 divide latency, barrier and cluster-barrier cost,
// shared-memory latency on one SM of a B200. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_fp64.bin tools/ubench_fp64.cu
#include <cooperative_groups.h>
#include <cstdio>
namespace cg = cooperative_groups;

__global__ void k_dfma_lat(double* out, long long* clk, double a, double b) {
    double x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) x = fma(x, a, b);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_dfma_thr(double* out, long long* clk, double a, double b) {
    double x[16];
    for (int j = 0; j < 16; j++) x[j] = out[threadIdx.x] + j;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) x[j] = fma(x[j], a, b);
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0;
    for (int j = 0; j < 16; j++) s += x[j];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_ddiv_lat(double* out, long long* clk, double a) {
    double x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) {
#pragma unroll
        for (int j = 0; j < 4; j++) x = a / x;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_ffma_lat(float* out, long long* clk, float a, float b) {
    float x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) x = fmaf(x, a, b);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_barrier(long long* clk) {
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; i++) __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_smem_lat(int* out, long long* clk) {
    __shared__ int chain[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) chain[i] = (i * 37 + 11) & 1023;
    __syncthreads();
    int p = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; i++) p = chain[p];
    long long t1 = clock64();
    out[threadIdx.x] = p;
    if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_cluster_barrier(long long* clk) {
    cg::cluster_group c = cg::this_cluster();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) c.sync();
    long long t1 = clock64();
    if (threadIdx.x == 0 && c.block_rank() == 0) clk[0] = t1 - t0;
}
__global__ void k_l2_lat(const int* chain, int* out, long long* clk) {
    int p = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) p = __ldcg(chain + p);
    long long t1 = clock64();
    out[threadIdx.x] = p;
    if (threadIdx.x == 0) clk[0] = t1 - t0;
}

int main() {
    double* d; float* f; long long* c; int* ii; int* chain;
    cudaMalloc(&d, 8192); cudaMalloc(&f, 4096); cudaMalloc(&c, 64); cudaMalloc(&ii, 4096); cudaMalloc(&chain, 1 << 22);
    cudaMemset(d, 0, 8192); cudaMemset(f, 0, 4096);
    { int* h = new int[1 << 20]; for (int i = 0; i < (1 << 20); i++) h[i] = (int)(((long long)i * 7919 + 104729) & ((1 << 20) - 1)); cudaMemcpy(chain, h, 1 << 22, cudaMemcpyHostToDevice); delete[] h; }
    long long h;
    auto rd = [&]() { cudaDeviceSynchronize(); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); return (double)h; };
    for (int rep = 0; rep < 2; rep++) {
        k_dfma_lat<<<1, 32>>>(d, c, 1.0000001, 1e-9); double a = rd();
        if (rep) printf("DFMA dependent latency      %.1f clk\n", a / 4096);
        for (int T : {32, 128, 256, 512, 1024}) { k_dfma_thr<<<1, T>>>(d, c, 1.0000001, 1e-9); a = rd(); if (rep) printf("DFMA throughput %4d threads  %.2f clk per warp-instruction (%.1f DFMA/clk/SM)\n", T, a / (4096.0 * (T / 32)), 4096.0 * T / a); }
        k_ddiv_lat<<<1, 32>>>(d, c, 1.37); a = rd(); if (rep) printf("DDIV dependent latency      %.1f clk\n", a / 1024);
        k_ffma_lat<<<1, 32>>>(f, c, 1.0000001f, 1e-9f); a = rd(); if (rep) printf("FFMA dependent latency      %.1f clk\n", a / 4096);
        for (int T : {32, 256, 1024}) { k_barrier<<<1, T>>>(c); a = rd(); if (rep) printf("__syncthreads %4d threads   %.1f clk\n", T, a / 1024); }
        k_smem_lat<<<1, 32>>>(ii, c); a = rd(); if (rep) printf("shared load latency         %.1f clk\n", a / 1024);
        k_l2_lat<<<1, 32>>>(chain, ii, c); a = rd(); if (rep) printf("L2 (ld.cg) latency          %.1f clk\n", a / 256);
        for (int C : {2, 4, 8}) {
            cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(C); cfg.blockDim = dim3(256);
            cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = C; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
            cfg.attrs = &at; cfg.numAttrs = 1;
            cudaLaunchKernelEx(&cfg, k_cluster_barrier, c); a = rd(); if (rep) printf("cluster.sync %d CTAs x 256    %.1f clk\n", C, a / 256);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

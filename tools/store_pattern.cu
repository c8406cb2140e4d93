// Write-only bandwidth ceilings for the Gram build: (a) linear STG.128 fill, (b) 64x64 tiles written by one CTA each
// (rows of 512 B), (c) symmetric pattern: lower tiles + mirrored tiles.  nvcc -arch=sm_100a -O3 -o store_pattern store_pattern.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void fill_linear(double2 *p, size_t n2) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n2; i += st) p[i] = make_double2(1.0, 2.0);
}
__global__ void fill_tiles(double *K, long n, int tiles, int sym) {
    int tm, tn;
    if (sym) {
        long b = blockIdx.x;
        tm = (int)((sqrtf(8.f * b + 1.f) - 1.f) * .5f);
        while ((long)(tm + 1) * (tm + 2) / 2 <= b) tm++;
        while ((long)tm * (tm + 1) / 2 > b) tm--;
        tn = (int)(b - (long)tm * (tm + 1) / 2);
    } else { tm = blockIdx.x / tiles; tn = blockIdx.x % tiles; }
    int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    for (int a = 0; a < 4; a++) {
        double *row = K + ((long)tm * 64 + ty + 16 * a) * n + (long)tn * 64 + 2 * tx;
        *(double2 *)row = make_double2(1.0, 2.0);
        *(double2 *)(row + 32) = make_double2(3.0, 4.0);
    }
    if (sym && tm != tn) {
        int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int rr = warp; rr < 64; rr += 8) {
            double *row = K + ((long)tn * 64 + rr) * n + (long)tm * 64 + lane;
            row[0] = 1.0; row[32] = 2.0;
        }
    }
}
int main() {
    long n = 20032; size_t bytes = (size_t)n * n * 8;
    double *K; cudaMalloc(&K, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0); fill_linear<<<148 * 8, 256>>>((double2 *)K, bytes / 16); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); printf("linear fill: %.3f ms %.0f GB/s\n", ms, bytes / ms / 1e6);
        cudaEventRecord(e0); cudaMemsetAsync(K, 0, bytes); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); printf("cudaMemset: %.3f ms %.0f GB/s\n", ms, bytes / ms / 1e6);
        int t = n / 64;
        cudaEventRecord(e0); fill_tiles<<<t * t, 256>>>(K, n, t, 0); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); printf("tiles full: %.3f ms %.0f GB/s\n", ms, bytes / ms / 1e6);
        cudaEventRecord(e0); fill_tiles<<<t * (t + 1) / 2, 256>>>(K, n, t, 1); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); printf("tiles sym: %.3f ms %.0f GB/s\n", ms, bytes / ms / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

// Microbenchmark: FP64 peaks on B200 (sm_100a) used as roofline denominators.
//   - DMMA m8n8k4 (mma.sync f64) register-resident loop
//   - DFMA register-resident loop
//   - exp() fp64 throughput
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --cudart shared -o peaks_fp64 peaks_fp64.cu
// (shared runtime: a statically linked binary carries the runtime's whole symbol table into the tree that is shipped to
// the GPU box).  The same loops are part of the library as lgp_peak_probe, which bench.py times in every run.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dmma(double *out, int iters) {
    double c[NACC][2];
    double a = threadIdx.x * 1e-9, b = 1.0 + threadIdx.x * 1e-9;
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dfma(double *out, int iters) {
    double c[NACC];
    double a = 1.0 + threadIdx.x * 1e-12, b = threadIdx.x * 1e-9;
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_exp(double *out, int iters) {
    double x = -1e-3 * (threadIdx.x + 1), s = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) { s += exp(x); x -= 1e-4; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sm = p.multiProcessorCount;
    printf("device %s SMs %d clock %d kHz\n", p.name, sm, p.clockRate);
    double *out; CK(cudaMalloc(&out, sizeof(double) * sm * 8 * 256));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int bps = 1; bps <= 4; bps *= 2) {
        int iters = 20000;
        // DMMA
        k_dmma<8><<<sm * bps, 256>>>(out, 100);
        CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int r = 0; r < 5; r++) {
            cudaEventRecord(e0); k_dmma<8><<<sm * bps, 256>>>(out, iters); cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        double flops = 2.0 * 256 * 8.0 * iters * (double)(sm * bps) * 8;  // per warp-MMA 8*8*4 FMA
        printf("DMMA m8n8k4  blocks/SM=%d  %.3f ms  %.2f TFLOP/s\n", bps, best, flops / best / 1e9);
        k_dmma<16><<<sm * bps, 256>>>(out, 100);
        best = 1e30f;
        for (int r = 0; r < 5; r++) {
            cudaEventRecord(e0); k_dmma<16><<<sm * bps, 256>>>(out, iters); cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        flops = 2.0 * 256 * 16.0 * iters * (double)(sm * bps) * 8;
        printf("DMMA m8n8k4 (16 acc) blocks/SM=%d  %.3f ms  %.2f TFLOP/s\n", bps, best, flops / best / 1e9);
        // DFMA
        k_dfma<16><<<sm * bps, 256>>>(out, 100);
        best = 1e30f;
        for (int r = 0; r < 5; r++) {
            cudaEventRecord(e0); k_dfma<16><<<sm * bps, 256>>>(out, iters); cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        flops = 2.0 * 16.0 * iters * (double)(sm * bps) * 256;
        printf("DFMA         blocks/SM=%d  %.3f ms  %.2f TFLOP/s\n", bps, best, flops / best / 1e9);
        // exp
        k_exp<<<sm * bps, 256>>>(out, 10);
        best = 1e30f;
        for (int r = 0; r < 5; r++) {
            cudaEventRecord(e0); k_exp<<<sm * bps, 256>>>(out, 2000); cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        double nexp = 8.0 * 2000 * (double)(sm * bps) * 256;
        printf("exp(f64)     blocks/SM=%d  %.3f ms  %.2f Gexp/s  (=%.2f TB/s of 8-byte outputs)\n", bps, best, nexp / best / 1e6, nexp * 8 / best / 1e9);
    }
    // sustained DMMA for ~2 s to see power-capped rate
    {
        cudaEventRecord(e0);
        for (int r = 0; r < 40; r++) k_dmma<16><<<sm * 2, 256>>>(out, 100000);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        double flops = 40.0 * 2.0 * 256 * 16.0 * 100000 * (double)(sm * 2) * 8;
        printf("DMMA sustained %.1f ms  %.2f TFLOP/s\n", ms, flops / ms / 1e9);
    }
    return 0;
}

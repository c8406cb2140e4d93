// Internal (non-ABI) declarations shared between the translation units of liblgpb200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lgp {

constexpr int NB = 128;  // Cholesky leaf / distribution block

// gemm_dmma.cu
struct GemmMirror;  // gemm_dmma.cuh: extra (peer / multicast) destinations of the epilogue
int gemm_launch(cudaStream_t stream, bool a_kmaj, bool b_kmaj, int M, int N, int K, double alpha, const double *A,
                int64_t lda, const double *B, int64_t ldb, double *C, int64_t ldc, int flags,
                const GemmMirror *mir = nullptr);

}  // namespace lgp

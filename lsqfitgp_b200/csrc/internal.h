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
                const GemmMirror *mir = nullptr, const double *scale = nullptr);


// ---- per-device bookkeeping (a process may drive several GPUs: stream pools, kernel attributes and events belong to ONE)
constexpr int MAX_DEVICES = 64;
int current_device();  // cudaGetDevice, -1 on error

// "done once on this device" flag (kernel attributes such as the >48 KB shared-memory opt-in are per device)
struct DeviceOnce {
    unsigned long long mask = 0;  // guarded by once_mutex() in capi.cu
    bool done(int dev) const;
    void set(int dev);
};

// Reusable disable-timing event of the calling host thread on the current device.  Events are handed out round-robin
// from a thread-local ring (64 per device): a record -> wait pair issued by one thread never has more than a few
// other acquisitions in between, and a wait captures the state of the event at the time it is enqueued, so re-recording
// a recycled event later is harmless.  Returns nullptr on failure.
cudaEvent_t ring_event();

}  // namespace lgp

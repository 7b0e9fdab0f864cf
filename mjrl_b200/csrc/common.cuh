// Shared device/host helpers for the mjrl_b200 CUDA engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace mjb {

constexpr int kThreads = 256;      // every tile kernel runs 256-thread CTAs
constexpr int kChunk = 32;         // input features staged per chunk of the first layer

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of a double; result valid in thread 0.  scratch: >= 32 doubles of shared memory.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    v = warp_sum(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    if (w == 0) {
        v = (lane < (int)(blockDim.x >> 5)) ? scratch[lane] : 0.0;
        v = warp_sum(v);
    }
    return v;
}

// Cluster-wide barrier with release/acquire at CLUSTER scope (what distributed-shared-memory exchange needs).
// cooperative_groups' cluster.sync() additionally emits a gpu-scope MEMBAR per call, which costs ~1-2k cycles.
__device__ __forceinline__ void cluster_sync_relacq() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// Layout of a "prepped" (padded / transposed) parameter set of a 2-hidden-layer net, in floats.
// All three kernels modes read the same layout for theta (new), theta_old and the CG tangent v.
struct PrepLayout {
    int H;        // padded hidden width (template value of the kernel)
    int K0;       // real input width (obs_dim, or obs_dim+4 for the baseline net)
    int K0P;      // K0 rounded up to kChunk
    int A;        // real output width
    int AP;       // A rounded up to 8
    int YR;       // rows of the natural last-layer matrix: AP rounded up to the weight-slice height
    int h1, h2;   // real hidden widths
    // offsets (floats) inside the prepped buffer
    int oW1T, ob1, oW2T, ob2, oW3T, ob3, oW2N, oW3N, oLS, total;
    // offsets inside the flat theta vector (reference layout, gaussian_mlp.py:37)
    int tW1, tb1, tW2, tb2, tW3, tb3, tLS, d;
};

inline PrepLayout make_prep_layout(int H, int K0, int A, int h1, int h2, bool has_log_std) {
    PrepLayout L;
    L.H = H; L.K0 = K0; L.K0P = round_up(K0, kChunk); L.A = A; L.AP = round_up(A, 8);
    int SR = 1024 / H;
    L.YR = round_up(L.AP, SR);
    L.h1 = h1; L.h2 = h2;
    int o = 0;
    L.oW1T = o; o += L.K0P * H;
    L.ob1 = o;  o += H;
    L.oW2T = o; o += H * H;
    L.ob2 = o;  o += H;
    L.oW3T = o; o += H * L.AP;
    L.ob3 = o;  o += L.AP;
    L.oW2N = o; o += H * H;
    L.oW3N = o; o += L.YR * H;
    L.oLS = o;  o += L.AP;
    L.total = round_up(o, 4);
    int t = 0;
    L.tW1 = t; t += h1 * K0;
    L.tb1 = t; t += h1;
    L.tW2 = t; t += h2 * h1;
    L.tb2 = t; t += h2;
    L.tW3 = t; t += A * h2;
    L.tb3 = t; t += A;
    L.tLS = t; if (has_log_std) t += A;
    L.d = t;
    return L;
}

}  // namespace mjb

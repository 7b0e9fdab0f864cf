// Developer probe: validates the hand-written tcgen05 plumbing (TMEM alloc, no-swizzle canonical smem layouts,
// K-major and MN-major descriptors, instruction descriptor, commit/mbarrier, tcgen05.ld) on one 128x128xK tile.
//   D[m][n] = sum_k A[m][k] * B[n][k]          fp16 inputs, fp32 accumulate in TMEM
// mode bit0: A is read MN-major from a buffer stored [k][m];  bit1: B is read MN-major from a buffer stored [k][n].
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o tools/libtcprobe.so tools/tc_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../mjrl_b200/csrc/tc_common.cuh"

using namespace mjb::tc;

// A_g: [128][K] row-major halves (or [K][128] when bit0), B_g likewise, D_g [128][128] floats
__global__ void __launch_bounds__(128, 1) tc_probe_kernel(const __half* A_g, const __half* B_g, float* D_g, int K, int mode) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    const bool a_mn = mode & 1, b_mn = mode & 2;
    // operand buffers, core-tiled: rows x cols halves, offset = (r/8)*128 + (c/8)*(16*rows) + (r%8)*16 + (c%8)*2
    const int a_rows = a_mn ? K : 128, a_cols = a_mn ? 128 : K;
    const int b_rows = b_mn ? K : 128, b_cols = b_mn ? 128 : K;
    unsigned char* sA = smem;
    unsigned char* sB = smem + (size_t)a_rows * a_cols * 2;
    for (int i = tid; i < a_rows * a_cols; i += 128) {
        const int r = i / a_cols, c = i % a_cols;
        *reinterpret_cast<__half*>(sA + core_offset(r, c, a_rows)) = A_g[i];
    }
    for (int i = tid; i < b_rows * b_cols; i += 128) {
        const int r = i / b_cols, c = i % b_cols;
        *reinterpret_cast<__half*>(sB + core_offset(r, c, b_rows)) = B_g[i];
    }
    if (warp == 0) tmem_alloc(&s_tmem, 128);
    if (tid == 0) mbar_init(&s_bar, 1);
    fence_proxy_async();                       // generic-proxy smem writes -> visible to the tensor core (async proxy)
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = s_tmem;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_f16(128, 128, a_mn, b_mn);
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        for (int k = 0; k < K / 16; ++k) {
            // K-major: 16 k's = 2 cores along the column-group axis; MN-major: 2 cores along the row-group axis
            const uint64_t da = a_mn ? make_desc(a0 + k * 2 * 128, /*lbo=*/128, /*sbo=*/16 * a_rows)
                                     : make_desc(a0 + k * 2 * 16 * a_rows, /*lbo=*/16 * a_rows, /*sbo=*/128);
            const uint64_t db = b_mn ? make_desc(b0 + k * 2 * 128, 128, 16 * b_rows)
                                     : make_desc(b0 + k * 2 * 16 * b_rows, 16 * b_rows, 128);
            mma_f16(tmem, da, db, idesc, k > 0);
        }
        mma_commit(&s_bar);
    }
    mbar_wait(&s_bar, 0);
    tcgen05_fence_after();
    // warp w owns TMEM lanes 32w..32w+31; each thread reads its row, 32 columns at a time
    for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) D_g[(size_t)tid * 128 + c0 + j] = __uint_as_float(v[j]);
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

extern "C" int tc_probe(const void* A_h, const void* B_h, float* D_h, int K, int mode) {
    __half *A, *B; float* D;
    cudaMalloc(&A, 128 * K * 2); cudaMalloc(&B, 128 * K * 2); cudaMalloc(&D, 128 * 128 * 4);
    cudaMemcpy(A, A_h, 128 * K * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(B, B_h, 128 * K * 2, cudaMemcpyHostToDevice);
    const size_t smem = 2 * (size_t)128 * K * 2;
    cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tc_probe_kernel<<<1, 128, smem>>>(A, B, D, K, mode);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D_h, D, 128 * 128 * 4, cudaMemcpyDeviceToHost);
    cudaFree(A); cudaFree(B); cudaFree(D);
    return (int)e;
}

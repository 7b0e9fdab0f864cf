// Developer probe: issue rate of tcgen05.mma kind::f16 (M = 128, K = 16 per instruction) as a function of the operand
// layout in shared memory (no-swizzle "core-tiled" vs SWIZZLE_128B), the N extent, the operand major-ness and the A
// source (shared memory vs TMEM).  64 accumulating MMAs per measurement, one issuing thread, clock64 around issue + wait.
// Rates only: the operand contents are arbitrary (finite) numbers and the result is not checked here.
// Build + run (GPU box): nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/mma_rate_probe tools/mma_rate_probe.cu && /tmp/mma_rate_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../mjrl_b200/csrc/tc_common.cuh"

using namespace mjb::tc;

__device__ __forceinline__ uint64_t make_desc_sw(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return make_desc(saddr, lbo, sbo) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}

struct Case { int N; int layout; int a_mn; int b_mn; int a_tmem; int two_acc; };   // layout: 0 none, 2 SWIZZLE_128B

__global__ void __launch_bounds__(128, 1) probe(const Case* cases, int n_cases, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (16384 + 32768) / 2; i += 128) reinterpret_cast<__half*>(smem)[i] = __float2half(((i * 37) % 101) * 0.01f - 0.5f);
    if (warp == 0) tmem_alloc(&s_tmem, 512);
    if (tid == 0) mbar_init(&s_bar, 1);
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = s_tmem, a0 = smem_u32(smem), b0 = a0 + 16384;
    uint32_t par = 0;
    for (int c = 0; c < n_cases; ++c) {
        const Case cs = cases[c];
        if (tid == 0) {
            const uint32_t idesc = make_idesc_f16(128, cs.N, cs.a_mn, cs.b_mn);
            // 64-wide K tile (4 k-steps).  no-swizzle K-major: column groups 16*rows apart; MN-major: 8-row groups 128 apart
            // SWIZZLE_128B K-major: rows of 128 B, 8-row atoms of 1024 B, k-step = +32 B inside the row
            uint64_t da[4], db[4];
            for (int k = 0; k < 4; ++k) {
                if (cs.layout == 0) {
                    da[k] = cs.a_mn ? make_desc(a0 + k * 2 * 128, 128, 16 * 64) : make_desc(a0 + k * 2 * 16 * 128, 16 * 128, 128);
                    db[k] = cs.b_mn ? make_desc(b0 + k * 2 * 128, 128, 16 * 64) : make_desc(b0 + k * 2 * 16 * cs.N, 16 * cs.N, 128);
                } else {
                    da[k] = make_desc_sw(a0 + k * 32, 16, 1024, cs.layout);
                    db[k] = make_desc_sw(b0 + k * 32, 16, 1024, cs.layout);
                }
            }
            const long long t0 = clock64();
            for (int r = 0; r < 16; ++r)
                for (int k = 0; k < 4; ++k) {
                    const uint32_t d = tmem + ((cs.two_acc && (k & 1)) ? 256 : 0);
                    const bool acc = cs.two_acc ? (r > 0 || k > 1) : (r | k) > 0;
                    if (cs.a_tmem) mma_f16_ts(d, tmem + 480 - 8 * k, db[k], idesc, acc);
                    else mma_f16(d, da[k], db[k], idesc, acc);
                }
            const long long t1 = clock64();
            mma_commit(&s_bar);
            mbar_wait(&s_bar, par);
            const long long t2 = clock64();
            out[2 * c] = t1 - t0;
            out[2 * c + 1] = t2 - t0;
        }
        par ^= 1;
        __syncthreads();
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// Same 64 MMAs, issued by `issuers` threads (lane 0 of warps 0 .. issuers-1), each into its own accumulator: is the ~92-cycle
// per-instruction floor a property of the issuing thread or of the tensor pipe?
__global__ void __launch_bounds__(128, 1) probe_multi(int N, int issuers, int converged, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar[4];
    __shared__ long long s_t[8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (16384 + 32768) / 2; i += 128) reinterpret_cast<__half*>(smem)[i] = __float2half(((i * 37) % 101) * 0.01f - 0.5f);
    if (warp == 0) tmem_alloc(&s_tmem, 512);
    if (tid < 4) mbar_init(&s_bar[tid], 1);
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = s_tmem, a0 = smem_u32(smem), b0 = a0 + 16384;
    const uint32_t idesc = make_idesc_f16(128, N, false, false);
    const int per = 64 / issuers;
    const long long t0 = clock64();
    if (warp < issuers) {
        if (converged) {                                   // whole warp runs the loop; one elected lane issues
            uint32_t pred;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
            for (int i = 0; i < per; ++i) {
                const int k = i & 3;
                const uint64_t da = make_desc(a0 + k * 2 * 16 * 128, 16 * 128, 128), db = make_desc(b0 + k * 2 * 16 * N, 16 * N, 128);
                if (pred) mma_f16(tmem + 128 * warp, da, db, idesc, i > 0);
            }
            if (pred) mma_commit(&s_bar[warp]);
        } else if (lane == 0) {
            for (int i = 0; i < per; ++i) {
                const int k = i & 3;
                const uint64_t da = make_desc(a0 + k * 2 * 16 * 128, 16 * 128, 128), db = make_desc(b0 + k * 2 * 16 * N, 16 * N, 128);
                mma_f16(tmem + 128 * warp, da, db, idesc, i > 0);
            }
            mma_commit(&s_bar[warp]);
        }
        if (lane == 0) s_t[warp] = clock64() - t0;
        mbar_wait(&s_bar[warp], 0);
        if (lane == 0) s_t[4 + warp] = clock64() - t0;
    }
    __syncthreads();
    if (tid == 0) {
        long long a = 0, b = 0;
        for (int w = 0; w < issuers; ++w) { a = s_t[w] > a ? s_t[w] : a; b = s_t[4 + w] > b ? s_t[4 + w] : b; }
        out[0] = a; out[1] = b;
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// 512-thread CTA, thread 0 issues 64 MMAs (pattern: N=128 then N=64 into the same accumulator when `mixed`), the other
// threads meanwhile: 0 = park on bar.sync, 1 = wait on the same mbarrier (try_wait loop), 2 = stream LDS.128 from another
// shared-memory region, 3 = stream tcgen05.ld from other TMEM columns, 4 = stream STS.128 into another region.
__global__ void __launch_bounds__(512, 1) probe_busy(int mode, int mixed, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ volatile int s_done;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (65536 + 32768) / 2; i += 512) reinterpret_cast<__half*>(smem)[i] = __float2half(((i * 37) % 101) * 0.01f - 0.5f);
    if (warp == 0) tmem_alloc(&s_tmem, 512);
    if (tid == 0) { mbar_init(&s_bar, 1); s_done = 0; }
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = s_tmem, a0 = smem_u32(smem), b0 = a0 + 65536;
    if (tid == 0) {
        const uint32_t id128 = make_idesc_f16(128, 128, false, false), id64 = make_idesc_f16(128, 64, false, false);
        const long long t0 = clock64();
        for (int i = 0; i < 64; ++i) {
            const int k = i & 7;                     // 8 k-steps over a 128-wide K (A 32 KB, B 32 KB: the fit's layer-2 shapes)
            const uint64_t da = make_desc(a0 + k * 2 * 16 * 128, 16 * 128, 128), db = make_desc(b0 + k * 2 * 16 * 128, 16 * 128, 128);
            mma_f16(tmem, da, db, (mixed && (i & 1)) ? id64 : id128, i > 0);
        }
        const long long t1 = clock64();
        mma_commit(&s_bar);
        mbar_wait(&s_bar, 0);
        const long long t2 = clock64();
        s_done = 1;
        out[0] = t1 - t0; out[1] = t2 - t0;
    } else if (warp == 0) {
        // (lanes 1..31 of the issuing warp: parked -- .sync.aligned instructions need the whole warp)
    } else if (mode == 1) {
        mbar_wait(&s_bar, 0);
    } else if (mode == 2) {
        uint4 acc = make_uint4(0, 0, 0, 0);
        const uint4* p = reinterpret_cast<const uint4*>(smem + 98304);
        while (!s_done) { for (int j = 0; j < 8; ++j) { const uint4 v = p[(tid + 64 * j) & 2047]; acc.x ^= v.x; acc.y ^= v.y; } }
        if (acc.x == 0x12345) out[7] = acc.y;
    } else if (mode == 3) {
        uint32_t v[16], x = 0;
        while (!s_done) { tmem_ld16(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + 256 + 16 * (warp >> 2), v); tmem_ld_wait(); x ^= v[0]; }
        if (x == 0x12345) out[7] = x;
    } else if (mode == 4) {
        uint4* p = reinterpret_cast<uint4*>(smem + 98304);
        uint32_t c = 0;
        while (!s_done) { for (int j = 0; j < 8; ++j) p[(tid + 64 * j) & 2047] = make_uint4(c, c, c, c); ++c; }
    }
    __syncthreads();
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
    Case h[] = {
        {64, 0, 0, 0, 0, 0}, {128, 0, 0, 0, 0, 0}, {256, 0, 0, 0, 0, 0},      // no swizzle, K-major
        {64, 2, 0, 0, 0, 0}, {128, 2, 0, 0, 0, 0}, {256, 2, 0, 0, 0, 0},      // SWIZZLE_128B, K-major
        {128, 0, 1, 1, 0, 0}, {128, 0, 0, 1, 0, 0}, {32, 0, 1, 1, 0, 0}, {16, 0, 1, 1, 0, 0},   // MN-major mixes (gradient GEMMs)
        {128, 0, 0, 0, 1, 0}, {128, 2, 0, 0, 1, 0}, {64, 0, 0, 0, 1, 0},      // A from TMEM
        {128, 0, 0, 0, 0, 1}, {128, 2, 0, 0, 0, 1},                          // two alternating accumulators
    };
    const int n = sizeof(h) / sizeof(h[0]);
    Case* d; long long* out;
    cudaMalloc(&d, sizeof(h)); cudaMalloc(&out, 2 * n * sizeof(long long));
    cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768);
    cudaFuncSetAttribute(probe_multi, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768);
    for (int rep = 0; rep < 2; ++rep) probe<<<1, 128, 16384 + 32768>>>(d, n, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    long long r[2 * 32];
    cudaMemcpy(r, out, 2 * n * sizeof(long long), cudaMemcpyDeviceToHost);
    printf("tcgen05.mma kind::f16 M=128 K=16, 64 accumulating MMAs per case (floor = N/2 cycles per MMA)\n");
    for (int c = 0; c < n; ++c)
        printf("N=%3d layout=%s A:%s-major%s B:%s-major%s : issue %.1f cyc/MMA, issue+drain %.1f cyc/MMA (floor %d)\n", h[c].N,
               h[c].layout ? "SW128" : "none ", h[c].a_mn ? "MN" : "K ", h[c].a_tmem ? " (TMEM)" : "", h[c].b_mn ? "MN" : "K ",
               h[c].two_acc ? " two accumulators" : "", r[2 * c] / 64.0, r[2 * c + 1] / 64.0, h[c].N / 2);
    cudaFuncSetAttribute(probe_busy, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304 + 32768);
    for (int mixed = 0; mixed < 2; ++mixed)
        for (int mode = 0; mode < 5; ++mode) {
            for (int rep = 0; rep < 2; ++rep) probe_busy<<<1, 512, 98304 + 32768>>>(mode, mixed, out);
            e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(r, out, 2 * sizeof(long long), cudaMemcpyDeviceToHost);
            const char* names[] = {"parked on bar.sync", "waiting on the same mbarrier", "streaming LDS.128", "streaming tcgen05.ld", "streaming STS.128"};
            printf("512-thread CTA, 64 MMAs %s, other 511 threads %s: issue %.1f, issue+drain %.1f cycles per MMA\n",
                   mixed ? "(N=128 / N=64 alternating)" : "(N=128)", names[mode], r[0] / 64.0, r[1] / 64.0);
            fflush(stdout);
        }
    for (int conv = 0; conv < 2; ++conv)
        for (int N = 64; N <= 128; N += 64)
            for (int iss = 1; iss <= 4; iss *= 2) {
                for (int rep = 0; rep < 2; ++rep) probe_multi<<<1, 128, 16384 + 32768>>>(N, iss, conv, out);
                e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(r, out, 2 * sizeof(long long), cudaMemcpyDeviceToHost);
                printf("N=%3d, %d issuing thread(s)%s, 64 MMAs in total: issue %.1f, issue+drain %.1f cycles per MMA\n", N, iss,
                       conv ? " (converged warp, elect.sync)" : " (lane 0 under a divergent branch)", r[0] / 64.0, r[1] / 64.0);
            }
    return 0;
}

// Fisher-vector product of the Gaussian LINEAR policy (policies/gaussian_linear.py) on tcgen05 tensor cores.
//
//   ydot = x~ V^T + c ;  dy = out_scale^2 * 2/(2 sigma^2 + 1e-8) * ydot ;  G += dy^T [x~ | 1]      (npg_cg.py:62-81)
//
// This path is HBM-bound (cfg5: 1.5 KB of observations vs 25.6 kFLOP per timestep), so the kernel is organised
// around streaming obs exactly once at full bandwidth:
//   * a tile is 64 consecutive timesteps; its fp32 rows are prefetched into registers one tile ahead (coalesced
//     64-byte row segments), then transformed, split into two fp16 terms and written into the no-swizzle canonical
//     "core-tiled" shared-memory layout (tc_common.cuh) -- conflict-free 8-byte stores;
//   * both GEMMs use the full-rate M = 128, N = 64 instruction shape by STACKING the two fp16 terms: the staged tile
//     has 128 rows [hi of the 64 samples ; lo of the 64 samples], the tangent 64 rows [V hi ; V lo], dy 64 columns
//     [dy hi | dy lo]; the cross terms land in separate accumulator rows / columns and are added in the epilogue
//     (an M = 64 SS-mode MMA measured ~90 cycles per instruction here, the stacked form ~50 for 2.7x the work);
//   * GEMM 1 reads the staged tile K-major; GEMM 2 reads THE SAME buffer MN-major (rows = features, reduction over
//     the 128 staged rows) -- the tile is neither re-read nor transposed;
//   * the [obs x act] gradient accumulators (3 x 128 x 64, incl. the bias gradient through a ones column) stay in
//     TMEM for the whole kernel; the tangent weights (48 KB) are loaded once per CTA by one TMA bulk copy.
// Two-term fp16 operands (hi*hi + lo*hi + hi*lo) keep fp32-level accuracy; the tangent is pre-scaled by 2^e.
#include <cuda_fp16.h>

#include "kernels.h"
#include "tc_common.cuh"

namespace mjb {

using namespace tc;

namespace {

constexpr int LM = 64;                    // samples per tile
constexpr int LKP = 384;                  // padded feature count (obs_dim + ones column <= 384)
constexpr int LNP = 32;                   // padded action count
constexpr int XR = 2 * LM;                // staged rows: [fp16 hi of the 64 samples ; fp16 lo of the 64 samples]
constexpr int X_LB = 16 * XR;             // bytes between 8-column groups of a staged block
constexpr int FB = 128;                   // features per staged block (= M of GEMM 2)
constexpr int NFB = LKP / FB;             // 3 blocks per tile
constexpr int XB_BYTES = XR * FB * 2;     // 32 KB per block
constexpr int RING = 3;                   // staged blocks of one tile
constexpr int VR = 2 * LNP;               // tangent rows: [V hi (32 padded actions) ; V lo]
constexpr int V_LB = 16 * VR;
constexpr int V_BYTES = VR * LKP * 2;
constexpr int DYC = 2 * LNP;              // dy columns: [dy hi | dy lo]
constexpr int DY_LB = 16 * XR;            // dy rows follow the staged rows (the reduction axis of GEMM 2)
constexpr int DY_BYTES = XR * DYC * 2;

// global prepped tangent block: [V hi ; V lo] core-tiled (rows = 64, cols = 384 features) + fp32 c[32]
constexpr int GL_V = 0, GL_C = V_BYTES, GL_TOTAL = GL_C + 32 * 4;

// shared memory map (bytes)
constexpr int SL_X = 0, SL_V = SL_X + RING * XB_BYTES, SL_DY = SL_V + V_BYTES, SL_SCR = SL_DY + DY_BYTES;
constexpr int SCR_PITCH = 33;
constexpr int SL_F32 = SL_SCR + LM * SCR_PITCH * 4;     // floats: shift[384] rinv[384] c[32] fac[32]
constexpr int SLF_SHIFT = 0, SLF_RINV = 384, SLF_C = 768, SLF_FAC = 800, SLF_END = 832;
constexpr int SL_BAR = SL_F32 + SLF_END * 4;
constexpr int SL_TOTAL = SL_BAR + 64;

constexpr int FLUSH_TILES = 24;           // tiles per TMEM accumulation group (see flush() in the kernel)
constexpr uint32_t TL_D1 = 0, TL_G = 64, TL_COLS = 256;   // TMEM columns: GEMM-1 output [0,64), G blocks 3 x 64

struct LinTcArgs {
    const unsigned char* T;      // prepped (scaled) tangent
    const float* theta;          // flat theta (for log_std)
    const float* in_shift; const float* in_scale; const float* out_scale;
    const float* obs; int K0; int A; const int* idx; long long n;
    float* gpartial; long long gstride; int tW, tb, tLS;
    unsigned long long* prof;    // developer aid: per-phase clock64 sums (nullptr = off)
};

unsigned long long* g_lin_prof = nullptr;

// streamed once: read-only path, no L1 allocation
__device__ __forceinline__ float4 ld_stream4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// One MMA shape for both GEMMs: M = 128, N = 64 (an M = 64 tile in SS mode runs ~4x slower per instruction).
//   GEMM 1:  D1[r][n] = sum_k X[r][k] V[n][k],  r = (hi|lo, sample), n = (hi|lo, action)
//            ydot[m][a] = D1[m][a] + D1[m][32+a] + D1[64+m][a]           (hi*hi + hi*lo + lo*hi)
//   GEMM 2:  G_j[f][n] += sum_r X[r][128 j + f] DY[r][n],  DY[m] = [dy hi | dy lo], DY[64+m] = [dy hi | 0]
//            G[f][a] = G_j[f][a] + G_j[f][32+a]                           (hi*hi + lo*hi + hi*lo)
// Per tile: convert the prefetched registers -> GEMM 1 (async) while the next tile's global loads are issued ->
// dy epilogue on all 16 warps -> GEMM 2 (async).  The next tile's loads are in flight under both GEMMs.
// Measured alternatives (tools/lin_fvp_profile.py): a 4-slot block ring that converts tile t+1 under
// the MMAs of tile t was slower -- ptxas puts every LDG on one scoreboard, so a conversion waits for the most recent
// loads, and ~96 KB of register-prefetched loads per SM exceed the LSU's outstanding-request capacity (lg_throttle).
template <bool IDENT>
__global__ void __launch_bounds__(512, 1) linear_tc_kernel(const LinTcArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sf = reinterpret_cast<float*>(smem + SL_F32);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SL_BAR);     // [0] GEMM 1, [1..3] GEMM 2 block j, [4] tangent
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K0 = a.K0, A = a.A;

    if (warp == 0) tmem_alloc(&s_tmem, TL_COLS);
    if (tid == 0) { for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1); }
    for (int i = tid; i < DY_BYTES / 16; i += 512) reinterpret_cast<uint4*>(smem + SL_DY)[i] = make_uint4(0, 0, 0, 0);
    {   // this CTA's gradient partial starts at zero (no memset node in front of the kernel); gstride is a multiple of 32
        float4* gz = reinterpret_cast<float4*>(a.gpartial + (size_t)blockIdx.x * a.gstride);
        for (int i = tid; i < (int)(a.gstride / 4); i += 512) gz[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int k = tid; k < LKP; k += 512) {
        sf[SLF_SHIFT + k] = k < K0 ? a.in_shift[k] : 0.0f;
        sf[SLF_RINV + k] = k < K0 ? 1.0f / (a.in_scale[k] + 1e-8f) : 0.0f;
    }
    if (tid < 32) {
        sf[SLF_C + tid] = reinterpret_cast<const float*>(a.T + GL_C)[tid];
        float f = 0.0f;
        if (tid < A) {
            const float sd = expf(a.theta[a.tLS + tid]);
            const float os = a.out_scale[tid];
            f = os * os * (2.0f / (2.0f * sd * sd + 1e-8f));
        }
        sf[SLF_FAC + tid] = f;
    }
    __syncthreads();
    if (tid == 0) {                                                  // tangent weights: one TMA bulk copy
        mbar_expect_tx(&bars[4], V_BYTES);
        bulk_g2s(smem + SL_V, a.T + GL_V, V_BYTES, &bars[4]);
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = s_tmem, sbase = smem_u32(smem);
    const uint32_t ID1 = make_idesc_f16(128, DYC, false, false);     // A = tile (K-major), B = V (K-major)
    const uint32_t ID2 = make_idesc_f16(128, DYC, true, true);       // A = tile (MN-major), B = dy (MN-major)
    uint32_t mma_par = 0;
    auto all_wait_mma = [&]() { mbar_wait(&bars[0], mma_par); mma_par ^= 1; tcgen05_fence_after(); };
    mbar_wait(&bars[4], 0);

    // staging map: warp-iteration = 8 rows x 16 features; lane -> (row r = lane%8, float4 column cidx = lane/8)
    // 64 x 384 tile = 8 row groups x 24 column blocks = 192 warp-iterations, 12 per warp (4 per feature block)
    const int r8 = lane & 7, cidx = lane >> 3;
    const int m_st = 8 * (warp & 7) + r8, k_st = 16 * (warp >> 3) + 4 * cidx;     // iteration u adds 32 features
    const uint32_t st_off = core_offset(m_st, k_st, XR);             // within a block; iteration u adds 4 column groups
    const int one_pos = K0 - k_st;                                   // ones column: iteration one_pos / 32, element one_pos % 32
    const bool vec = (K0 & 3) == 0;
    float4 pre[12];
    auto issue_loads = [&](long long tile) {
        const long long row = tile * LM + m_st;
        const bool rv = row < a.n;
        const long long rr = rv ? (a.idx ? (long long)a.idx[row] : row) : 0;
        const float* p0 = a.obs + rr * K0 + k_st;
#pragma unroll
        for (int u = 0; u < 12; ++u) {
            const int k = k_st + 32 * u;
            const float* p = p0 + 32 * u;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rv && k < K0) {
                if (vec) v = ld_stream4(p);
                else { v.x = p[0]; if (k + 1 < K0) v.y = p[1]; if (k + 2 < K0) v.z = p[2]; if (k + 3 < K0) v.w = p[3]; }
            }
            pre[u] = v;
        }
    };
    // converts the prefetched tile into the three staged blocks (padding columns: zero, ones column: one)
    auto stage = [&]() {
#pragma unroll
        for (int u = 0; u < 12; ++u) {
            float x0 = pre[u].x, x1 = pre[u].y, x2 = pre[u].z, x3 = pre[u].w;
            if (!IDENT) {
                const int k = k_st + 32 * u;
                const float4 sh = *reinterpret_cast<const float4*>(sf + SLF_SHIFT + k);
                const float4 ri = *reinterpret_cast<const float4*>(sf + SLF_RINV + k);
                x0 = (x0 - sh.x) * ri.x; x1 = (x1 - sh.y) * ri.y; x2 = (x2 - sh.z) * ri.z; x3 = (x3 - sh.w) * ri.w;
            }
            if ((one_pos >> 5) == u && one_pos >= 0) {
                const int e = one_pos & 31;
                if (e == 0) x0 = 1.0f; else if (e == 1) x1 = 1.0f; else if (e == 2) x2 = 1.0f; else if (e == 3) x3 = 1.0f;
            }
            const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
            const float2 b01 = __half22float2(h01), b23 = __half22float2(h23);
            const __half2 l01 = __floats2half2_rn(x0 - b01.x, x1 - b01.y), l23 = __floats2half2_rn(x2 - b23.x, x3 - b23.y);
            uint2 hv, lv;
            hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
            lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
            unsigned char* o = smem + SL_X + (u >> 2) * XB_BYTES + st_off + (u & 3) * 4 * X_LB;
            *reinterpret_cast<uint2*>(o) = hv;
            *reinterpret_cast<uint2*>(o + LM * 16) = lv;                 // row 64 + m
        }
    };

    // Adds the TMEM gradient accumulators to the per-CTA partial in global memory (zeroed before the launch, owned by this
    // CTA): lane = feature, columns = action.  Called every FLUSH_TILES tiles because the tensor core's fp32 accumulation
    // truncates -- an accumulator that lives for the whole kernel drifts by ~1.5e-7 (relative) per tile.
    auto flush = [&]() {
        float* gp = a.gpartial + (size_t)blockIdx.x * a.gstride;
        if (warp < 4) {
            const int kl = 32 * warp + lane;
#pragma unroll 1
            for (int j = 0; j < NFB; ++j) {
                const int k = FB * j + kl;
#pragma unroll 1
                for (int c8 = 0; 8 * c8 < A; ++c8) {                     // 8 actions at a time: keeps the register footprint small
                    uint32_t g0[8], g1[8];
                    tmem_ld8(tmem + ((uint32_t)(32 * warp) << 16) + TL_G + DYC * j + 8 * c8, g0);
                    tmem_ld8(tmem + ((uint32_t)(32 * warp) << 16) + TL_G + DYC * j + 32 + 8 * c8, g1);
                    tmem_ld_wait();
                    // red.global.add: fire-and-forget (no read round trip); the slice belongs to this CTA
#pragma unroll
                    for (int o8 = 0; o8 < 8; ++o8) {
                        const int o = 8 * c8 + o8;
                        const float gv = __uint_as_float(g0[o8]) + __uint_as_float(g1[o8]);
                        if (o < A) {
                            if (k < K0) atomicAdd(&gp[a.tW + o * K0 + k], gv);
                            else if (k == K0) atomicAdd(&gp[a.tb + o], gv);
                        }
                    }
                }
            }
        }
        tcgen05_fence_before();
    };
    const long long n_tiles = (a.n + LM - 1) / LM;
    const long long G = gridDim.x;
    long long it = 0;
    long long t_last = clock64();
    unsigned long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define LIN_PROF(i) do { if (a.prof && tid == 0) { const long long _t = clock64(); pacc[i] += _t - t_last; t_last = _t; } } while (0)
    if ((long long)blockIdx.x < n_tiles) issue_loads(blockIdx.x);
    for (long long tile = blockIdx.x; tile < n_tiles; tile += G, ++it) {
        const long long base = tile * LM;
        const bool first = (it % FLUSH_TILES == 0);                  // first tile of an accumulation group
        stage();                                                     // consumes the prefetched registers
        LIN_PROF(0);
        fence_proxy_async();
        tcgen05_fence_before();
        __syncthreads();
        // ---- GEMM 1 ----
        if (tid == 0) {
            tcgen05_fence_after();
#pragma unroll
            for (int i = 0; i < NFB; ++i) {
                const uint32_t xb = sbase + SL_X + (uint32_t)i * XB_BYTES;
#pragma unroll
                for (int jj = 0; jj < FB / 16; ++jj)
                    mma_f16(tmem + TL_D1, make_desc(xb + 2 * jj * X_LB, X_LB, 128),
                            make_desc(sbase + SL_V + 2 * (8 * i + jj) * V_LB, V_LB, 128), ID1, (i | jj) > 0);
            }
            mma_commit(&bars[0]);
        }
        LIN_PROF(1);
        if (tile + G < n_tiles) issue_loads(tile + G);               // next tile: in flight under the MMAs
        all_wait_mma();
        LIN_PROF(2);
        // ---- dy = fac * (ydot + c), masked.  TMEM lanes 0..63: hi rows of the samples, 64..127: lo rows.
        //      16 warps: lane quadrant q = warp % 4, action group cq = warp / 4 (8 actions each) ----
        {
            const int q = warp & 3, cq = warp >> 2;
            float* scr = reinterpret_cast<float*>(smem + SL_SCR);
            uint32_t y0[8], y1[8];
            const uint32_t tl = tmem + ((uint32_t)(32 * q) << 16) + TL_D1 + 8 * cq;
            tmem_ld8(tl, y0);
            if (q < 2) tmem_ld8(tl + 32, y1);
            tmem_ld_wait();
            if (q >= 2) {
                float* sp = scr + (32 * (q - 2) + lane) * SCR_PITCH + 8 * cq;
#pragma unroll
                for (int o = 0; o < 8; ++o) sp[o] = __uint_as_float(y0[o]);
            }
            __syncthreads();
            if (q < 2) {
                const int m = 32 * q + lane;
                const bool valid = (base + m) < a.n;
                const float* sp = scr + m * SCR_PITCH + 8 * cq;
                __half2 h[4], l[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int o = 2 * j, a0 = 8 * cq + o;
                    const float s0f = __uint_as_float(y0[o]) + __uint_as_float(y1[o]) + sp[o];
                    const float s1f = __uint_as_float(y0[o + 1]) + __uint_as_float(y1[o + 1]) + sp[o + 1];
                    const float d0 = valid ? sf[SLF_FAC + a0] * (s0f + sf[SLF_C + a0]) : 0.0f;
                    const float d1 = valid ? sf[SLF_FAC + a0 + 1] * (s1f + sf[SLF_C + a0 + 1]) : 0.0f;
                    h[j] = __floats2half2_rn(d0, d1);
                    const float2 bk = __half22float2(h[j]);
                    l[j] = __floats2half2_rn(d0 - bk.x, d1 - bk.y);
                }
                const uint32_t o = (uint32_t)cq * DY_LB + (uint32_t)((m >> 3) * 128 + (m & 7) * 16);
                *reinterpret_cast<uint4*>(smem + SL_DY + o) = *reinterpret_cast<const uint4*>(h);                 // row m, hi
                *reinterpret_cast<uint4*>(smem + SL_DY + o + 4 * DY_LB) = *reinterpret_cast<const uint4*>(l);     // row m, lo
                *reinterpret_cast<uint4*>(smem + SL_DY + o + LM * 16) = *reinterpret_cast<const uint4*>(h);       // row 64+m, hi
            }
        }
        fence_proxy_async();
        tcgen05_fence_before();
        __syncthreads();
        LIN_PROF(3);
        // ---- GEMM 2 (reduction over the 128 staged rows) ----
        if (tid == 0) {
            tcgen05_fence_after();
#pragma unroll
            for (int j = 0; j < NFB; ++j) {
                const uint32_t xb = sbase + SL_X + (uint32_t)j * XB_BYTES;
#pragma unroll
                for (int ks = 0; ks < XR / 16; ++ks)
                    mma_f16(tmem + TL_G + DYC * j, make_desc(xb + 2 * ks * 128, 128, X_LB),
                            make_desc(sbase + SL_DY + 2 * ks * 128, 128, DY_LB), ID2, !first || ks > 0);
            }
            mma_commit(&bars[0]);
        }
        LIN_PROF(4);
        all_wait_mma();                                              // the tile buffer is restaged next
        LIN_PROF(5);
        if ((it + 1) % FLUSH_TILES == 0) flush();
    }
    if (a.prof && tid == 0) {
        for (int i = 0; i < 6; ++i) atomicAdd(a.prof + i, pacc[i]);
        atomicAdd(a.prof + 6, (unsigned long long)it);
    }

    if (it > 0 && it % FLUSH_TILES != 0) flush();
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, TL_COLS);
}

// tangent (flat theta layout: W [A][K0], b [A], log_std [A]) -> fp16 [hi ; lo] core-tiled [64 x 384] + fp32 c[32]
__global__ void lin_tc_prep_kernel(const float* __restrict__ v, int K0, int A, const float* __restrict__ scale_dev,
                                   unsigned char* __restrict__ out) {
    const float sc = scale_dev ? *scale_dev : 1.0f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < LNP * LKP + 32; i += gridDim.x * blockDim.x) {
        if (i < LNP * LKP) {
            const int o = i / LKP, k = i % LKP;
            const float val = (o < A && k < K0) ? v[o * K0 + k] * sc : 0.0f;
            const __half h = __float2half_rn(val), l = __float2half_rn(val - __half2float(h));
            *reinterpret_cast<__half*>(out + GL_V + core_offset(o, k, VR)) = h;
            *reinterpret_cast<__half*>(out + GL_V + core_offset(LNP + o, k, VR)) = l;
        } else {
            const int o = i - LNP * LKP;
            reinterpret_cast<float*>(out + GL_C)[o] = o < A ? v[A * K0 + o] * sc : 0.0f;
        }
    }
}

}  // namespace

void lin_tc_set_prof(unsigned long long* p) { g_lin_prof = p; }

size_t lin_tc_prep_bytes() { return (size_t)round_up(GL_TOTAL, 256); }
bool lin_tc_supported(int K0, int A) { return K0 < LKP && A <= LNP && K0 >= 1; }

void launch_lin_tc_prep(const float* v, int K0, int A, const float* scale_dev, unsigned char* out, cudaStream_t s) {
    lin_tc_prep_kernel<<<48, 256, 0, s>>>(v, K0, A, scale_dev, out);
}

cudaError_t launch_linear_tc(const unsigned char* T, const float* theta, const float* in_shift, const float* in_scale,
                             const float* out_scale, bool identity_in, const float* obs, int K0, int A, const int* idx,
                             long long n, float* gpartial, long long gstride, int tW, int tb, int tLS, int grid, cudaStream_t s) {
    LinTcArgs a;
    a.T = T; a.theta = theta; a.in_shift = in_shift; a.in_scale = in_scale; a.out_scale = out_scale; a.obs = obs;
    a.K0 = K0; a.A = A; a.idx = idx; a.n = n; a.gpartial = gpartial; a.gstride = gstride; a.tW = tW; a.tb = tb; a.tLS = tLS;
    a.prof = g_lin_prof;
    auto kern = identity_in ? linear_tc_kernel<true> : linear_tc_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SL_TOTAL);
    if (e != cudaSuccess) return e;
    kern<<<grid, 512, SL_TOTAL, s>>>(a);
    return cudaGetLastError();
}

}  // namespace mjb

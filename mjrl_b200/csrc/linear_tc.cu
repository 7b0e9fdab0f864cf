// Fisher-vector product of the Gaussian LINEAR policy (policies/gaussian_linear.py) on tcgen05 tensor cores.
//
//   ydot = x~ V^T + c ;  dy = out_scale^2 * 2/(2 sigma^2 + 1e-8) * ydot ;  G += dy^T [x~ | 1]      (npg_cg.py:62-81)
//
// This path is HBM-bound (cfg5: 1.5 KB of observations vs 25.6 kFLOP per timestep), so the kernel is organised
// around streaming obs exactly once at full bandwidth:
//   * a tile is 64 consecutive timesteps; its fp32 rows are prefetched into registers one tile ahead (coalesced
//     64-byte row segments), then transformed, split into two fp16 terms and written into the no-swizzle canonical
//     "core-tiled" shared-memory layout (tc_common.cuh) -- conflict-free 8-byte stores;
//   * GEMM 1 (M=64 samples, N=32 padded actions, K=obs) reads the staged tile K-major; GEMM 2 reads THE SAME buffer
//     MN-major (rows = features, reduction over the 64 samples) -- the tile is neither re-read nor transposed;
//   * the [obs x act] gradient accumulators (3 x 128 x 32, incl. the bias gradient through a ones column) stay in
//     TMEM for the whole kernel; the tangent weights (fp16 hi/lo, 48 KB) are loaded once per CTA by TMA bulk copy.
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
constexpr int X_LB = 16 * LM;             // bytes between 8-column groups of the staged tile (rows = 64)
constexpr int V_LB = 16 * LNP;            // ... of the tangent weights (rows = 32)
constexpr int X_BYTES = LM * LKP * 2;     // one fp16 term of the tile
constexpr int V_BYTES = LNP * LKP * 2;
constexpr int DY_BYTES = LM * LNP * 2;

// global prepped tangent block: [V hi][V lo] core-tiled (rows = 32 actions, cols = 384 features) + fp32 c[32]
constexpr int GL_V = 0, GL_C = 2 * V_BYTES, GL_TOTAL = GL_C + 32 * 4;

// shared memory map (bytes)
constexpr int SL_XHI = 0, SL_XLO = SL_XHI + X_BYTES, SL_VHI = SL_XLO + X_BYTES, SL_VLO = SL_VHI + V_BYTES;
constexpr int SL_DYHI = SL_VLO + V_BYTES, SL_DYLO = SL_DYHI + DY_BYTES;
constexpr int SL_F32 = SL_DYLO + DY_BYTES;     // floats: shift[384] rinv[384] c[32] fac[32] gb[4][32]
constexpr int SLF_SHIFT = 0, SLF_RINV = 384, SLF_C = 768, SLF_FAC = 800, SLF_END = 832;
constexpr int SL_BAR = SL_F32 + SLF_END * 4;
constexpr int SL_TOTAL = SL_BAR + 32;

constexpr uint32_t TL_DY = 0, TL_G = 32;       // TMEM columns: ydot [0,32), G chunks [32,128)

struct LinTcArgs {
    const unsigned char* T;      // prepped (scaled) tangent
    const float* theta;          // flat theta (for log_std)
    const float* in_shift; const float* in_scale; const float* out_scale;
    const float* obs; int K0; int A; const int* idx; long long n;
    float* gpartial; long long gstride; int tW, tb, tLS;
};

__global__ void __launch_bounds__(512, 1) linear_tc_kernel(const LinTcArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sf = reinterpret_cast<float*>(smem + SL_F32);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SL_BAR);     // [0] mma done, [1] tangent weights landed
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K0 = a.K0, A = a.A;

    if (warp == 0) tmem_alloc(&s_tmem, 128);
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); }
    for (int i = tid; i < (SL_VHI - SL_XHI) / 16; i += 512) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < 2 * DY_BYTES / 16; i += 512) reinterpret_cast<uint4*>(smem + SL_DYHI)[i] = make_uint4(0, 0, 0, 0);
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
    if (tid < LM) *reinterpret_cast<__half*>(smem + SL_XHI + core_offset(tid, K0, LM)) = __float2half_rn(1.0f);   // ones column
    if (tid == 0) {                                                  // tangent weights: one TMA bulk copy per term
        mbar_expect_tx(&bars[1], 2 * V_BYTES);
        bulk_g2s(smem + SL_VHI, a.T + GL_V, V_BYTES, &bars[1]);
        bulk_g2s(smem + SL_VLO, a.T + GL_V + V_BYTES, V_BYTES, &bars[1]);
    }
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = s_tmem, sbase = smem_u32(smem);
    const uint32_t ID1 = make_idesc_f16(64, LNP, false, false);      // ydot: A = tile (K-major), B = V (K-major)
    const uint32_t ID2 = make_idesc_f16(128, LNP, true, true);       // G: A = tile (MN-major), B = dy (MN-major)
    uint32_t mma_par = 0;
    auto all_wait_mma = [&]() { mbar_wait(&bars[0], mma_par); mma_par ^= 1; tcgen05_fence_after(); };
    mbar_wait(&bars[1], 0);

    // staging map: warp-iteration = 8 rows x 16 features; lane -> (row r = lane%8, float4 column cidx = lane/8)
    // 64 x 384 tile = 8 row groups x 24 column blocks = 192 warp-iterations, 12 per warp
    const int r8 = lane & 7, cidx = lane >> 3;
    float4 pre[12];
    auto issue_loads = [&](long long base) {
#pragma unroll
        for (int u = 0; u < 12; ++u) {
            const int wi = warp + 16 * u, rg = wi & 7, cb = wi >> 3;
            const int m = 8 * rg + r8, k = 16 * cb + 4 * cidx;
            const long long row = base + m;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < a.n && k < K0) {
                const long long rr = a.idx ? (long long)a.idx[row] : row;
                const float* p = a.obs + rr * K0 + k;
                if (k + 3 < K0 && ((K0 & 3) == 0)) v = *reinterpret_cast<const float4*>(p);
                else { v.x = p[0]; if (k + 1 < K0) v.y = p[1]; if (k + 2 < K0) v.z = p[2]; if (k + 3 < K0) v.w = p[3]; }
            }
            pre[u] = v;
        }
    };
    auto stage = [&]() {
#pragma unroll
        for (int u = 0; u < 12; ++u) {
            const int wi = warp + 16 * u, rg = wi & 7, cb = wi >> 3;
            const int m = 8 * rg + r8, k = 16 * cb + 4 * cidx;
            if (k >= K0) continue;                                  // padding columns stay zero / one
            const float4 sh = *reinterpret_cast<const float4*>(sf + SLF_SHIFT + k);
            const float4 ri = *reinterpret_cast<const float4*>(sf + SLF_RINV + k);
            const float x0 = (pre[u].x - sh.x) * ri.x, x1 = (pre[u].y - sh.y) * ri.y;
            const float x2 = (pre[u].z - sh.z) * ri.z, x3 = (pre[u].w - sh.w) * ri.w;
            __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
            const float2 b01 = __half22float2(h01), b23 = __half22float2(h23);
            __half2 l01 = __floats2half2_rn(x0 - b01.x, x1 - b01.y), l23 = __floats2half2_rn(x2 - b23.x, x3 - b23.y);
            if (k + 3 >= K0) {                                       // keep the ones column / zero padding intact
                __half* hp = reinterpret_cast<__half*>(&h01); __half* lp = reinterpret_cast<__half*>(&l01);
                __half* hq = reinterpret_cast<__half*>(&h23); __half* lq = reinterpret_cast<__half*>(&l23);
                if (k + 1 >= K0) { hp[1] = __float2half_rn(k + 1 == K0 ? 1.0f : 0.0f); lp[1] = __float2half_rn(0.0f); }
                if (k + 2 >= K0) { hq[0] = __float2half_rn(k + 2 == K0 ? 1.0f : 0.0f); lq[0] = __float2half_rn(0.0f); }
                if (k + 3 >= K0) { hq[1] = __float2half_rn(k + 3 == K0 ? 1.0f : 0.0f); lq[1] = __float2half_rn(0.0f); }
            }
            const uint32_t o = core_offset(m, k, LM);
            uint2 hv, lv;
            hv.x = *reinterpret_cast<uint32_t*>(&h01); hv.y = *reinterpret_cast<uint32_t*>(&h23);
            lv.x = *reinterpret_cast<uint32_t*>(&l01); lv.y = *reinterpret_cast<uint32_t*>(&l23);
            *reinterpret_cast<uint2*>(smem + SL_XHI + o) = hv;
            *reinterpret_cast<uint2*>(smem + SL_XLO + o) = lv;
        }
    };

    const long long n_tiles = (a.n + LM - 1) / LM;
    long long it = 0;
    if ((long long)blockIdx.x < n_tiles) issue_loads((long long)blockIdx.x * LM);
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const long long base = tile * LM;
        stage();                                                     // consumes the prefetched registers
        fence_proxy_async();
        tcgen05_fence_before();
        __syncthreads();
        if (tile + gridDim.x < n_tiles) issue_loads((tile + gridDim.x) * LM);   // next tile: in flight under the MMAs
        // ---- GEMM 1: ydot[m][a] = sum_k x~[m][k] V[a][k] ----
        if (tid == 0) {
            tcgen05_fence_after();
#pragma unroll 4
            for (int j = 0; j < LKP / 16; ++j) {
                const uint64_t ah = make_desc(sbase + SL_XHI + 2 * j * X_LB, X_LB, 128), al = make_desc(sbase + SL_XLO + 2 * j * X_LB, X_LB, 128);
                const uint64_t bh = make_desc(sbase + SL_VHI + 2 * j * V_LB, V_LB, 128), bl = make_desc(sbase + SL_VLO + 2 * j * V_LB, V_LB, 128);
                mma_f16(tmem + TL_DY, ah, bh, ID1, j > 0);
                mma_f16(tmem + TL_DY, al, bh, ID1, true);
                mma_f16(tmem + TL_DY, ah, bl, ID1, true);
            }
            mma_commit(&bars[0]);
        }
        all_wait_mma();
        // ---- dy = fac * (ydot + c), masked; M=64 accumulators sit in lanes 0..15 of each TMEM quadrant ----
        if (warp < 4) {                                              // warp-uniform: the TMEM load is .sync.aligned
            const int m = 16 * warp + (lane & 15);
            uint32_t y[32];
            tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + TL_DY, y);
            tmem_ld_wait();
            if (lane < 16) {
            const bool valid = (base + m) < a.n;
            const int rowoff = (m >> 3) * 128 + (m & 7) * 16;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                __half2 h[4], l[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int a0 = 8 * g + 2 * j;
                    const float d0 = valid ? sf[SLF_FAC + a0] * (__uint_as_float(y[a0]) + sf[SLF_C + a0]) : 0.0f;
                    const float d1 = valid ? sf[SLF_FAC + a0 + 1] * (__uint_as_float(y[a0 + 1]) + sf[SLF_C + a0 + 1]) : 0.0f;
                    h[j] = __floats2half2_rn(d0, d1);
                    const float2 bk = __half22float2(h[j]);
                    l[j] = __floats2half2_rn(d0 - bk.x, d1 - bk.y);
                }
                const int o = g * X_LB + rowoff;
                *reinterpret_cast<uint4*>(smem + SL_DYHI + o) = *reinterpret_cast<const uint4*>(h);
                *reinterpret_cast<uint4*>(smem + SL_DYLO + o) = *reinterpret_cast<const uint4*>(l);
            }
            }
        }
        fence_proxy_async();
        tcgen05_fence_before();
        __syncthreads();
        // ---- GEMM 2: G_j[k][a] += sum_m x~[m][128 j + k] dy[m][a]   (j = 0..2; reduction over the 64 samples) ----
        if (tid == 0) {
            tcgen05_fence_after();
#pragma unroll
            for (int j = 0; j < 3; ++j)
#pragma unroll
                for (int ks = 0; ks < LM / 16; ++ks) {
                    const uint32_t xo = 16 * j * X_LB + 2 * ks * 128;        // feature block 128 j, samples 16 ks..
                    const uint64_t ah = make_desc(sbase + SL_XHI + xo, 128, X_LB), al = make_desc(sbase + SL_XLO + xo, 128, X_LB);
                    const uint64_t bh = make_desc(sbase + SL_DYHI + 2 * ks * 128, 128, X_LB), bl = make_desc(sbase + SL_DYLO + 2 * ks * 128, 128, X_LB);
                    const bool acc = (it > 0) || ks > 0;
                    mma_f16(tmem + TL_G + 32 * j, ah, bh, ID2, acc);
                    mma_f16(tmem + TL_G + 32 * j, al, bh, ID2, true);
                    mma_f16(tmem + TL_G + 32 * j, ah, bl, ID2, true);
                }
            mma_commit(&bars[0]);
        }
        all_wait_mma();                                              // the tile buffer is restaged next
    }

    if (it > 0) {                                                    // write the per-CTA partial: lane = feature, cols = action
        float* gp = a.gpartial + (size_t)blockIdx.x * a.gstride;
        if (warp < 4) {
            const int kl = 32 * warp + lane;
            for (int j = 0; j < 3; ++j) {
                uint32_t g[32];
                tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + TL_G + 32 * j, g);
                tmem_ld_wait();
                const int k = 128 * j + kl;
                for (int o = 0; o < A; ++o) {
                    if (k < K0) gp[a.tW + o * K0 + k] = __uint_as_float(g[o]);
                    else if (k == K0) gp[a.tb + o] = __uint_as_float(g[o]);
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

// tangent (flat theta layout: W [A][K0], b [A], log_std [A]) -> fp16 hi/lo core-tiled [32 x 384] + fp32 c[32]
__global__ void lin_tc_prep_kernel(const float* __restrict__ v, int K0, int A, const float* __restrict__ scale_dev,
                                   unsigned char* __restrict__ out) {
    const float sc = scale_dev ? *scale_dev : 1.0f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < LNP * LKP + 32; i += gridDim.x * blockDim.x) {
        if (i < LNP * LKP) {
            const int o = i / LKP, k = i % LKP;
            const float val = (o < A && k < K0) ? v[o * K0 + k] * sc : 0.0f;
            const __half h = __float2half_rn(val), l = __float2half_rn(val - __half2float(h));
            const uint32_t off = core_offset(o, k, LNP);
            *reinterpret_cast<__half*>(out + GL_V + off) = h;
            *reinterpret_cast<__half*>(out + GL_V + V_BYTES + off) = l;
        } else {
            const int o = i - LNP * LKP;
            reinterpret_cast<float*>(out + GL_C)[o] = o < A ? v[A * K0 + o] * sc : 0.0f;
        }
    }
}

}  // namespace

size_t lin_tc_prep_bytes() { return (size_t)round_up(GL_TOTAL, 256); }
bool lin_tc_supported(int K0, int A) { return K0 < LKP && A <= LNP && K0 >= 1; }

void launch_lin_tc_prep(const float* v, int K0, int A, const float* scale_dev, unsigned char* out, cudaStream_t s) {
    lin_tc_prep_kernel<<<48, 256, 0, s>>>(v, K0, A, scale_dev, out);
}

cudaError_t launch_linear_tc(const unsigned char* T, const float* theta, const float* in_shift, const float* in_scale,
                             const float* out_scale, const float* obs, int K0, int A, const int* idx, long long n,
                             float* gpartial, long long gstride, int tW, int tb, int tLS, int grid, cudaStream_t s) {
    LinTcArgs a;
    a.T = T; a.theta = theta; a.in_shift = in_shift; a.in_scale = in_scale; a.out_scale = out_scale; a.obs = obs;
    a.K0 = K0; a.A = A; a.idx = idx; a.n = n; a.gpartial = gpartial; a.gstride = gstride; a.tW = tW; a.tb = tb; a.tLS = tLS;
    cudaError_t e = cudaFuncSetAttribute(linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SL_TOTAL);
    if (e != cudaSuccess) return e;
    linear_tc_kernel<<<grid, 512, SL_TOTAL, s>>>(a);
    return cudaGetLastError();
}

}  // namespace mjb

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
#include <cstdlib>

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
// `ap` = rows per fp16 term: 32 for the register-prefetch kernel, round_up(A, 8) for the TMA-fed kernel (whose MMA
// N = 2 ap); the fp32 c[32] block always sits at GL_C.
__global__ void lin_tc_prep_kernel(const float* __restrict__ v, int K0, int A, int ap, const float* __restrict__ scale_dev,
                                   unsigned char* __restrict__ out) {
    const float sc = scale_dev ? *scale_dev : 1.0f;
    const int vr = 2 * ap;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ap * LKP + 32; i += gridDim.x * blockDim.x) {
        if (i < ap * LKP) {
            const int o = i / LKP, k = i % LKP;
            const float val = (o < A && k < K0) ? v[o * K0 + k] * sc : 0.0f;
            const __half h = __float2half_rn(val), l = __float2half_rn(val - __half2float(h));
            *reinterpret_cast<__half*>(out + GL_V + core_offset(o, k, vr)) = h;
            *reinterpret_cast<__half*>(out + GL_V + core_offset(ap + o, k, vr)) = l;
        } else {
            const int o = i - ap * LKP;
            reinterpret_cast<float*>(out + GL_C)[o] = o < A ? v[A * K0 + o] * sc : 0.0f;
        }
    }
}

// =====================================================================================================================
// TMA-fed variant (the default for dense batches): the observation rows are streamed by a dedicated producer thread
// with cp.async.bulk into a raw fp32 shared-memory ring, so HBM loads stay in flight through every phase of a tile --
// the register-prefetch kernel above stalls its loader whenever the warps are busy converting or waiting for an MMA
// (ncu: stall_lg_throttle) and reaches 2.5 TB/s; this one is bounded by the shared-memory data path instead.
//
//   warps 0-15  converters + dy epilogue      warp 16 (one thread)  TMA producer      warp 17 (one thread)  MMA issuer
//
//   raw ring slot = 16 consecutive timesteps x K0 floats = one contiguous bulk copy (dense rows; the converters' lane map
//   keeps their LDS.128 phases on distinct banks); 2-4 slots (whatever fits next to the tile).
//   staged tile  = as above, but rows are permuted so that the fp16 hi and lo rows of a sample sit in the SAME TMEM lane
//   quadrant, 8 lanes apart: row(m, t) = 16 (m / 8) + 8 t + m % 8  -> the hi*hi + hi*lo + lo*hi sum is one shuffle,
//   no shared-memory exchange and no block barrier in the epilogue.
//   The action axis is padded to AP = round_up(A, 8) instead of 32: MMA N = 2 AP, V and dy shrink accordingly (cfg5,
//   A = 17: N = 48, 36 KB of tangent), which is what makes room for the ring.
//   All hand-offs are mbarriers: full/empty per ring slot, staged (converters -> issuer), d1 (GEMM 1 -> epilogue),
//   dy (epilogue -> issuer), g2 (GEMM 2 -> converters: the tile buffer may be restaged).
constexpr int TW_CONV = 16;                       // converter warps
constexpr int T_THREADS = 32 * (TW_CONV + 2);     // 576
constexpr int CHUNK_ROWS = 16;                    // rows per ring slot; LM / CHUNK_ROWS = 4 slots' worth per tile.  (8-row slots
constexpr int MAX_SLOTS = 4;                      //  were measured slower: a cp.async.bulk costs its issuer ~250 cycles whatever its size)

struct LinTmaArgs {
    const unsigned char* T; const float* theta;
    const float* in_shift; const float* in_scale; const float* out_scale;
    const float* obs; int K0; int A; long long n;
    float* gpartial; long long gstride; int tW, tb, tLS;
    int ap;            // padded actions per fp16 term (8, 16, 24, 32)
    int nfb;           // feature blocks of 128 in use: ceil((K0 + 1) / 128)
    int pitch;         // floats between raw rows in a ring slot
    int slots;         // ring slots
    int map4x2;        // converter lane map: 0 = a phase reads 8 rows x 1 float4 column, 1 = 4 rows x 2 columns
    int off_v, off_dy, off_f32, off_ring, off_bar;   // shared-memory map (bytes); the staged tile sits at 0
    unsigned long long* prof;                        // developer aid: per-role clock64 sums (nullptr = off), 16 slots
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

template <bool IDENT>
__global__ void __launch_bounds__(T_THREADS, 1) linear_tc_tma_kernel(const LinTmaArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sf = reinterpret_cast<float*>(smem + a.off_f32);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.off_bar);
    // barrier indices
    enum { B_FULL = 0, B_EMPTY = MAX_SLOTS, B_STAGED = 2 * MAX_SLOTS, B_D1, B_DY, B_G2, B_V, B_COUNT };
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K0 = a.K0, A = a.A, AP = a.ap, NN = 2 * a.ap, NFB_ = a.nfb, S = a.slots, pitch = a.pitch;
    const int VRr = 2 * AP;                                           // rows of the tangent buffer
    const uint32_t v_lb = 16u * (uint32_t)VRr;
    const int v_bytes = VRr * LKP * 2, dy_bytes = XR * NN * 2;
    const int chunk_bytes = CHUNK_ROWS * pitch * 4;

    if (warp == 0) tmem_alloc(&s_tmem, TL_COLS);
    if (tid == 0) {
        for (int i = 0; i < MAX_SLOTS; ++i) { mbar_init(&bars[B_FULL + i], 1); mbar_init(&bars[B_EMPTY + i], TW_CONV); }
        mbar_init(&bars[B_STAGED], TW_CONV); mbar_init(&bars[B_D1], 1); mbar_init(&bars[B_DY], TW_CONV);
        mbar_init(&bars[B_G2], 1); mbar_init(&bars[B_V], 1);
    }
    for (int i = tid; i < dy_bytes / 16; i += T_THREADS) reinterpret_cast<uint4*>(smem + a.off_dy)[i] = make_uint4(0, 0, 0, 0);
    {   // this CTA's gradient partial starts at zero
        float4* gz = reinterpret_cast<float4*>(a.gpartial + (size_t)blockIdx.x * a.gstride);
        for (int i = tid; i < (int)(a.gstride / 4); i += T_THREADS) gz[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int k = tid; k < LKP; k += T_THREADS) {
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
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = s_tmem, sbase = smem_u32(smem);
    const long long n_tiles = (a.n + LM - 1) / LM;
    const long long G = gridDim.x;

    if (warp == TW_CONV) {
        // =============================== TMA producer ===============================
        // lane 0 arms the slot's barrier and issues ONE bulk copy for the slot's 16 contiguous rows (the other lanes only
        // keep the warp converged)
        {
            if (lane == 0) {
                mbar_expect_tx(&bars[B_V], (uint32_t)v_bytes);
                bulk_g2s(smem + a.off_v, a.T + GL_V, (uint32_t)v_bytes, &bars[B_V]);
            }
            long long cc = 0;                                          // chunks issued by this CTA
            const uint32_t row_bytes = (uint32_t)K0 * 4u;
            unsigned long long p_wait = 0, p_issue = 0;
            long long tp = clock64();
#define TMA_PROF(acc) do { if (a.prof) { const long long _t = clock64(); acc += (unsigned long long)(_t - tp); tp = _t; } } while (0)
            for (long long tile = blockIdx.x; tile < n_tiles; tile += G) {
                for (int c = 0; c < LM / CHUNK_ROWS; ++c, ++cc) {
                    const int slot = (int)(cc % S);
                    if (cc >= S) mbar_wait(&bars[B_EMPTY + slot], (uint32_t)(((cc / S) - 1) & 1));
                    if (lane == 0) TMA_PROF(p_wait);
                    const long long row0 = tile * LM + (long long)c * CHUNK_ROWS;
                    const int rows = (int)max(0LL, min((long long)CHUNK_ROWS, a.n - row0));
                    if (lane == 0) {
                        if (rows > 0) {
                            mbar_expect_tx(&bars[B_FULL + slot], row_bytes * (uint32_t)rows);
                            bulk_g2s(smem + a.off_ring + (size_t)slot * chunk_bytes, a.obs + row0 * K0,
                                     row_bytes * (uint32_t)rows, &bars[B_FULL + slot]);
                        } else {
                            mbar_arrive(&bars[B_FULL + slot]);         // nothing to load: complete the phase by hand
                        }
                        TMA_PROF(p_issue);
                    }
                    __syncwarp();
                }
            }
            if (a.prof && lane == 0) { atomicAdd(a.prof + 9, p_wait); atomicAdd(a.prof + 10, p_issue); }
        }
    } else if (warp == TW_CONV + 1) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            const uint32_t ID1 = make_idesc_f16(128, NN, false, false);   // A = tile (K-major), B = V (K-major)
            const uint32_t ID2 = make_idesc_f16(128, NN, true, true);     // A = tile (MN-major), B = dy (MN-major)
            mbar_wait(&bars[B_V], 0);
            long long it = 0;
            unsigned long long i_ws = 0, i_g1 = 0, i_wd = 0, i_g2 = 0;
            long long tp = clock64();
            for (long long tile = blockIdx.x; tile < n_tiles; tile += G, ++it) {
                const uint32_t par = (uint32_t)(it & 1);
                const bool first = (it % FLUSH_TILES == 0);
                mbar_wait(&bars[B_STAGED], par);
                TMA_PROF(i_ws);
                tcgen05_fence_after();
                for (int i = 0; i < NFB_; ++i) {
                    uint64_t da = make_desc(sbase + (uint32_t)i * XB_BYTES, X_LB, 128);
                    uint64_t db = make_desc(sbase + a.off_v + 16u * (uint32_t)i * v_lb, v_lb, 128);
#pragma unroll
                    for (int jj = 0; jj < FB / 16; ++jj) {
                        mma_f16(tmem + TL_D1, da, db, ID1, (i | jj) > 0);
                        da = desc_adv(da, 2 * X_LB); db = desc_adv(db, 2 * v_lb);
                    }
                }
                mma_commit(&bars[B_D1]);
                TMA_PROF(i_g1);
                mbar_wait(&bars[B_DY], par);
                TMA_PROF(i_wd);
                tcgen05_fence_after();
                for (int j = 0; j < NFB_; ++j) {
                    uint64_t da = make_desc(sbase + (uint32_t)j * XB_BYTES, 128, X_LB);
                    uint64_t db = make_desc(sbase + a.off_dy, 128, DY_LB);
#pragma unroll
                    for (int ks = 0; ks < XR / 16; ++ks) {
                        mma_f16(tmem + TL_G + (uint32_t)(NN * j), da, db, ID2, !first || ks > 0);
                        da = desc_adv(da, 256); db = desc_adv(db, 256);
                    }
                }
                mma_commit(&bars[B_G2]);
                TMA_PROF(i_g2);
            }
            if (a.prof) { atomicAdd(a.prof + 5, i_ws); atomicAdd(a.prof + 6, i_g1); atomicAdd(a.prof + 7, i_wd); atomicAdd(a.prof + 8, i_g2); }
        }
    } else {
        // =============================== converters + dy epilogue (warps 0-15) ===============================
        // lane -> (row r8 of the 8-row group, float4 column cidx of the 16-feature block).  Either map gives every
        // half-warp 8 rows x 2 columns (conflict-free 8-byte stores into one core-matrix row group) ...
        const int r8 = a.map4x2 ? ((lane & 3) + 4 * ((lane >> 3) & 1)) : (lane & 7);
        const int cidx = a.map4x2 ? (((lane >> 2) & 1) + 2 * (lane >> 4)) : (lane >> 3);
        const int q = warp & 3, cq = warp >> 2;                      // epilogue: TMEM lane quadrant, 8-action group
        const int e_grp = lane >> 4, e_t = (lane >> 3) & 1;          // epilogue lane -> (sample group, hi/lo row)
        const int e_m = 16 * q + 8 * e_grp + (lane & 7);             // sample of this lane's TMEM row (row = 32 q + lane)
        auto flush = [&]() {
            float* gp = a.gpartial + (size_t)blockIdx.x * a.gstride;
            if (warp < 4) {
                const int kl = 32 * warp + lane;
#pragma unroll 1
                for (int j = 0; j < NFB_; ++j) {
                    const int k = FB * j + kl;
#pragma unroll 1
                    for (int c8 = 0; 8 * c8 < A; ++c8) {
                        uint32_t g0[8], g1[8];
                        tmem_ld8(tmem + ((uint32_t)(32 * warp) << 16) + TL_G + NN * j + 8 * c8, g0);
                        tmem_ld8(tmem + ((uint32_t)(32 * warp) << 16) + TL_G + NN * j + AP + 8 * c8, g1);
                        tmem_ld_wait();
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
        long long it = 0, cc = 0;
        unsigned long long c_wg = 0, c_wf = 0, c_cv = 0, c_wd = 0, c_ep = 0;
        long long tp = clock64();
#define CV_PROF(acc) do { if (a.prof && tid == 0) { const long long _t = clock64(); acc += (unsigned long long)(_t - tp); tp = _t; } } while (0)
        for (long long tile = blockIdx.x; tile < n_tiles; tile += G, ++it) {
            const long long base = tile * LM;
            const uint32_t par = (uint32_t)(it & 1);
            if (it > 0) {                                            // GEMM 2 of the previous tile done: buffers free
                mbar_wait(&bars[B_G2], (uint32_t)((it - 1) & 1));
                tcgen05_fence_after();
                if (it % FLUSH_TILES == 0) flush();
            }
            CV_PROF(c_wg);
            // ---- convert the tile: 4 ring slots of 16 rows; per slot this warp does NFB warp-iterations of
            //      8 rows x 16 features (lane -> row r8, float4 column cidx)
            for (int c = 0; c < LM / CHUNK_ROWS; ++c, ++cc) {
                const int slot = (int)(cc % S);
                mbar_wait(&bars[B_FULL + slot], (uint32_t)((cc / S) & 1));
                CV_PROF(c_wf);
                const float* raw = reinterpret_cast<const float*>(smem + a.off_ring + (size_t)slot * chunk_bytes);
#pragma unroll
                for (int u = 0; u < NFB; ++u) {
                    if (u >= NFB_) break;
                    const int wi = warp + TW_CONV * u;
                    const int rg = wi & 1, cb = wi >> 1;             // row group of the slot, 16-feature column block
                    const int rl = 8 * rg + r8;                      // row inside the slot
                    const int m = CHUNK_ROWS * c + rl;               // sample inside the tile
                    const int k = 16 * cb + 4 * cidx;                // first feature of this lane's float4
                    float x0 = 0.f, x1 = 0.f, x2 = 0.f, x3 = 0.f;
                    const bool rv = (base + m) < a.n;
                    if (k < K0) {
                        if (rv) {
                            const float4 v = *reinterpret_cast<const float4*>(raw + (size_t)rl * pitch + k);
                            x0 = v.x; x1 = v.y; x2 = v.z; x3 = v.w;
                            if (!IDENT) {
                                const float4 sh = *reinterpret_cast<const float4*>(sf + SLF_SHIFT + k);
                                const float4 ri = *reinterpret_cast<const float4*>(sf + SLF_RINV + k);
                                x0 = (x0 - sh.x) * ri.x; x1 = (x1 - sh.y) * ri.y; x2 = (x2 - sh.z) * ri.z; x3 = (x3 - sh.w) * ri.w;
                            }
                        }
                    } else if (k == K0) {
                        x0 = 1.0f;                                   // ones column (bias gradient); K0 % 4 == 0
                    }
                    const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
                    const float2 b01 = __half22float2(h01), b23 = __half22float2(h23);
                    const __half2 l01 = __floats2half2_rn(x0 - b01.x, x1 - b01.y), l23 = __floats2half2_rn(x2 - b23.x, x3 - b23.y);
                    uint2 hv, lv;
                    hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
                    lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
                    // staged row of (sample m, term t): 16 (m / 8) + 8 t + m % 8
                    const int row_hi = 16 * (m >> 3) + (m & 7);
                    const int kk = k & (FB - 1);
                    unsigned char* o = smem + (size_t)(k >> 7) * XB_BYTES + core_offset(row_hi, kk, XR);
                    *reinterpret_cast<uint2*>(o) = hv;
                    *reinterpret_cast<uint2*>(o + 128) = lv;         // row + 8: the next core matrix
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[B_EMPTY + slot]);   // this warp is done reading the slot
                CV_PROF(c_cv);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[B_STAGED]);
            // ---- dy = fac * (ydot + c), masked ----
            mbar_wait(&bars[B_D1], par);
            CV_PROF(c_wd);
            tcgen05_fence_after();
            if (8 * cq < AP) {
                uint32_t y0[8], y1[8];
                const uint32_t tl = tmem + ((uint32_t)(32 * q) << 16) + TL_D1 + 8 * cq;
                tmem_ld8(tl, y0);
                tmem_ld8(tl + AP, y1);
                tmem_ld_wait();
                const bool valid = (base + e_m) < a.n;
                __half2 h[4], l[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int o = 2 * j, a0 = 8 * cq + o;
                    // hi row: x_hi v_hi + x_hi v_lo ; lo row: x_lo v_hi ; the pair sits 8 lanes apart
                    float s0f = __uint_as_float(y0[o]) + (e_t == 0 ? __uint_as_float(y1[o]) : 0.0f);
                    float s1f = __uint_as_float(y0[o + 1]) + (e_t == 0 ? __uint_as_float(y1[o + 1]) : 0.0f);
                    s0f += __shfl_xor_sync(0xffffffffu, s0f, 8);
                    s1f += __shfl_xor_sync(0xffffffffu, s1f, 8);
                    const float d0 = valid ? sf[SLF_FAC + a0] * (s0f + sf[SLF_C + a0]) : 0.0f;
                    const float d1 = valid ? sf[SLF_FAC + a0 + 1] * (s1f + sf[SLF_C + a0 + 1]) : 0.0f;
                    h[j] = __floats2half2_rn(d0, d1);
                    const float2 bk = __half22float2(h[j]);
                    l[j] = __floats2half2_rn(d0 - bk.x, d1 - bk.y);
                }
                // DY row r = 32 q + lane: hi rows get [dy hi | dy lo], lo rows [dy hi | 0] (the zero half is never written)
                const int r = 32 * q + lane;
                const uint32_t o = (uint32_t)cq * DY_LB + (uint32_t)((r >> 3) * 128 + (r & 7) * 16);
                *reinterpret_cast<uint4*>(smem + a.off_dy + o) = *reinterpret_cast<const uint4*>(h);
                if (e_t == 0) *reinterpret_cast<uint4*>(smem + a.off_dy + o + (uint32_t)(AP / 8) * DY_LB) = *reinterpret_cast<const uint4*>(l);
            }
            fence_proxy_async();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[B_DY]);
            CV_PROF(c_ep);
        }
        if (a.prof && tid == 0) {
            atomicAdd(a.prof + 0, c_wg); atomicAdd(a.prof + 1, c_wf); atomicAdd(a.prof + 2, c_cv); atomicAdd(a.prof + 3, c_wd);
            atomicAdd(a.prof + 4, c_ep); atomicAdd(a.prof + 11, (unsigned long long)it);
        }
        if (it > 0) {
            mbar_wait(&bars[B_G2], (uint32_t)((it - 1) & 1));
            tcgen05_fence_after();
            flush();
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, TL_COLS);
}


}  // namespace

void lin_tc_set_prof(unsigned long long* p) { g_lin_prof = p; }

size_t lin_tc_prep_bytes() { return (size_t)round_up(GL_TOTAL, 256); }
bool lin_tc_supported(int K0, int A) { return K0 < LKP && A <= LNP && K0 >= 1; }

// The TMA-fed kernel covers dense batches (no subsample gather) whose rows are 16-byte multiples.
static bool lin_tma_plan(int K0, int A, bool has_idx, LinTmaArgs* out) {
    if (has_idx || (K0 & 3) != 0 || K0 + 1 > LKP || A > LNP) return false;
    if (const char* env = getenv("MJRL_B200_LIN_TMA")) if (atoi(env) == 0) return false;
    LinTmaArgs a;
    a.ap = round_up(A, 8);
    a.nfb = (K0 + 1 + FB - 1) / FB;
    // ONE bulk copy per ring slot: the 16 rows of a slot are contiguous in global memory and land densely (pitch = K0).
    // (One copy per row into a padded pitch was measured first: the TMA unit needs ~80 cycles per cp.async.bulk whatever
    //  its size, 5 k cycles per 64-row tile, and the ring ran dry.)  The converters' LDS.128 stays conflict-free by
    //  choosing which rows x float4-columns the 8 lanes of a phase touch: K0/4 odd -> 8 rows x 1 column, else 4 rows x 2.
    a.pitch = K0;
    const int chunk_bytes = CHUNK_ROWS * a.pitch * 4;
    const int v_bytes = 2 * a.ap * LKP * 2, dy_bytes = XR * 2 * a.ap * 2;
    a.off_v = RING * XB_BYTES;
    a.off_dy = a.off_v + v_bytes;
    a.off_f32 = a.off_dy + dy_bytes;
    a.off_ring = round_up(a.off_f32 + SLF_END * 4, 128);
    a.map4x2 = (((K0 / 4) & 1) == 0) ? 1 : 0;
    const int max_smem = 232448 - 1024;                               // 227 KB opt-in limit minus static shared + slack
    int slots = (max_smem - 256 - a.off_ring) / chunk_bytes;
    if (slots > MAX_SLOTS) slots = MAX_SLOTS;
    if (slots < 2) return false;
    a.slots = slots;
    a.off_bar = a.off_ring + slots * chunk_bytes;
    if (out) *out = a;
    return true;
}

void launch_lin_tc_prep(const float* v, int K0, int A, bool has_idx, const float* scale_dev, unsigned char* out, cudaStream_t s) {
    const int ap = lin_tma_plan(K0, A, has_idx, nullptr) ? round_up(A, 8) : LNP;
    lin_tc_prep_kernel<<<48, 256, 0, s>>>(v, K0, A, ap, scale_dev, out);
}

cudaError_t launch_linear_tc(const unsigned char* T, const float* theta, const float* in_shift, const float* in_scale,
                             const float* out_scale, bool identity_in, const float* obs, int K0, int A, const int* idx,
                             long long n, float* gpartial, long long gstride, int tW, int tb, int tLS, int grid, cudaStream_t s) {
    LinTmaArgs ta;
    if (lin_tma_plan(K0, A, idx != nullptr, &ta)) {
        ta.T = T; ta.theta = theta; ta.in_shift = in_shift; ta.in_scale = in_scale; ta.out_scale = out_scale; ta.obs = obs;
        ta.K0 = K0; ta.A = A; ta.n = n; ta.gpartial = gpartial; ta.gstride = gstride; ta.tW = tW; ta.tb = tb; ta.tLS = tLS;
        ta.prof = g_lin_prof;
        const int smem = ta.off_bar + 256;
        auto kern = identity_in ? linear_tc_tma_kernel<true> : linear_tc_tma_kernel<false>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, T_THREADS, smem, s>>>(ta);
        return cudaGetLastError();
    }
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

// Fisher-vector product on 5th-generation tensor cores (tcgen05 + TMEM), for the 128x128 Gaussian MLP.
//
//   F v = (1/N) sum_n J_n^T W J_n v   (algos/npg_cg.py:62-81 in closed form)
//
// 89 % of the FVP's flops are dense 128x128x128 contractions per 128-sample tile (forward, tangent-forward,
// dgrad, wgrad).  They run as tcgen05.mma kind::f16 with fp32 accumulation in TMEM.  To stay at fp32-level
// accuracy every operand is split into two fp16 terms (x = hi + lo, 22 significant bits) and each logical
// product is three MMAs (hi*hi + lo*hi + hi*lo); the tangent vector is pre-scaled by a power of two so its
// magnitude is O(1) (exact to undo because the FVP is linear in v).
//
// Per CTA (one per SM, persistent over tiles of 128 consecutive timesteps):
//   * operands live in shared memory in the no-swizzle canonical "core-tiled" layout (tc_common.cuh), written by
//     the epilogue threads straight from TMEM; one buffer is read K-major by the forward/dgrad GEMMs and MN-major
//     by the weight-gradient GEMMs (reduction over samples), so nothing is ever transposed;
//   * weights (fp16 hi/lo, pre-split and pre-tiled in global memory by tc_prep_kernel) are streamed from L2 in
//     16 KB K-slices by TMA bulk copies through a two-slot ring (mbarrier expect_tx / tcgen05.commit);
//   * the weight-gradient accumulators G1 (128x32), G2 (128x144, the extra column block carries the bias
//     gradient via a ones column) and G3 (128x16) stay in TMEM for the whole kernel and are written once;
//   * tanh / tangent / delta computations are the epilogues: tcgen05.ld -> registers -> fp16 split -> smem.
// TMEM: D1 [0,128) D2 [128,256) G2 [256,400) G1 [400,432) G3 [432,448) of 512 allocated columns.
#include <cuda_fp16.h>

#include "kernels.h"
#include "tc_common.cuh"

namespace mjb {

using namespace tc;

namespace {

constexpr int TH = 128;                 // hidden width handled by this kernel
constexpr int TK0 = 32;                 // padded input width
constexpr int TM = 128;                 // samples per tile
constexpr int LBY = 16 * 128;           // bytes between 8-column groups of a 128-row core-tiled buffer
constexpr int CHUNK = 16384;            // one streamed weight K-slice: [hi 8 KB][lo 8 KB] of a [128 x 32] block

// Power-of-two scale of the back-propagated deltas (delta2, dh1, delta1 and hence G2 / G1; undone exactly at the
// output): keeps the fp16 LOW term of these small operands (|delta2| ~ 1e-3 at init) out of the subnormal range,
// where the two-term split would lose its 22 significant bits.
constexpr float SD = 64.0f;
constexpr int FLUSH_TILES = 12;           // tiles per TMEM accumulation group (see flush() in the kernel)

// ---- global (prepped) parameter block, bytes ----
constexpr int G_W1S = 0;
constexpr int G_W2S = G_W1S + CHUNK;
constexpr int G_W2TS = G_W2S + 4 * CHUNK;
constexpr int G_F32 = G_W2TS + 4 * CHUNK;
// fp32 block (floats): b1[128] b2[128] W3[8][128] b3[8] ls[8]
constexpr int F_B1 = 0, F_B2 = 128, F_W3 = 256, F_B3 = 256 + 1024, F_LS = F_B3 + 8, F_TOTAL = F_LS + 8;
constexpr int G_TOTAL = G_F32 + F_TOTAL * 4;

// ---- shared memory map, bytes ----
constexpr int S_PHI = 0, S_PLO = S_PHI + 128 * 144 * 2;
constexpr int S_QHI = S_PLO + 128 * 144 * 2, S_QLO = S_QHI + 128 * 128 * 2;
constexpr int S_XHI = S_QLO + 128 * 128 * 2, S_XLO = S_XHI + 128 * 32 * 2;
constexpr int S_DYHI = S_XLO + 128 * 32 * 2, S_DYLO = S_DYHI + 128 * 16 * 2;
constexpr int S_RING = S_DYLO + 128 * 16 * 2;
constexpr int S_F32 = S_RING + 2 * CHUNK;
// fp32 area (floats)
constexpr int SF_W3 = 0, SF_V3 = 1024, SF_B1 = 2048, SF_B2 = 2176, SF_C1 = 2304, SF_C2 = 2432, SF_B3 = 2560, SF_C3 = 2568,
              SF_FAC = 2576, SF_YSC = 2592, SF_DYS = SF_YSC + 4 * 128 * 8, SF_GB3 = SF_DYS + 128 * 8, SF_END = SF_GB3 + 32;
constexpr int S_BAR = S_F32 + SF_END * 4;
constexpr int S_TOTAL = S_BAR + 64;

// TMEM columns
constexpr uint32_t T_D1 = 0, T_D2 = 128, T_G2 = 256, T_G1 = 400, T_G3 = 432;

struct TcFvpArgs {
    const unsigned char* P;      // prepped theta
    const unsigned char* T;      // prepped (scaled) tangent
    const float* in_shift; const float* in_scale; const float* out_scale;
    const float* obs; int obs_dim; int A; const int* idx; long long n;
    float* gpartial; long long gstride;
    int tW1, tb1, tW2, tb2, tW3, tb3, K0, h1, h2;
};

// tanh via one ex2 and one rcp: 1 - 2/(exp(2|x|)+1) with the sign restored.  Absolute error ~5e-8 (the same order as
// the fp32 rounding of values near 1); used only on the tensor-core path, whose operands are rounded to 22 bits anyway.
__device__ __forceinline__ float tanh_fast(float x) {
    const float e = __expf(2.0f * fabsf(x));
    const float t = 1.0f - __fdividef(2.0f, e + 1.0f);
    return copysignf(t, x);
}

__device__ __forceinline__ void split16(float v, __half& hi, __half& lo) {
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}

// write 16 consecutive columns [c0, c0+16) of row m (fp32 values) as fp16 hi / lo into a core-tiled buffer pair
__device__ __forceinline__ void store_split16(unsigned char* smem, int off_hi, int off_lo, int m, int c0, const float (&v)[16]) {
    const int rowoff = (m >> 3) * 128 + (m & 7) * 16;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        __half2 h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float a = v[8 * g + 2 * j], b = v[8 * g + 2 * j + 1];
            h[j] = __floats2half2_rn(a, b);                       // one cvt.rn.f16x2.f32 for two values
            const float2 back = __half22float2(h[j]);
            l[j] = __floats2half2_rn(a - back.x, b - back.y);
        }
        const int o = ((c0 >> 3) + g) * LBY + rowoff;
        *reinterpret_cast<uint4*>(smem + off_hi + o) = *reinterpret_cast<const uint4*>(h);
        *reinterpret_cast<uint4*>(smem + off_lo + o) = *reinterpret_cast<const uint4*>(l);
    }
}

// 8 consecutive columns [c0, c0+8) of row m -> one 16-byte core-matrix row of the hi buffer and one of the lo buffer
__device__ __forceinline__ void store_split8(unsigned char* smem, int off_hi, int off_lo, int m, int c0, const float (&v)[8]) {
    __half2 h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float a = v[2 * j], b = v[2 * j + 1];
        h[j] = __floats2half2_rn(a, b);                           // one cvt.rn.f16x2.f32 for two values
        const float2 back = __half22float2(h[j]);
        l[j] = __floats2half2_rn(a - back.x, b - back.y);
    }
    const int o = (c0 >> 3) * LBY + (m >> 3) * 128 + (m & 7) * 16;
    *reinterpret_cast<uint4*>(smem + off_hi + o) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(smem + off_lo + o) = *reinterpret_cast<const uint4*>(l);
}

// D (+)= A * B^T with two-term fp16 operands: hi*hi + lo*hi + hi*lo
__device__ __forceinline__ void mma3(uint32_t d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo, uint32_t idesc, bool acc) {
    mma_f16(d, a_hi, b_hi, idesc, acc);
    mma_f16(d, a_lo, b_hi, idesc, true);
    mma_f16(d, a_hi, b_lo, idesc, true);
}

__global__ void __launch_bounds__(512, 1) fvp_tc_kernel(const TcFvpArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sf = reinterpret_cast<float*>(smem + S_F32);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S_BAR);     // [0] mma done, [1..2] ring full, [3..4] ring empty
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, cq = warp >> 2;                          // TMEM lane quadrant, column quarter (32 columns)
    const int m = 32 * q + lane;                                     // sample row of this thread inside the tile
    const int A = a.A;

    // ---- one-time setup ----
    if (warp == 0) tmem_alloc(&s_tmem, 512);
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
        mbar_init(&bars[3], 1); mbar_init(&bars[4], 1);
    }
    for (int i = tid; i < (S_RING - S_PHI) / 16; i += 512) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    {   // this CTA's gradient partial starts at zero (no memset node in front of the kernel); gstride is a multiple of 32
        float4* gz = reinterpret_cast<float4*>(a.gpartial + (size_t)blockIdx.x * a.gstride);
        for (int i = tid; i < (int)(a.gstride / 4); i += 512) gz[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    if (tid < 128) {                                                 // ones column (feature 128) of P: carries d/db2 through G2
        const __half one = __float2half_rn(1.0f);
        *reinterpret_cast<__half*>(smem + S_PHI + core_offset(tid, 128, 128)) = one;
    }
    {   // fp32 side tables
        const float* pf = reinterpret_cast<const float*>(a.P + G_F32);
        const float* tf = reinterpret_cast<const float*>(a.T + G_F32);
        // packed last-layer table: sf[SF_W3 + n*16 + {0..7}] = W3[a][n], sf[SF_W3 + n*16 + 8 + {0..7}] = V3[a][n]
        for (int i = tid; i < 1024; i += 512) {
            const int aa = i / 128, n = i % 128;
            sf[SF_W3 + n * 16 + aa] = pf[F_W3 + i];
            sf[SF_W3 + n * 16 + 8 + aa] = tf[F_W3 + i];
        }
        if (tid < 128) { sf[SF_B1 + tid] = pf[F_B1 + tid]; sf[SF_B2 + tid] = pf[F_B2 + tid]; sf[SF_C1 + tid] = tf[F_B1 + tid]; sf[SF_C2 + tid] = tf[F_B2 + tid]; }
        if (tid < 8) {
            sf[SF_B3 + tid] = pf[F_B3 + tid]; sf[SF_C3 + tid] = tf[F_B3 + tid];
            float f = 0.0f;
            if (tid < A) {
                const float sd = expf(pf[F_LS + tid]);
                const float os = a.out_scale[tid];
                f = os * os * (2.0f / (2.0f * sd * sd + 1e-8f));
            }
            sf[SF_FAC + tid] = f;
        }
    }
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t tlane = (uint32_t)(32 * q) << 16;
    const uint32_t sbase = smem_u32(smem);

    // ---- instruction descriptors ----
    const uint32_t ID_KK128 = make_idesc_f16(128, 128, false, false);
    const uint32_t ID_MM16 = make_idesc_f16(128, 16, true, true);
    const uint32_t ID_MM144 = make_idesc_f16(128, 144, true, true);
    const uint32_t ID_MM32 = make_idesc_f16(128, 32, true, true);

    // ---- ring state (thread 0 only) ----
    uint32_t mma_par = 0, fpar[2] = {0, 0}, epar[2] = {0, 0};
    bool busy[2] = {false, false};
    auto ring_load = [&](int slot, const unsigned char* src) {
        if (busy[slot]) { mbar_wait(&bars[3 + slot], epar[slot]); epar[slot] ^= 1; busy[slot] = false; }
        mbar_expect_tx(&bars[1 + slot], CHUNK);
        bulk_g2s(smem + S_RING + slot * CHUNK, src, CHUNK, &bars[1 + slot]);
    };
    auto ring_wait = [&](int slot) { mbar_wait(&bars[1 + slot], fpar[slot]); fpar[slot] ^= 1; tcgen05_fence_after(); };
    auto ring_release = [&](int slot) { mma_commit(&bars[3 + slot]); busy[slot] = true; };
    // K-major descriptors: A from a sample-row buffer at feature column c (multiple of 16); B from a ring slot, k-step j
    auto descA_k = [&](int off, int c) { return make_desc(sbase + off + (c >> 3) * LBY, LBY, 128); };
    auto descB_k = [&](int slot, int lo, int j) { return make_desc(sbase + S_RING + slot * CHUNK + lo * 8192 + 2 * j * LBY, LBY, 128); };
    // MN-major descriptor (MMA rows = feature, reduction = samples 16j..16j+15)
    auto desc_mn = [&](int off, int j) { return make_desc(sbase + off + 2 * j * 128, 128, LBY); };
    auto all_wait_mma = [&]() { mbar_wait(&bars[0], mma_par); mma_par ^= 1; tcgen05_fence_after(); };

    // Adds the TMEM gradient accumulators to the per-CTA partial in global memory (reference theta layout; the partial
    // is zeroed before the launch and owned by this CTA).  Called every FLUSH_TILES tiles: the tensor core's fp32
    // accumulation truncates, so an accumulator that lives for the whole kernel drifts by ~4e-7 (relative) per tile --
    // 2.5e-5 after the 53 tiles per SM of a 1e6-timestep batch (tools/fvp_accuracy.py); short groups + fp32 adds keep
    // the kernel at ~4e-6.
    // (red.global.add: fire-and-forget, no read round trip; the slice belongs to this CTA, so the order is fixed)
    auto flush = [&]() {
        float* gp = a.gpartial + (size_t)blockIdx.x * a.gstride;
        // G2: lane = n (row of W2), columns k < h1 ; column 128 = d/db2[n]
        for (int cc = 0; cc < 2; ++cc) {
            const int c0 = 32 * cq + 16 * cc;
            uint32_t g[16];
            tmem_ld16(tmem + tlane + T_G2 + c0, g);
            tmem_ld_wait();
            if (m < a.h2) {
                float* row = gp + a.tW2 + m * a.h1 + c0;
                if (c0 + 16 <= a.h1 && ((a.tW2 + m * a.h1) & 3) == 0) {     // 64 contiguous, aligned bytes: vector read-modify-write
                    float4 o4[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) o4[j] = reinterpret_cast<const float4*>(row)[j];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        o4[j].x += __uint_as_float(g[4 * j]) * (1.0f / SD); o4[j].y += __uint_as_float(g[4 * j + 1]) * (1.0f / SD);
                        o4[j].z += __uint_as_float(g[4 * j + 2]) * (1.0f / SD); o4[j].w += __uint_as_float(g[4 * j + 3]) * (1.0f / SD);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) reinterpret_cast<float4*>(row)[j] = o4[j];
                } else {
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < a.h1) atomicAdd(&row[j], __uint_as_float(g[j]) * (1.0f / SD));
                }
            }
        }
        if (cq == 0) {
            uint32_t g[16];
            tmem_ld16(tmem + tlane + T_G2 + 128, g);
            tmem_ld_wait();
            if (m < a.h2) atomicAdd(&gp[a.tb2 + m], __uint_as_float(g[0]) * (1.0f / SD));
            uint32_t g1[32];
            tmem_ld32(tmem + tlane + T_G1, g1);
            tmem_ld_wait();
            if (m < a.h1) {
                for (int j = 0; j < a.K0; ++j) atomicAdd(&gp[a.tW1 + m * a.K0 + j], __uint_as_float(g1[j]) * (1.0f / SD));
                atomicAdd(&gp[a.tb1 + m], __uint_as_float(g1[a.K0]) * (1.0f / SD));
            }
            uint32_t g3[16];
            tmem_ld16(tmem + tlane + T_G3, g3);                       // lane = k (unit of h2), columns a
            tmem_ld_wait();
            if (m < a.h2)
                for (int i = 0; i < A; ++i) atomicAdd(&gp[a.tW3 + i * a.h2 + m], __uint_as_float(g3[i]));
        }
        tcgen05_fence_before();
    };
    float gb3_acc = 0.0f;                                             // thread a < 8: running sum_m dy[m][a]
    const long long n_tiles = (a.n + TM - 1) / TM;
    // Input staging: thread -> (row = tid / 4, 8 consecutive features (tid % 4) * 8 ..), i.e. one 16-byte core-matrix row
    // of the staged tile per thread (one vector store for the hi half, one for the lo half); four threads read a row's
    // <= 124 contiguous bytes.  The raw values of the NEXT tile are fetched into registers while the current tile
    // computes (their L2 / HBM latency used to sit in front of every tile).  The ones column (bias gradient through
    // G1) sits at feature obs_dim.
    const int x_r = tid >> 2, x_k0 = (tid & 3) * 8;
    float x_raw[8], x_sh[8], x_ri[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = x_k0 + j;
        x_raw[j] = 0.0f;
        x_sh[j] = k < a.obs_dim ? a.in_shift[k] : 0.0f;
        x_ri[j] = k < a.obs_dim ? 1.0f / (a.in_scale[k] + 1e-8f) : 0.0f;
    }
    auto fetch_tile = [&](long long tile) {
        const long long row = tile * TM + x_r;
        const bool rv = row < a.n;
        const long long rr = rv ? (a.idx ? (long long)a.idx[row] : row) : 0;
        const float* src = a.obs + rr * a.obs_dim + x_k0;
#pragma unroll
        for (int j = 0; j < 8; ++j) x_raw[j] = (rv && x_k0 + j < a.obs_dim) ? __ldg(src + j) : 0.0f;
    };
    if ((long long)blockIdx.x < n_tiles) fetch_tile(blockIdx.x);
    long long it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const long long base = tile * TM;
        const bool first = (it % FLUSH_TILES == 0);                  // first tile of an accumulation group
        // ================= P0: stage the input tile (transform, split) =================
        {
            float xv[8];
            const bool rv = base + x_r < a.n;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                xv[j] = rv ? (x_raw[j] - x_sh[j]) * x_ri[j] : 0.0f;
                if (x_k0 + j == a.obs_dim) xv[j] = 1.0f;
            }
            store_split8(smem, S_XHI, S_XLO, x_r, x_k0, xv);
        }
        if (tile + gridDim.x < n_tiles) fetch_tile(tile + gridDim.x);     // next tile: in flight under this tile's phases
        fence_proxy_async();
        tcgen05_fence_before();
        __syncthreads();
        // ================= P1: z1 = x W1^T -> D1, zd1 = x V1^T -> D2 =================
        if (tid == 0) {
            tcgen05_fence_after();
            if (it == 0) { ring_load(0, a.P + G_W1S); ring_load(1, a.T + G_W1S); }   // later tiles: prefetched at the end of P7
            ring_wait(0);
#pragma unroll
            for (int j = 0; j < 2; ++j)
                mma3(tmem + T_D1, descA_k(S_XHI, 16 * j), descA_k(S_XLO, 16 * j), descB_k(0, 0, j), descB_k(0, 1, j), ID_KK128, j > 0);
            ring_release(0);
            ring_wait(1);
#pragma unroll
            for (int j = 0; j < 2; ++j)
                mma3(tmem + T_D2, descA_k(S_XHI, 16 * j), descA_k(S_XLO, 16 * j), descB_k(1, 0, j), descB_k(1, 1, j), ID_KK128, j > 0);
            ring_release(1);
            mma_commit(&bars[0]);
        }
        all_wait_mma();
        // the ring is idle through the epilogue: the first two weight slices of the next GEMM phase stream in under it
        if (tid == 0) { ring_load(0, a.P + G_W2S); ring_load(1, a.P + G_W2S + CHUNK); }
        // ================= P2: h1 = tanh(z1+b1) -> P ; hd1 = (1-h1^2)(zd1+c1) -> Q =================
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
            const int c0 = 32 * cq + 16 * cc;
            uint32_t z[16], zd[16];
            tmem_ld16(tmem + tlane + T_D1 + c0, z);
            tmem_ld16(tmem + tlane + T_D2 + c0, zd);
            tmem_ld_wait();
            float h[16], hd[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float hv = tanh_fast(__uint_as_float(z[j]) + sf[SF_B1 + c0 + j]);
                h[j] = hv;
                hd[j] = (1.0f - hv * hv) * (__uint_as_float(zd[j]) + sf[SF_C1 + c0 + j]);
            }
            store_split16(smem, S_PHI, S_PLO, m, c0, h);
            store_split16(smem, S_QHI, S_QLO, m, c0, hd);
        }
        fence_proxy_async();
        tcgen05_fence_before();
        __syncthreads();
        // ================= P3: z2 = h1 W2^T -> D1 ; zd2 = hd1 W2^T + h1 V2^T -> D2 =================
        if (tid == 0) {
            tcgen05_fence_after();
            for (int s = 0; s < 8; ++s) {                             // 4 slices of W2, then 4 slices of V2 (0, 1 prefetched)
                const int slot = s & 1;
                ring_wait(slot);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int c = 32 * (s & 3) + 16 * j;
                    if (s < 4) {
                        mma3(tmem + T_D1, descA_k(S_PHI, c), descA_k(S_PLO, c), descB_k(slot, 0, j), descB_k(slot, 1, j), ID_KK128, (s | j) > 0);
                        mma3(tmem + T_D2, descA_k(S_QHI, c), descA_k(S_QLO, c), descB_k(slot, 0, j), descB_k(slot, 1, j), ID_KK128, (s | j) > 0);
                    } else {
                        mma3(tmem + T_D2, descA_k(S_PHI, c), descA_k(S_PLO, c), descB_k(slot, 0, j), descB_k(slot, 1, j), ID_KK128, true);
                    }
                }
                ring_release(slot);
                if (s + 2 < 8) ring_load(slot, (s + 2 < 4) ? a.P + G_W2S + (s + 2) * CHUNK : a.T + G_W2S + (s + 2 - 4) * CHUNK);
            }
            mma_commit(&bars[0]);
        }
        all_wait_mma();
        if (tid == 0) { ring_load(0, a.P + G_W2TS); ring_load(1, a.P + G_W2TS + CHUNK); }   // for P7, under P4 .. P6
        // ================= P4: h2 -> Q (operand) and back into D1 (fp32); ydot on the CUDA cores; delta_y =================
        {
            float yacc[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) yacc[i] = 0.0f;
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                const int c0 = 32 * cq + 16 * cc;
                uint32_t z[16], zd[16];
                tmem_ld16(tmem + tlane + T_D1 + c0, z);
                tmem_ld16(tmem + tlane + T_D2 + c0, zd);
                tmem_ld_wait();
                float h[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int n = c0 + j;
                    const float hv = tanh_fast(__uint_as_float(z[j]) + sf[SF_B2 + n]);
                    const float hdv = (1.0f - hv * hv) * (__uint_as_float(zd[j]) + sf[SF_C2 + n]);
                    h[j] = hv;
                    z[j] = __float_as_uint(hv);
                    const float4* wv = reinterpret_cast<const float4*>(sf + SF_W3 + n * 16);
                    const float4 w0 = wv[0], w1 = wv[1], v0 = wv[2], v1 = wv[3];   // rows >= A are zero
                    yacc[0] = fmaf(hdv, w0.x, fmaf(hv, v0.x, yacc[0])); yacc[1] = fmaf(hdv, w0.y, fmaf(hv, v0.y, yacc[1]));
                    yacc[2] = fmaf(hdv, w0.z, fmaf(hv, v0.z, yacc[2])); yacc[3] = fmaf(hdv, w0.w, fmaf(hv, v0.w, yacc[3]));
                    if (A > 4) {
                        yacc[4] = fmaf(hdv, w1.x, fmaf(hv, v1.x, yacc[4])); yacc[5] = fmaf(hdv, w1.y, fmaf(hv, v1.y, yacc[5]));
                        yacc[6] = fmaf(hdv, w1.z, fmaf(hv, v1.z, yacc[6])); yacc[7] = fmaf(hdv, w1.w, fmaf(hv, v1.w, yacc[7]));
                    }
                }
                tmem_st16(tmem + tlane + T_D1 + c0, z);               // keep h2 (fp32) in TMEM for the delta2 epilogue
                store_split16(smem, S_QHI, S_QLO, m, c0, h);
            }
            tmem_st_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) sf[SF_YSC + (cq * 128 + m) * 8 + i] = yacc[i];
        }
        __syncthreads();
        if (cq == 0) {
            const bool valid = (base + m) < a.n;
            __half dh[8], dl[8];
            float dyv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float y = (sf[SF_YSC + m * 8 + i] + sf[SF_YSC + (128 + m) * 8 + i]) +
                          (sf[SF_YSC + (256 + m) * 8 + i] + sf[SF_YSC + (384 + m) * 8 + i]) + sf[SF_C3 + i];
                y = (valid && i < A) ? sf[SF_FAC + i] * y : 0.0f;
                dyv[i] = y;
                sf[SF_DYS + m * 8 + i] = y;
                split16(y, dh[i], dl[i]);
            }
            const int o = (m >> 3) * 128 + (m & 7) * 16;
            *reinterpret_cast<uint4*>(smem + S_DYHI + o) = *reinterpret_cast<const uint4*>(dh);
            *reinterpret_cast<uint4*>(smem + S_DYLO + o) = *reinterpret_cast<const uint4*>(dl);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float s = warp_sum(dyv[i]);
                if (lane == 0) sf[SF_GB3 + q * 8 + i] = s;
            }
        }
        fence_proxy_async();
        tcgen05_fence_before();
        __syncthreads();
        if (tid < 8) gb3_acc += (sf[SF_GB3 + tid] + sf[SF_GB3 + 8 + tid]) + (sf[SF_GB3 + 16 + tid] + sf[SF_GB3 + 24 + tid]);
        // ================= P5: G3[k][a] += sum_m h2[m][k] dy[m][a] =================
        if (tid == 0) {
            tcgen05_fence_after();
#pragma unroll
            for (int j = 0; j < 8; ++j)
                mma3(tmem + T_G3, desc_mn(S_QHI, j), desc_mn(S_QLO, j), desc_mn(S_DYHI, j), desc_mn(S_DYLO, j), ID_MM16, !first || j > 0);
            mma_commit(&bars[0]);
        }
        all_wait_mma();
        // ================= P6: delta2 = (1-h2^2) (dy W3) -> Q =================
        {
            float dyr[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) dyr[i] = sf[SF_DYS + m * 8 + i];
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                const int c0 = 32 * cq + 16 * cc;
                uint32_t hz[16];
                tmem_ld16(tmem + tlane + T_D1 + c0, hz);
                tmem_ld_wait();
                float d[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int n = c0 + j;
                    const float hv = __uint_as_float(hz[j]);
                    const float4* wv = reinterpret_cast<const float4*>(sf + SF_W3 + n * 16);
                    const float4 w0 = wv[0];
                    float t = fmaf(dyr[0], w0.x, fmaf(dyr[1], w0.y, fmaf(dyr[2], w0.z, dyr[3] * w0.w)));
                    if (A > 4) {
                        const float4 w1 = wv[1];
                        t += fmaf(dyr[4], w1.x, fmaf(dyr[5], w1.y, fmaf(dyr[6], w1.z, dyr[7] * w1.w)));
                    }
                    d[j] = SD * ((1.0f - hv * hv) * t);          // scaled: delta2 ~ dy W3 is small (W3 ~ 1e-3 at init)
                }
                store_split16(smem, S_QHI, S_QLO, m, c0, d);
            }
        }
        fence_proxy_async();
        tcgen05_fence_before();
        __syncthreads();
        // ================= P7: dh1 = delta2 W2 -> D2 ; G2[n][k|1] += sum_m delta2[m][n] [h1|1][m][k] =================
        if (tid == 0) {
            tcgen05_fence_after();
#pragma unroll
            for (int j = 0; j < 8; ++j)
                mma3(tmem + T_G2, desc_mn(S_QHI, j), desc_mn(S_QLO, j), desc_mn(S_PHI, j), desc_mn(S_PLO, j), ID_MM144, !first || j > 0);
            for (int s = 0; s < 4; ++s) {
                const int slot = s & 1;
                ring_wait(slot);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int c = 32 * s + 16 * j;
                    mma3(tmem + T_D2, descA_k(S_QHI, c), descA_k(S_QLO, c), descB_k(slot, 0, j), descB_k(slot, 1, j), ID_KK128, (s | j) > 0);
                }
                ring_release(slot);
                if (s + 2 < 4) ring_load(slot, a.P + G_W2TS + (s + 2) * CHUNK);
            }
            mma_commit(&bars[0]);
        }
        all_wait_mma();
        if (tid == 0 && tile + gridDim.x < n_tiles) { ring_load(0, a.P + G_W1S); ring_load(1, a.T + G_W1S); }   // next tile's P1
        // ================= P8: delta1 = (1-h1^2) dh1 -> Q =================
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
            const int c0 = 32 * cq + 16 * cc;
            uint32_t dz[16];
            tmem_ld16(tmem + tlane + T_D2 + c0, dz);
            tmem_ld_wait();
            float d[16];
            const int rowoff = (m >> 3) * 128 + (m & 7) * 16;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const int o = ((c0 >> 3) + g) * LBY + rowoff;
                const uint4 uh = *reinterpret_cast<const uint4*>(smem + S_PHI + o);
                const uint4 ul = *reinterpret_cast<const uint4*>(smem + S_PLO + o);
                const __half2* hh = reinterpret_cast<const __half2*>(&uh);
                const __half2* hl = reinterpret_cast<const __half2*>(&ul);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 a2 = __half22float2(hh[j]), b2 = __half22float2(hl[j]);
                    const float h0 = a2.x + b2.x, h1v = a2.y + b2.y;
                    d[8 * g + 2 * j] = (1.0f - h0 * h0) * __uint_as_float(dz[8 * g + 2 * j]);
                    d[8 * g + 2 * j + 1] = (1.0f - h1v * h1v) * __uint_as_float(dz[8 * g + 2 * j + 1]);
                }
            }
            store_split16(smem, S_QHI, S_QLO, m, c0, d);
        }
        fence_proxy_async();
        tcgen05_fence_before();
        __syncthreads();
        // ================= P9: G1[n][k0|1] += sum_m delta1[m][n] [x|1][m][k0] =================
        if (tid == 0) {
            tcgen05_fence_after();
#pragma unroll
            for (int j = 0; j < 8; ++j)
                mma3(tmem + T_G1, desc_mn(S_QHI, j), desc_mn(S_QLO, j), desc_mn(S_XHI, j), desc_mn(S_XLO, j), ID_MM32, !first || j > 0);
            mma_commit(&bars[0]);
        }
        all_wait_mma();                                               // X / Q are rewritten by the next tile
        if ((it + 1) % FLUSH_TILES == 0) flush();                     // (every thread reads its own TMEM lanes / columns)
    }

    if (it > 0 && it % FLUSH_TILES != 0) flush();
    if (it > 0 && tid < A) a.gpartial[(size_t)blockIdx.x * a.gstride + a.tb3 + tid] = gb3_acc;
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// theta (or tangent) in the reference flat layout -> fp16 hi/lo core-tiled K-slices + fp32 side tables
__global__ void tc_prep_kernel(const float* __restrict__ th, PrepLayout L, const float* __restrict__ scale_dev,
                               unsigned char* __restrict__ out) {
    const float sc = scale_dev ? *scale_dev : 1.0f;
    const int n_w1 = 128 * 32, n_w2 = 128 * 128;
    const int total = n_w1 + 2 * n_w2 + F_TOTAL;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        if (i < n_w1) {                                              // W1s: rows n, cols k
            const int n = i / 32, k = i % 32;
            const float v = (n < L.h1 && k < L.K0) ? th[L.tW1 + n * L.K0 + k] * sc : 0.0f;
            __half h, l; split16(v, h, l);
            const uint32_t o = core_offset(n, k, 128);
            *reinterpret_cast<__half*>(out + G_W1S + o) = h;
            *reinterpret_cast<__half*>(out + G_W1S + 8192 + o) = l;
        } else if (i < n_w1 + n_w2) {                                // W2s: rows n, cols k (slice k/32)
            const int j = i - n_w1, n = j / 128, k = j % 128;
            const float v = (n < L.h2 && k < L.h1) ? th[L.tW2 + n * L.h1 + k] * sc : 0.0f;
            __half h, l; split16(v, h, l);
            const uint32_t o = (k / 32) * CHUNK + core_offset(n, k % 32, 128);
            *reinterpret_cast<__half*>(out + G_W2S + o) = h;
            *reinterpret_cast<__half*>(out + G_W2S + 8192 + o) = l;
        } else if (i < n_w1 + 2 * n_w2) {                            // W2Ts: rows k, cols n (slice n/32)
            const int j = i - n_w1 - n_w2, k = j / 128, n = j % 128;
            const float v = (n < L.h2 && k < L.h1) ? th[L.tW2 + n * L.h1 + k] * sc : 0.0f;
            __half h, l; split16(v, h, l);
            const uint32_t o = (n / 32) * CHUNK + core_offset(k, n % 32, 128);
            *reinterpret_cast<__half*>(out + G_W2TS + o) = h;
            *reinterpret_cast<__half*>(out + G_W2TS + 8192 + o) = l;
        } else {
            const int j = i - n_w1 - 2 * n_w2;
            float v = 0.0f;
            if (j < F_B2) { if (j < L.h1) v = th[L.tb1 + j] * sc; }
            else if (j < F_W3) { const int n = j - F_B2; if (n < L.h2) v = th[L.tb2 + n] * sc; }
            else if (j < F_B3) { const int o = (j - F_W3) / 128, k = (j - F_W3) % 128; if (o < L.A && k < L.h2) v = th[L.tW3 + o * L.h2 + k] * sc; }
            else if (j < F_LS) { const int o = j - F_B3; if (o < L.A) v = th[L.tb3 + o] * sc; }
            else { const int o = j - F_LS; if (o < L.A) v = th[L.tLS + o]; }
            reinterpret_cast<float*>(out + G_F32)[j] = v;
        }
    }
}

// power-of-two scale that brings max|v| into [1, 2); out[0] = scale, out[1] = 1/scale
__global__ void tc_vscale_kernel(const float* __restrict__ v, int d, float* out) {
    __shared__ float red[32];
    float mx = 0.0f;
    for (int i = threadIdx.x; i < d; i += blockDim.x) mx = fmaxf(mx, fabsf(v[i]));
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
        int e = 0;
        float s = 1.0f;
        if (mx > 0.0f && isfinite(mx)) { frexpf(mx, &e); s = ldexpf(1.0f, 1 - e); }
        out[0] = s;
        out[1] = 1.0f / s;
    }
}

}  // namespace

size_t fvp_tc_prep_bytes() { return (size_t)round_up(G_TOTAL, 256); }

bool fvp_tc_supported(const PrepLayout& L) {
    return L.H == 128 && L.K0 < 32 && L.A <= 8 && L.h1 <= 128 && L.h2 <= 128;
}

void launch_tc_prep(const float* theta, const PrepLayout& L, const float* scale_dev, unsigned char* out, cudaStream_t s) {
    tc_prep_kernel<<<148, 256, 0, s>>>(theta, L, scale_dev, out);
}
void launch_tc_vscale(const float* v, int d, float* out2, cudaStream_t s) { tc_vscale_kernel<<<1, 1024, 0, s>>>(v, d, out2); }

cudaError_t launch_fvp_tc(const PrepLayout& L, const unsigned char* P, const unsigned char* T, const float* in_shift,
                          const float* in_scale, const float* out_scale, const float* obs, const int* idx, long long n,
                          float* gpartial, long long gstride, int grid, cudaStream_t s) {
    TcFvpArgs a;
    a.P = P; a.T = T; a.in_shift = in_shift; a.in_scale = in_scale; a.out_scale = out_scale;
    a.obs = obs; a.obs_dim = L.K0; a.A = L.A; a.idx = idx; a.n = n; a.gpartial = gpartial; a.gstride = gstride;
    a.tW1 = L.tW1; a.tb1 = L.tb1; a.tW2 = L.tW2; a.tb2 = L.tb2; a.tW3 = L.tW3; a.tb3 = L.tb3; a.K0 = L.K0; a.h1 = L.h1; a.h2 = L.h2;
    cudaError_t e = cudaFuncSetAttribute(fvp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S_TOTAL);
    if (e != cudaSuccess) return e;
    fvp_tc_kernel<<<grid, 512, S_TOTAL, s>>>(a);
    return cudaGetLastError();
}

}  // namespace mjb

// MLPBaseline.fit, model-parallel cluster version (baselines/mlp_baseline.py:61-95, utils/optimize_model.py:7-36).
//
// The minibatch-Adam chain is sequential, so per step there are only 64 rows x 19.6 k parameters of work and the
// step time is set by latency, not throughput.  This kernel minimises what has to cross SMs per step:
//
//   * a thread-block cluster of C CTAs splits the HIDDEN UNITS: CTA c owns units [c*u, (c+1)*u) of both hidden
//     layers, i.e. rows of W1 / W2, the matching biases, entries of W3 -- and their Adam moments.  Weights and
//     optimizer state never leave the owner's shared memory (no gradient all-reduce, no weight broadcast);
//   * what is exchanged per step, through distributed shared memory (st.shared::cluster) between three hardware
//     cluster barriers, is small: the owner's slice of h1 to everybody (64 x u floats), 64 partial outputs, and
//     the owner-destined slices of the partial dgrad (64 x u floats per peer);
//   * every CTA gathers the same 64 minibatch rows itself (prefetched two steps / one step ahead).
//
// All cross-CTA sums run in a fixed order (deterministic, independent of timing).  Semantics (minibatch order,
// 1/B scaling, L2-in-gradient weight decay, bias correction, state persistence) are the reference's.
#include <cooperative_groups.h>

#include "kernels.h"

namespace cg = cooperative_groups;

namespace mjb {

namespace {

constexpr int MT_ = 512;       // threads per CTA
constexpr int BP = 68;         // pitch of [feature][64 rows] buffers (floats); 68 % 32 == 4 -> conflict-free LDS.128
constexpr int NB = 64;         // minibatch rows

struct MpArgs {
    int K, H1, H2, u1, u2, obs_dim, steps;
    const float* feat;                       // [N][K] fp32 features (mlp_baseline.py:36-58), built once per fit
    const float* ret32;                      // [N] float32(returns)
    const int* perm;
    float lr, reg, beta1, beta2, eps;
    float* w; float* m; float* v;            // natural nn.Sequential layout, global
    const float2* consts;                    // per-step {sqrt(1-b2^t), -lr/(1-b1^t)}
    long long* prof;
};

struct AdamP { float one_m_b1, b2, one_m_b2, bc2_sqrt, eps, neg_step, reg; };

__device__ __forceinline__ float adam_apply(float g, float w, float* m, float* v, const AdamP& c) {
    g = fmaf(c.reg, w, g);
    const float mn = *m + c.one_m_b1 * (g - *m);
    const float vn = fmaf(c.one_m_b2 * g, g, *v * c.b2);
    *m = mn; *v = vn;
    return fmaf(c.neg_step, mn / (sqrtf(vn) / c.bc2_sqrt + c.eps), w);
}

// remote (distributed shared memory) helpers: addresses are 32-bit shared::cluster addresses
__device__ __forceinline__ uint32_t map_cluster(const void* local_smem, int cta) {
    uint32_t l = (uint32_t)__cvta_generic_to_shared(local_smem), r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(l), "r"(cta));
    return r;
}
// asynchronous 16-byte store into another CTA's shared memory; completion is counted (16 bytes) on that CTA's mbarrier
__device__ __forceinline__ void st_async_v4(uint32_t raddr, float4 v, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];\n"
                 ::"r"(raddr), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)),
                   "r"(__float_as_uint(v.w)), "r"(rbar) : "memory");
}
__device__ __forceinline__ void st_async_f32(uint32_t raddr, float v, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];\n"
                 ::"r"(raddr), "r"(__float_as_uint(v)), "r"(rbar) : "memory");
}
__device__ __forceinline__ void bar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arm(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait_cluster(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D_%=;\n\tbra W_%=;\n\t"
        "D_%=:\n\t}\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}

// C CTAs, hidden width H (both layers), u = H/C owned units per CTA.  K (input features) is a run-time value.
template <int C, int H>
__global__ void __launch_bounds__(MT_, 1) vf_fit_mp_kernel(const MpArgs a) {
    constexpr int U = H / C;                      // owned units per layer
    constexpr int W2P = H + 4;                    // pitch of owned W2 rows
    constexpr int ITEMS = U * 16;                 // (unit, row-quad) work items of an owned slice
    constexpr int KS = MT_ / ITEMS;               // reduction split of the layer-2 slice
    constexpr int KR = H / KS;
    constexpr int NG = MT_ / H;                   // thread groups along the owned rows in the W2 weight gradient
    constexpr int PER = (U + NG - 1) / NG;
    static_assert(ITEMS * KS == MT_ && KR * KS == H && PER <= 4 && NG * H == MT_, "shape not covered");
    uint32_t c;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(c));
    const int K = a.K;
    const int tid = threadIdx.x;
    extern __shared__ __align__(16) float sm[];
    // ---- owned parameters + Adam state (persist for the whole epoch) ----
    float* W1o = sm;                              // [U][K]
    float* b1o = W1o + U * K;                     // [U]
    float* W2o = b1o + U;                         // [U][W2P]
    float* b2o = W2o + U * W2P;                   // [U]
    float* W3o = b2o + U;                         // [U]
    float* b3r = W3o + U;                         // [1] replicated in every CTA (identical updates)
    const int np = U * K + U + U * W2P + 2 * U + 1;
    const int npp = round_up(np, 4);
    float* Mo = sm + npp;                         // Adam exp_avg, same indexing as the parameter block
    float* Vo = Mo + npp;                         // Adam exp_avg_sq
    // ---- per-step buffers; the three exchange buffers are double-buffered by step parity ----
    float* xT = Vo + npp;                         // [K][BP]        minibatch features (all 64 rows)
    float* h1o = xT + K * BP;                     // [U][BP]        owned slice of h1
    float* h2o = h1o + U * BP;                    // [U][BP]        owned slice of h2, then delta2
    float* d1o = h2o + U * BP;                    // [U][BP]        owned slice of delta1
    float* red = d1o + U * BP;                    // [MT_][4]       k-split partials
    float* tv = red + MT_ * 4;                    // [NB] targets
    float* dy = tv + NB;                          // [NB]
    float* h1f = dy + NB;                         // [2][H][BP]     full h1 (every CTA's slice lands here)
    float* dg = h1f + 2 * H * BP;                 // [2][C][U][BP]  partial dgrad slices from every CTA
    float* yp = dg + 2 * C * U * BP;              // [2][C][NB]     partial outputs from every CTA
    __shared__ AdamP s_c;
    __shared__ __align__(8) uint64_t bars[3][2];  // [exchange][step parity]

    const int oW1 = 0, ob1 = H * K, oW2 = ob1 + H, ob2 = oW2 + H * H, oW3 = ob2 + H, ob3 = oW3 + H;
    auto nat_index = [&](int p) -> int {          // owned-block index -> natural flat index
        if (p < U * K) return oW1 + ((int)c * U + p / K) * K + p % K;
        p -= U * K;
        if (p < U) return ob1 + (int)c * U + p;
        p -= U;
        if (p < U * W2P) { const int n = p / W2P, k = p % W2P; return k < H ? oW2 + ((int)c * U + n) * H + k : -1; }
        p -= U * W2P;
        if (p < U) return ob2 + (int)c * U + p;
        p -= U;
        if (p < U) return oW3 + (int)c * U + p;
        return ob3;
    };
    for (int p = tid; p < np; p += MT_) {
        const int j = nat_index(p);
        W1o[p] = j >= 0 ? a.w[j] : 0.0f;
        Mo[p] = j >= 0 ? a.m[j] : 0.0f;
        Vo[p] = j >= 0 ? a.v[j] : 0.0f;
    }
    const int nbuf = (K + 3 * U) * BP + MT_ * 4 + 2 * NB + 2 * H * BP + 2 * C * U * BP + 2 * C * NB;
    for (int i = tid; i < nbuf; i += MT_) xT[i] = 0.0f;
    if (tid < 6) bar_init(&bars[tid >> 1][tid & 1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    __syncthreads();

    // ---- minibatch gather pipeline: indices two steps ahead, feature rows one step ahead, in registers ----
    constexpr int NPF = 4;                         // NB*K <= NPF*MT_  (K <= 32)
    float pre_x[NPF];
    float pre_t = 0.f;
    int r_nxt[NPF], rt_nxt = 0;
    int e_row[NPF], e_col[NPF];                    // element slot -> (minibatch row, feature), fixed for the whole epoch
#pragma unroll
    for (int u = 0; u < NPF; ++u) { const int e = tid + MT_ * u; e_row[u] = (e < NB * K) ? e / K : -1; e_col[u] = e % K; }
    auto load_idx = [&](int s) {
        const int* pidx = a.perm + (size_t)s * NB;
#pragma unroll
        for (int u = 0; u < NPF; ++u) r_nxt[u] = (e_row[u] >= 0) ? pidx[e_row[u]] : 0;
        rt_nxt = (tid < NB) ? pidx[tid] : 0;
    };
    auto load_vals = [&]() {
#pragma unroll
        for (int u = 0; u < NPF; ++u)
            if (e_row[u] >= 0) pre_x[u] = a.feat[(size_t)r_nxt[u] * K + e_col[u]];
        if (tid < NB) pre_t = a.ret32[rt_nxt];
    };
    auto commit = [&]() {
#pragma unroll
        for (int u = 0; u < NPF; ++u)
            if (e_row[u] >= 0) xT[e_col[u] * BP + e_row[u]] = pre_x[u];
        if (tid < NB) tv[tid] = pre_t;
    };
    load_idx(0);
    load_vals();
    commit();
    if (a.steps > 1) load_idx(1);
    float2 cc_next = make_float2(1.f, 0.f);
    if (tid == 0) cc_next = a.consts[0];
    cluster_sync_relacq();                         // barriers initialised and buffers zeroed everywhere before any remote store

    // phase profiler: accumulates in shared memory (a global read-modify-write per phase would itself cost ~1k cycles)
    __shared__ long long s_prof[16];
    if (tid < 16) s_prof[tid] = 0;
    long long t_last = clock64();
#define MP_PROF(i) do { if (a.prof && tid == 0 && c == 0) { const long long _t = clock64(); s_prof[i] += _t - t_last; t_last = _t; } } while (0)
    for (int s = 0; s < a.steps; ++s) {
        const int par = s & 1;
        const uint32_t ph = (uint32_t)(s >> 1) & 1u;          // phase parity of the barriers of this step parity
        float* h1f_s = h1f + par * H * BP;
        float* dg_s = dg + par * C * U * BP;
        float* yp_s = yp + par * C * NB;
        if (tid == 0) {
            const float2 cc = cc_next;
            s_c.one_m_b1 = (float)(1.0 - (double)a.beta1); s_c.b2 = a.beta2; s_c.one_m_b2 = (float)(1.0 - (double)a.beta2);
            s_c.bc2_sqrt = cc.x; s_c.eps = a.eps; s_c.neg_step = cc.y; s_c.reg = a.reg;
            // arm this step's three hand-off barriers with the bytes every CTA (incl. myself) will deliver
            bar_arm(&bars[0][par], C * U * NB * 4);
            bar_arm(&bars[1][par], C * NB * 4);
            bar_arm(&bars[2][par], C * U * NB * 4);
        }
        // ---- P1: owned slice of layer 1 ----
        if (tid < ITEMS) {
            const int n = tid % U, q = tid / U;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            const float* wr = W1o + n * K;
#pragma unroll 4
            for (int k = 0; k < K; ++k) {
                const float w = wr[k];
                const float4 x = *reinterpret_cast<const float4*>(xT + k * BP + 4 * q);
                acc.x = fmaf(x.x, w, acc.x); acc.y = fmaf(x.y, w, acc.y); acc.z = fmaf(x.z, w, acc.z); acc.w = fmaf(x.w, w, acc.w);
            }
            const float bb = b1o[n];
            acc.x = fmaxf(acc.x + bb, 0.f); acc.y = fmaxf(acc.y + bb, 0.f); acc.z = fmaxf(acc.z + bb, 0.f); acc.w = fmaxf(acc.w + bb, 0.f);
            *reinterpret_cast<float4*>(h1o + n * BP + 4 * q) = acc;
        }
        __syncthreads();
        // ---- E1: my h1 slice -> every CTA's full h1 (asynchronous remote stores, counted on the receiver's barrier) ----
#pragma unroll
        for (int i = tid; i < ITEMS * C; i += MT_) {
            const int item = i % ITEMS, r = i / ITEMS;
            const int n = item % U, q = item / U;
            const float4 v = *reinterpret_cast<const float4*>(h1o + n * BP + 4 * q);
            st_async_v4(map_cluster(h1f_s + ((int)c * U + n) * BP + 4 * q, r), v, map_cluster(&bars[0][par], r));
        }
        MP_PROF(0);
        bar_wait_cluster(&bars[0][par], ph);       // full h1 has arrived
        MP_PROF(1);
        // ---- P2: owned slice of layer 2 (reduction over H split KS ways) ----
        {
            const int item = tid % ITEMS, ks = tid / ITEMS;
            const int n = item % U, q = item / U;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            const float* wr = W2o + n * W2P + ks * KR;
            const float* hp = h1f_s + (ks * KR) * BP + 4 * q;
#pragma unroll 8
            for (int i = 0; i < KR; ++i) {
                const float w = wr[i];
                const float4 x = *reinterpret_cast<const float4*>(hp + i * BP);
                acc.x = fmaf(x.x, w, acc.x); acc.y = fmaf(x.y, w, acc.y); acc.z = fmaf(x.z, w, acc.z); acc.w = fmaf(x.w, w, acc.w);
            }
            *reinterpret_cast<float4*>(red + tid * 4) = acc;
            __syncthreads();
            if (tid < ITEMS) {
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int k2 = 0; k2 < KS; ++k2) {
                    const float4 r = *reinterpret_cast<const float4*>(red + (k2 * ITEMS + tid) * 4);
                    t.x += r.x; t.y += r.y; t.z += r.z; t.w += r.w;
                }
                const float bb = b2o[n];
                t.x = fmaxf(t.x + bb, 0.f); t.y = fmaxf(t.y + bb, 0.f); t.z = fmaxf(t.z + bb, 0.f); t.w = fmaxf(t.w + bb, 0.f);
                *reinterpret_cast<float4*>(h2o + n * BP + 4 * q) = t;
            }
        }
        __syncthreads();
        // ---- E2: partial outputs of my units -> everybody ----
        for (int i = tid; i < NB * C; i += MT_) {
            const int b = i % NB, r = i / NB;
            float t = 0.0f;
#pragma unroll
            for (int n = 0; n < U; ++n) t = fmaf(h2o[n * BP + b], W3o[n], t);
            st_async_f32(map_cluster(yp_s + (int)c * NB + b, r), t, map_cluster(&bars[1][par], r));
        }
        MP_PROF(2);
        bar_wait_cluster(&bars[1][par], ph);
        MP_PROF(3);
        if (tid < NB) {
            float y = 0.0f;
#pragma unroll
            for (int r = 0; r < C; ++r) y += yp_s[r * NB + tid];
            dy[tid] = 2.0f * ((y + b3r[0]) - tv[tid]) / (float)NB;
        }
        __syncthreads();
        const AdamP ck = s_c;
        // ---- P3: small gradients (one warp per owned unit, shuffle reductions), delta2 (in place of h2) ----
        // warp w < U: gW3[w] = sum_b dy[b] h2[w][b] and, after the ReLU mask, gb2[w] = sum_b delta2[w][b]; warp 0 also gb3
        float g_w3 = 0.0f, g_b2 = 0.0f, g_b3 = 0.0f;
        {
            const int w = tid >> 5, lane = tid & 31;
            if (w < U) {
                const float w3 = W3o[w];
                const float h0 = h2o[w * BP + lane], h1v = h2o[w * BP + 32 + lane];
                const float d0 = dy[lane], d1 = dy[32 + lane];
                g_w3 = warp_sum(fmaf(d0, h0, d1 * h1v));
                const float e0 = h0 > 0.f ? d0 * w3 : 0.f, e1 = h1v > 0.f ? d1 * w3 : 0.f;
                g_b2 = warp_sum(e0 + e1);
                __syncwarp();
                h2o[w * BP + lane] = e0;                    // delta2 in place of h2 (this warp owns the whole row)
                h2o[w * BP + 32 + lane] = e1;
            }
            if (w == 0) g_b3 = warp_sum(dy[lane] + dy[32 + lane]);
        }
        __syncthreads();
        MP_PROF(8);
        // ---- partial dgrad over my units, scattered to the owners of each h1 unit (E3) ----
        // Warp w handles the 8 consecutive h1 units k = 8w..8w+7 (one owner CTA per warp, so every st.async warp
        // instruction has a single destination): lane -> (k = 8w + lane%8, row-quad q = lane/8 + 4j, j < 4).
        for (int kb = (tid >> 5) * 8; kb < H; kb += (MT_ >> 5) * 8) {
            const int k = kb + (tid & 7), ql = (tid & 31) >> 3;
            float4 acc[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int n = 0; n < U; ++n) {
                const float w = W2o[n * W2P + k];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 d = *reinterpret_cast<const float4*>(h2o + n * BP + 4 * (ql + 4 * j));
                    acc[j].x = fmaf(d.x, w, acc[j].x); acc[j].y = fmaf(d.y, w, acc[j].y);
                    acc[j].z = fmaf(d.z, w, acc[j].z); acc[j].w = fmaf(d.w, w, acc[j].w);
                }
            }
            const int r = k / U;
            const uint32_t dst = map_cluster(dg_s + ((int)c * U + k % U) * BP, r), rb = map_cluster(&bars[2][par], r);
#pragma unroll
            for (int j = 0; j < 4; ++j) st_async_v4(dst + 16 * (ql + 4 * j), acc[j], rb);
        }
        MP_PROF(9);
        // ---- wgrad of my W2 rows: g[n][k] = sum_b delta2[n][b] h1[k][b]  (kept in registers until the update) ----
        float gw2[4] = {0.f, 0.f, 0.f, 0.f};
        {
            const int k = tid % H, ng = tid / H;
#pragma unroll 4
            for (int q = 0; q < 16; ++q) {
                const float4 h = *reinterpret_cast<const float4*>(h1f_s + k * BP + 4 * q);
#pragma unroll
                for (int j = 0; j < PER; ++j) {
                    const int n = ng * PER + j;
                    if (n < U) {
                        const float4 d = *reinterpret_cast<const float4*>(h2o + n * BP + 4 * q);
                        gw2[j] = fmaf(d.x, h.x, gw2[j]); gw2[j] = fmaf(d.y, h.y, gw2[j]);
                        gw2[j] = fmaf(d.z, h.z, gw2[j]); gw2[j] = fmaf(d.w, h.w, gw2[j]);
                    }
                }
            }
        }
        MP_PROF(10);
        // Gather for the next step is issued here: nothing below waits on a fence that would stall on these loads.
        if (s + 1 < a.steps) load_vals();
        if (s + 2 < a.steps) load_idx(s + 2);
        if (tid == 0 && s + 1 < a.steps) cc_next = a.consts[s + 1];
        MP_PROF(4);
        bar_wait_cluster(&bars[2][par], ph);       // every partial dgrad slice destined to me has arrived
        MP_PROF(5);
        // ---- P4: delta1 of my units (fixed-order sum over the C sources) ----
        if (tid < ITEMS) {
            const int n = tid % U, q = tid / U;
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 0; r < C; ++r) {
                const float4 d = *reinterpret_cast<const float4*>(dg_s + (r * U + n) * BP + 4 * q);
                t.x += d.x; t.y += d.y; t.z += d.z; t.w += d.w;
            }
            const float4 h = *reinterpret_cast<const float4*>(h1o + n * BP + 4 * q);
            t.x = h.x > 0.f ? t.x : 0.f; t.y = h.y > 0.f ? t.y : 0.f; t.z = h.z > 0.f ? t.z : 0.f; t.w = h.w > 0.f ? t.w : 0.f;
            *reinterpret_cast<float4*>(d1o + n * BP + 4 * q) = t;
        }
        __syncthreads();                           // also: every thread is past its reads of W2o (dgrad) -> safe to update
        // ---- P5: Adam on everything I own ----
        {
            const int k = tid % H, ng = tid / H;
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                const int n = ng * PER + j;
                if (n < U) {
                    float* p = W2o + n * W2P + k;
                    *p = adam_apply(gw2[j], *p, Mo + (p - sm), Vo + (p - sm), ck);
                }
            }
        }
        for (int o = tid; o < U * K; o += MT_) {    // W1 rows
            const int n = o % U, k = o / U;
            float g = 0.0f;
#pragma unroll 4
            for (int q = 0; q < 16; ++q) {
                const float4 d = *reinterpret_cast<const float4*>(d1o + n * BP + 4 * q);
                const float4 x = *reinterpret_cast<const float4*>(xT + k * BP + 4 * q);
                g = fmaf(d.x, x.x, g); g = fmaf(d.y, x.y, g); g = fmaf(d.z, x.z, g); g = fmaf(d.w, x.w, g);
            }
            float* p = W1o + n * K + k;
            *p = adam_apply(g, *p, Mo + (p - sm), Vo + (p - sm), ck);
        }
        {                                          // b1: one warp per owned unit (taken from the far end of the block:
            const int w = (MT_ >> 5) - 1 - (tid >> 5), lane = tid & 31;   // the W1 loop keeps the low warps busy)
            if (w < U) {
                const float g = warp_sum(d1o[w * BP + lane] + d1o[w * BP + 32 + lane]);
                if (lane == 0) { float* p = b1o + w; *p = adam_apply(g, *p, Mo + (p - sm), Vo + (p - sm), ck); }
            }
        }
        if ((tid & 31) == 0) {                     // lane 0 of warp w holds the reduced small gradients of unit w
            const int w = tid >> 5;
            if (w < U) {
                float* p = W3o + w; *p = adam_apply(g_w3, *p, Mo + (p - sm), Vo + (p - sm), ck);
                float* pb = b2o + w; *pb = adam_apply(g_b2, *pb, Mo + (pb - sm), Vo + (pb - sm), ck);
            }
            if (w == 0) { float* p = b3r; *p = adam_apply(g_b3, *p, Mo + (p - sm), Vo + (p - sm), ck); }
        }
        __syncthreads();                           // W1 wgrad reads of xT are done before the next minibatch lands
        if (s + 1 < a.steps) commit();
        __syncthreads();
        MP_PROF(6);
    }
    if (a.prof && c == 0 && tid < 16) a.prof[tid] += s_prof[tid];
    // ---- write the owned parameters / moments back to the natural layout ----
    for (int p = tid; p < np; p += MT_) {
        const int j = nat_index(p);
        if (j >= 0 && (j != ob3 || c == 0)) { a.w[j] = W1o[p]; a.m[j] = Mo[p]; a.v[j] = Vo[p]; }
    }
    cluster_sync_relacq();                         // nobody exits while a peer might still be storing into it
}

// fp32 feature matrix of the whole batch, in the reference's dtypes: fp64 feature map, then .astype(float32)
__global__ void vf_features_kernel(const float* __restrict__ obs, const int* __restrict__ tstep,
                                   const double* __restrict__ returns, long long n, int obs_dim, int K,
                                   float* __restrict__ feat, float* __restrict__ ret32) {
    const long long total = n * K;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / K;
        const int k = (int)(i - r * K);
        float val;
        if (k < obs_dim) {
            double x = (double)obs[r * obs_dim + k];
            x = fmin(fmax(x, -10.0), 10.0) / 10.0;
            val = (float)x;
        } else {
            const double tau = (double)tstep[r] / 1000.0;
            double p = tau;
            for (int q = obs_dim; q < k; ++q) p *= tau;
            val = (float)p;
        }
        feat[i] = val;
        if (k == 0) ret32[r] = (float)returns[r];
    }
}

__global__ void mp_adam_consts_kernel(float2* out, int steps, long long step0, float lr, float beta1, float beta2) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= steps) return;
    const double t = (double)(step0 + s + 1);
    const double bc1 = 1.0 - pow((double)beta1, t), bc2 = 1.0 - pow((double)beta2, t);
    out[s] = make_float2((float)sqrt(bc2), (float)(-((double)lr / bc1)));
}

size_t mp_smem_bytes(int K, int H1, int H2, int C) {
    const int H = H1, U = H / C, W2P = H + 4;
    const int np = U * K + U + U * W2P + 2 * U + 1, npp = round_up(np, 4);
    const size_t fl = 3 * (size_t)npp + (size_t)(K + 3 * U) * BP + MT_ * 4 + 2 * NB + 2 * (size_t)H * BP + 2 * (size_t)C * U * BP +
                      2 * (size_t)C * NB + 64;
    return fl * 4;
}

template <int C, int H>
cudaError_t launch_mp(const MpArgs& a, size_t smem, cudaStream_t s) {
    auto kern = vf_fit_mp_kernel<C, H>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (C > 8) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(C); cfg.blockDim = dim3(MT_); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, a);
}

long long* g_mp_prof = nullptr;

}  // namespace

void vf_mp_set_prof(long long* dev16) { g_mp_prof = dev16; }

bool vf_mp_supported(int K, int H1, int H2, int batch, int C) {
    if (batch != NB || (C != 8 && C != 16)) return false;
    if (H1 != H2 || (H1 != 128 && H1 != 64)) return false;    // instantiated widths (the reference default is 128)
    if (NB * K > 4 * MT_) return false;                       // gather prefetch registers (K <= 32)
    return mp_smem_bytes(K, H1, H2, C) <= 200 * 1024;
}

// Build the fp32 feature matrix + fp32 targets once per fit (all epochs reuse them).
cudaError_t vf_build_features(const VfFitArgs& v, float* feat, float* ret32, cudaStream_t s) {
    vf_features_kernel<<<148 * 8, 256, 0, s>>>(v.obs, v.tstep, v.returns, v.n, v.obs_dim, v.K, feat, ret32);
    return cudaGetLastError();
}

cudaError_t launch_vf_fit_mp(const VfFitArgs& v, const float* feat, const float* ret32, float2* consts, int C, cudaStream_t s) {
    mp_adam_consts_kernel<<<(v.steps + 255) / 256, 256, 0, s>>>(consts, v.steps, v.step0, v.lr, v.beta1, v.beta2);
    MpArgs a;
    a.K = v.K; a.H1 = v.H1; a.H2 = v.H2; a.u1 = v.H1 / C; a.u2 = v.H2 / C; a.obs_dim = v.obs_dim; a.steps = v.steps;
    a.feat = feat; a.ret32 = ret32; a.perm = v.perm;
    a.lr = v.lr; a.reg = v.reg; a.beta1 = v.beta1; a.beta2 = v.beta2; a.eps = v.eps;
    a.w = v.w; a.m = v.m; a.v = v.v; a.consts = consts; a.prof = g_mp_prof;
    const size_t smem = mp_smem_bytes(v.K, v.H1, v.H2, C);
    cudaError_t e = cudaErrorInvalidValue;
    if (v.H1 == 128) e = (C == 8) ? launch_mp<8, 128>(a, smem, s) : launch_mp<16, 128>(a, smem, s);
    else if (v.H1 == 64) e = (C == 8) ? launch_mp<8, 64>(a, smem, s) : launch_mp<16, 64>(a, smem, s);
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace mjb

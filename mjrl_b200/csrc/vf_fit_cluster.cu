// MLPBaseline.fit, cluster version (baselines/mlp_baseline.py:61-95 + utils/optimize_model.py:7-36).
//
// The Adam chain is sequential across minibatches, so the only parallelism inside one step is over the 64
// rows of the minibatch and over the parameters.  One thread-block CLUSTER of C CTAs (C SMs) runs the whole
// epoch as a persistent kernel:
//   * every CTA keeps the full weight set (78 KB for 21->128->128->1) in shared memory, in a k-major padded
//     layout so forward, dgrad and wgrad all read it conflict-free;
//   * the minibatch is split over the CTAs (64/C rows each): forward + backward of those rows is local;
//   * the per-CTA gradient partials are exchanged through L2 (st.global.cg / ld.global.cg) between two
//     hardware cluster barriers; each CTA owns 1/C of the parameters (Adam moments live in its shared
//     memory), sums the C partials for its slice, applies torch.optim.Adam's update and publishes the new
//     weights, which every CTA then reloads with one flat coalesced copy;
//   * the next minibatch's rows (random rows of the 68 MB batch: HBM latency) are prefetched into registers
//     while the current step computes.
// Two cluster barriers per step; no grid-wide sync, no atomics.  Minibatch order, 1/B scaling, L2-in-gradient
// weight decay and bias-correction follow the reference exactly; only the summation order inside the
// minibatch gradient differs (C partial sums), which is fp32-rounding-level.
#include <cooperative_groups.h>

#include "kernels.h"

namespace cg = cooperative_groups;

namespace mjb {

namespace {

constexpr int CT = 512;        // threads per CTA (16 warps: the per-step GEMVs are latency-bound)
constexpr int AP = 12;         // row pitch (floats) of feature-major activations: up to 8 local rows, conflict-free LDS.128

struct KL {                    // kernel layout of the value net parameters (floats)
    int K, H1, H2, P1, P2;
    int oW1T, ob1, oW2T, ob2, oW3, ob3, total, per;
};

__host__ __device__ inline KL make_kl(int K, int H1, int H2, int C) {
    KL L;
    L.K = K; L.H1 = H1; L.H2 = H2; L.P1 = H1 + 1; L.P2 = H2 + 1;
    int o = 0;
    L.oW1T = o; o += K * L.P1;
    L.ob1 = o;  o += H1;
    L.oW2T = o; o += H1 * L.P2;
    L.ob2 = o;  o += H2;
    L.oW3 = o;  o += H2;
    L.ob3 = o;  o += 1;
    L.total = round_up(o, 4 * C);
    L.per = L.total / C;
    return L;
}

// natural (nn.Sequential.parameters()) index -> kernel-layout index
__device__ __forceinline__ int nat_to_kl(int i, const KL& L) {
    const int nW1 = L.H1 * L.K, nW2 = L.H2 * L.H1;
    if (i < nW1) { const int n = i / L.K, k = i % L.K; return L.oW1T + k * L.P1 + n; }
    i -= nW1;
    if (i < L.H1) return L.ob1 + i;
    i -= L.H1;
    if (i < nW2) { const int n = i / L.H1, k = i % L.H1; return L.oW2T + k * L.P2 + n; }
    i -= nW2;
    if (i < L.H2) return L.ob2 + i;
    i -= L.H2;
    if (i < L.H2) return L.oW3 + i;
    return L.ob3;
}

__global__ void vf_relayout_kernel(float* nat_w, float* nat_m, float* nat_v, float* kw, float* km, float* kv, KL L,
                                   int n_nat, int to_kernel) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nat) return;
    const int j = nat_to_kl(i, L);
    if (to_kernel) { kw[j] = nat_w[i]; km[j] = nat_m[i]; kv[j] = nat_v[i]; }
    else { nat_w[i] = kw[j]; nat_m[i] = km[j]; nat_v[i] = kv[j]; }
}

struct AdamK { float one_m_b1, b2, one_m_b2, bc2_sqrt, eps, neg_step, reg; };

// per-step bias-correction constants {sqrt(1-b2^t), -lr/(1-b1^t)} computed off the critical path (fp64 pow)
__global__ void vf_adam_consts_kernel(float2* out, int steps, long long step0, float lr, float beta1, float beta2) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= steps) return;
    const double t = (double)(step0 + s + 1);
    const double bc1 = 1.0 - pow((double)beta1, t), bc2 = 1.0 - pow((double)beta2, t);
    out[s] = make_float2((float)sqrt(bc2), (float)(-((double)lr / bc1)));
}

struct ClArgs {
    KL L;
    int obs_dim, batch, steps;
    const float* obs; const int* tstep; const double* returns; const int* perm;
    float lr, reg, beta1, beta2, eps; long long step0;
    float* kw; float* km; float* kv;      // kernel-layout weights / moments (global)
    float* gpart;                         // [C][total] gradient partials
    long long* prof;                      // optional [16] per-phase cycle counters (CTA 0, thread 0)
    const float2* consts;                 // [steps] Adam bias-correction constants
    float* red;                           // unused (reserved)
};

__device__ __forceinline__ float feature_value(const ClArgs& a, long long r, int k) {
    if (k < a.obs_dim) {
        double x = (double)a.obs[r * a.obs_dim + k];
        x = fmin(fmax(x, -10.0), 10.0) / 10.0;
        return (float)x;
    }
    const double tau = (double)a.tstep[r] / 1000.0;
    double p = tau;
    for (int q = a.obs_dim; q < k; ++q) p *= tau;
    return (float)p;
}

template <int C>
__global__ void __launch_bounds__(CT, 1) vf_fit_cluster_kernel(const ClArgs a) {
    cg::cluster_group cluster = cg::this_cluster();
    const int c = (int)cluster.block_rank();
    const KL L = a.L;
    const int K = L.K, H1 = L.H1, H2 = L.H2, P1 = L.P1, P2 = L.P2;
    const int BL = a.batch / C;                        // local rows
    constexpr int NQ = (C == 8) ? 2 : 1;               // quads of local rows (batch 64)
    extern __shared__ __align__(16) float sm[];
    float* sw = sm;                                   // [total] weights, kernel layout
    float* s_m = sw + L.total;                        // [per] Adam exp_avg of the owned slice
    float* s_v = s_m + L.per;                         // [per] Adam exp_avg_sq
    float* red = s_v + L.per;                         // [CT][8] k-split partial sums
    float* xT = red + CT * 8;                         // [K][AP]
    float* h1T = xT + K * AP;                         // [H1][AP]
    float* h2T = h1T + H1 * AP;                       // [H2][AP]  (becomes delta2)
    float* d1T = h2T + H2 * AP;                       // [H1][AP]
    float* tv = d1T + H1 * AP;                        // [8] targets
    float* dy = tv + 8;                               // [8]
    __shared__ AdamK s_c;
    const int tid = threadIdx.x;
    const int own0 = c * L.per;
    float* gp = a.gpart + (size_t)c * L.total;

    for (int i = tid; i < L.total; i += CT) sw[i] = __ldcg(a.kw + i);
    for (int i = tid; i < L.per; i += CT) { s_m[i] = a.km[own0 + i]; s_v[i] = a.kv[own0 + i]; }
    for (int i = tid; i < (K + 2 * H1 + H2) * AP + 16; i += CT) xT[i] = 0.0f;
    __syncthreads();

    // Two-deep software pipeline for the minibatch gather: row indices are loaded two steps ahead (L2 latency),
    // the rows themselves one step ahead (HBM latency), both into registers, so no step stalls on a dependent load.
    // element slot e = tid of the local [BL][K] feature block (BL*K <= CT).
    float pre_x = 0.f;
    double pre_t64 = 0.0;
    int r_nxt = 0, rt_nxt = 0;
    auto load_idx = [&](int s) {
        const int* pidx = a.perm + (size_t)s * a.batch + c * BL;
        r_nxt = (tid < BL * K) ? pidx[tid / K] : 0;
        rt_nxt = (tid < BL) ? pidx[tid] : 0;
    };
    auto load_vals = [&]() {                            // raw loads only: nothing here consumes the value
        if (tid < BL * K) {
            const int k = tid % K;
            const long long r = r_nxt;
            pre_x = (k < a.obs_dim) ? a.obs[r * a.obs_dim + k] : __int_as_float(a.tstep[r]);
        }
        if (tid < BL) pre_t64 = a.returns[rt_nxt];
    };
    auto commit_prefetch = [&]() {                       // feature map of mlp_baseline.py:36-58 applied here
        if (tid < BL * K) {
            const int k = tid % K;
            float val;
            if (k < a.obs_dim) {
                double x = (double)pre_x;
                x = fmin(fmax(x, -10.0), 10.0) / 10.0;
                val = (float)x;
            } else {
                const double tau = (double)__float_as_int(pre_x) / 1000.0;
                double p = tau;
                for (int q = a.obs_dim; q < k; ++q) p *= tau;
                val = (float)p;
            }
            xT[k * AP + tid / K] = val;
        }
        if (tid < BL) tv[tid] = (float)pre_t64;
    };
    load_idx(0);
    load_vals();
    commit_prefetch();
    if (a.steps > 1) load_idx(1);
    __syncthreads();

    // k-split GEMV helper: partial[t] = sum_{r in my range} in[r][0..7] * w(r), stored to red[t*8..]
    // thread t -> (o = t % NO: output unit, ks = t / NO: reduction slice)
    float2 cc_next = make_float2(1.f, 0.f);
    if (tid == 0) cc_next = a.consts[0];
    long long t_last = clock64();
#define MJB_PROF(i) do { if (a.prof && tid == 0 && c == 0) { const long long _t = clock64(); a.prof[i] += _t - t_last; t_last = _t; } } while (0)
    for (int s = 0; s < a.steps; ++s) {
        if (tid == 0) {
            const float2 cc = cc_next;
            s_c.one_m_b1 = (float)(1.0 - (double)a.beta1);
            s_c.b2 = a.beta2;
            s_c.one_m_b2 = (float)(1.0 - (double)a.beta2);
            s_c.bc2_sqrt = cc.x;
            s_c.eps = a.eps;
            s_c.neg_step = cc.y;
            s_c.reg = a.reg;
        }
        // ---- forward layer 1: h1[n][q] = relu(sum_k x[k][q] W1T[k][n] + b1[n]) ----
        for (int o = tid; o < H1 * NQ; o += CT) {
            const int n = o % H1, q = o / H1;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
            for (int k = 0; k < K; ++k) {
                const float w = sw[L.oW1T + k * P1 + n];
                const float4 x = *reinterpret_cast<const float4*>(xT + k * AP + 4 * q);
                acc.x = fmaf(x.x, w, acc.x); acc.y = fmaf(x.y, w, acc.y); acc.z = fmaf(x.z, w, acc.z); acc.w = fmaf(x.w, w, acc.w);
            }
            const float bb = sw[L.ob1 + n];
            acc.x = fmaxf(acc.x + bb, 0.f); acc.y = fmaxf(acc.y + bb, 0.f); acc.z = fmaxf(acc.z + bb, 0.f); acc.w = fmaxf(acc.w + bb, 0.f);
            *reinterpret_cast<float4*>(h1T + n * AP + 4 * q) = acc;
        }
        __syncthreads();
        MJB_PROF(0);
        // ---- forward layer 2, reduction split over CT/H2 thread groups ----
        {
            const int n = tid % H2, ks = tid / H2, KS = CT / H2, kr = H1 / KS;
            float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
            const float* wp = sw + L.oW2T + (ks * kr) * P2 + n;
            const float* hp = h1T + (ks * kr) * AP;
#pragma unroll 8
            for (int i = 0; i < kr; ++i) {
                const float w = wp[i * P2];
                const float4 x0 = *reinterpret_cast<const float4*>(hp + i * AP);
                a0.x = fmaf(x0.x, w, a0.x); a0.y = fmaf(x0.y, w, a0.y); a0.z = fmaf(x0.z, w, a0.z); a0.w = fmaf(x0.w, w, a0.w);
                if (NQ == 2) {
                    const float4 x1 = *reinterpret_cast<const float4*>(hp + i * AP + 4);
                    a1.x = fmaf(x1.x, w, a1.x); a1.y = fmaf(x1.y, w, a1.y); a1.z = fmaf(x1.z, w, a1.z); a1.w = fmaf(x1.w, w, a1.w);
                }
            }
            *reinterpret_cast<float4*>(red + tid * 8) = a0;
            *reinterpret_cast<float4*>(red + tid * 8 + 4) = a1;
        }
        __syncthreads();
        for (int o = tid; o < H2 * NQ; o += CT) {
            const int n = o % H2, q = o / H2, KS = CT / H2;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int ks = 0; ks < KS; ++ks) {
                const float4 r = *reinterpret_cast<const float4*>(red + (ks * H2 + n) * 8 + 4 * q);
                acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w;
            }
            const float bb = sw[L.ob2 + n];
            acc.x = fmaxf(acc.x + bb, 0.f); acc.y = fmaxf(acc.y + bb, 0.f); acc.z = fmaxf(acc.z + bb, 0.f); acc.w = fmaxf(acc.w + bb, 0.f);
            *reinterpret_cast<float4*>(h2T + n * AP + 4 * q) = acc;
        }
        __syncthreads();
        MJB_PROF(1);
        // ---- output + loss gradient: one warp per local row ----
        {
            const int b = tid >> 5, lane = tid & 31;
            if (b < BL) {
                float t = 0.0f;
                for (int n = lane; n < H2; n += 32) t = fmaf(h2T[n * AP + b], sw[L.oW3 + n], t);
                t = warp_sum(t);
                if (lane == 0) dy[b] = 2.0f * ((t + sw[L.ob3]) - tv[b]) / (float)a.batch;
            }
        }
        __syncthreads();
        MJB_PROF(2);
        const AdamK ck = s_c;
        // ---- W3 / b3 partial gradients, then delta2 in place of h2 ----
        float g3 = 0.0f;
        if (tid < H2) { for (int b = 0; b < BL; ++b) g3 = fmaf(dy[b], h2T[tid * AP + b], g3); }
        else if (tid == H2) { for (int b = 0; b < BL; ++b) g3 += dy[b]; }
        __syncthreads();
        if (tid < H2) __stcg(gp + L.oW3 + tid, g3);
        else if (tid == H2) __stcg(gp + L.ob3, g3);
        for (int o = tid; o < H2 * NQ; o += CT) {
            const int n = o % H2, q = o / H2;
            const float w3 = sw[L.oW3 + n];
            float4 h = *reinterpret_cast<const float4*>(h2T + n * AP + 4 * q);
            const float4 d = *reinterpret_cast<const float4*>(dy + 4 * q);
            h.x = h.x > 0.f ? d.x * w3 : 0.f; h.y = h.y > 0.f ? d.y * w3 : 0.f;
            h.z = h.z > 0.f ? d.z * w3 : 0.f; h.w = h.w > 0.f ? d.w * w3 : 0.f;
            *reinterpret_cast<float4*>(h2T + n * AP + 4 * q) = h;
        }
        __syncthreads();
        MJB_PROF(3);
        // ---- wgrad W2 first (its 64 KB of partial stores drain to L2 while the dgrad runs) ----
        // G2[n][k] = sum_b delta2[n][b] h1[k][b]; thread tile n = ng+16i (i<8), k = 4kg+j (j<4)
        {
            const int ng = tid & 15, kg = tid >> 4;
            float g[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) g[i][j] = 0.0f;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                float4 dv[8], hv[4];
#pragma unroll
                for (int i = 0; i < 8; ++i) dv[i] = *reinterpret_cast<const float4*>(h2T + min(ng + 16 * i, H2 - 1) * AP + 4 * q);
#pragma unroll
                for (int j = 0; j < 4; ++j) hv[j] = *reinterpret_cast<const float4*>(h1T + min(kg * 4 + j, H1 - 1) * AP + 4 * q);
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float t = g[i][j];
                        t = fmaf(dv[i].x, hv[j].x, t); t = fmaf(dv[i].y, hv[j].y, t);
                        t = fmaf(dv[i].z, hv[j].z, t); t = fmaf(dv[i].w, hv[j].w, t);
                        g[i][j] = t;
                    }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int n = ng + 16 * i, k = kg * 4 + j;
                    if (n < H2 && k < H1) __stcg(gp + L.oW2T + k * P2 + n, g[i][j]);
                }
        }
        if (tid < H2) {
            float t = 0.0f;
            for (int b = 0; b < BL; ++b) t += h2T[tid * AP + b];
            __stcg(gp + L.ob2 + tid, t);
        }
        // ---- dgrad: delta1[k][q] = relu'(h1) * sum_n delta2[n][q] W2T[k][n], reduction split over CT/H1 groups ----
        {
            const int k = tid % H1, ns = tid / H1, NS = CT / H1, nr = H2 / NS;
            float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
            const float* wp = sw + L.oW2T + k * P2 + ns * nr;
            const float* dp = h2T + (ns * nr) * AP;
#pragma unroll 8
            for (int i = 0; i < nr; ++i) {
                const float w = wp[i];
                const float4 x0 = *reinterpret_cast<const float4*>(dp + i * AP);
                a0.x = fmaf(x0.x, w, a0.x); a0.y = fmaf(x0.y, w, a0.y); a0.z = fmaf(x0.z, w, a0.z); a0.w = fmaf(x0.w, w, a0.w);
                if (NQ == 2) {
                    const float4 x1 = *reinterpret_cast<const float4*>(dp + i * AP + 4);
                    a1.x = fmaf(x1.x, w, a1.x); a1.y = fmaf(x1.y, w, a1.y); a1.z = fmaf(x1.z, w, a1.z); a1.w = fmaf(x1.w, w, a1.w);
                }
            }
            *reinterpret_cast<float4*>(red + tid * 8) = a0;
            *reinterpret_cast<float4*>(red + tid * 8 + 4) = a1;
        }
        __syncthreads();
        for (int o = tid; o < H1 * NQ; o += CT) {
            const int k = o % H1, q = o / H1, NS = CT / H1;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int ns = 0; ns < NS; ++ns) {
                const float4 r = *reinterpret_cast<const float4*>(red + (ns * H1 + k) * 8 + 4 * q);
                acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w;
            }
            const float4 h = *reinterpret_cast<const float4*>(h1T + k * AP + 4 * q);
            acc.x = h.x > 0.f ? acc.x : 0.f; acc.y = h.y > 0.f ? acc.y : 0.f; acc.z = h.z > 0.f ? acc.z : 0.f; acc.w = h.w > 0.f ? acc.w : 0.f;
            *reinterpret_cast<float4*>(d1T + k * AP + 4 * q) = acc;
        }
        __syncthreads();                                  // d1T complete
        MJB_PROF(4);
        // ---- wgrad W1 / b1 ----
        for (int o = tid; o < H1 * K; o += CT) {
            const int k = o / H1, n = o - k * H1;
            float t = 0.0f;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const float4 d = *reinterpret_cast<const float4*>(d1T + n * AP + 4 * q);
                const float4 x = *reinterpret_cast<const float4*>(xT + k * AP + 4 * q);
                t = fmaf(d.x, x.x, t); t = fmaf(d.y, x.y, t); t = fmaf(d.z, x.z, t); t = fmaf(d.w, x.w, t);
            }
            __stcg(gp + L.oW1T + k * P1 + n, t);
        }
        if (tid < H1) {
            float t = 0.0f;
            for (int b = 0; b < BL; ++b) t += d1T[tid * AP + b];
            __stcg(gp + L.ob1 + tid, t);
        }
        MJB_PROF(5);
        cluster_sync_relacq();                            // #1: every CTA's partial is in L2
        MJB_PROF(6);
        // ---- owner: sum the C partials of my slice, Adam, publish (all loads issued before any use) ----
        {
            float gs[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int i = tid + CT * j;
                float g = 0.0f;
                if (i < L.per) {
#pragma unroll
                    for (int cc = 0; cc < C; ++cc) g += __ldcg(a.gpart + (size_t)cc * L.total + own0 + i);
                }
                gs[j] = g;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int i = tid + CT * j;
                if (i < L.per) {
                    const float w = sw[own0 + i];
                    const float g = fmaf(ck.reg, w, gs[j]);
                    const float mn = s_m[i] + ck.one_m_b1 * (g - s_m[i]);
                    const float vn = fmaf(ck.one_m_b2 * g, g, s_v[i] * ck.b2);
                    s_m[i] = mn; s_v[i] = vn;
                    const float denom = sqrtf(vn) / ck.bc2_sqrt + ck.eps;
                    __stcg(a.kw + own0 + i, fmaf(ck.neg_step, mn / denom, w));
                }
            }
        }
        MJB_PROF(7);
        cluster_sync_relacq();                            // #2: new weights visible
        MJB_PROF(8);
        // gather for the next step: issued after the last cluster barrier of this step (a barrier's release fence waits
        // for every outstanding load of the thread), consumed by commit_prefetch() below
        if (s + 1 < a.steps) load_vals();
        if (s + 2 < a.steps) load_idx(s + 2);
        if (tid == 0 && s + 1 < a.steps) cc_next = a.consts[s + 1];
        {   // reload all weights: one flat coalesced copy, loads batched ahead of the shared-memory stores
            constexpr int RB = 12;
            float4 r[RB];
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                const int i = (tid + CT * j) * 4;
                if (i < L.total) r[j] = __ldcg(reinterpret_cast<const float4*>(a.kw + i));
            }
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                const int i = (tid + CT * j) * 4;
                if (i < L.total) *reinterpret_cast<float4*>(sw + i) = r[j];
            }
            for (int i = (tid + CT * RB) * 4; i < L.total; i += CT * 4)
                *reinterpret_cast<float4*>(sw + i) = __ldcg(reinterpret_cast<const float4*>(a.kw + i));
        }
        if (s + 1 < a.steps) commit_prefetch();
        __syncthreads();
        MJB_PROF(9);
    }
    for (int i = tid; i < L.per; i += CT) { a.km[own0 + i] = s_m[i]; a.kv[own0 + i] = s_v[i]; }
}

template <int C>
cudaError_t launch_cluster(const ClArgs& a, size_t smem, cudaStream_t s) {
    auto kern = vf_fit_cluster_kernel<C>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (C > 8) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(C); cfg.blockDim = dim3(CT); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, a);
}

}  // namespace

// Scratch needed by the cluster path, in floats: kw, km, kv (3 x total) + gpart (C x total).
size_t vf_cluster_scratch_floats(int K, int H1, int H2, int C) {
    const KL L = make_kl(K, H1, H2, C);
    return (size_t)(3 + C) * L.total;
}

bool vf_cluster_supported(int K, int H1, int H2, int batch, int C) {
    if (batch != 64 || (C != 8 && C != 16)) return false;
    auto ok = [](int h) { return h == 16 || h == 32 || h == 64 || h == 128; };
    if (!ok(H1) || !ok(H2)) return false;
    const KL L = make_kl(K, H1, H2, C);
    if ((batch / C) * K > CT || L.per > 8 * CT) return false;
    const size_t smem = ((size_t)L.total + 2 * L.per + CT * 8 + (size_t)(K + 2 * H1 + H2) * AP + 16) * 4;
    return smem <= 200 * 1024;
}

static long long* g_vf_prof = nullptr;
void vf_cluster_set_prof(long long* dev16) { g_vf_prof = dev16; }

cudaError_t launch_vf_fit_cluster(const VfFitArgs& v, float* scratch, int C, cudaStream_t s) {
    const KL L = make_kl(v.K, v.H1, v.H2, C);
    float* kw = scratch; float* km = kw + L.total; float* kv = km + L.total; float* gpart = kv + L.total;
    static float2* consts = nullptr;
    static int consts_cap = 0;
    if (v.steps > consts_cap) {
        if (consts) cudaFree(consts);
        consts_cap = v.steps + 1024;
        cudaError_t ce = cudaMalloc(&consts, sizeof(float2) * consts_cap);
        if (ce != cudaSuccess) { consts = nullptr; consts_cap = 0; return ce; }
    }
    vf_adam_consts_kernel<<<(v.steps + 255) / 256, 256, 0, s>>>(consts, v.steps, v.step0, v.lr, v.beta1, v.beta2);
    const int n_nat = v.H1 * v.K + v.H1 + v.H2 * v.H1 + v.H2 + v.H2 + 1;
    cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(float) * (size_t)(3 + C) * L.total, s);
    if (e != cudaSuccess) return e;
    vf_relayout_kernel<<<(n_nat + 255) / 256, 256, 0, s>>>(v.w, v.m, v.v, kw, km, kv, L, n_nat, 1);
    ClArgs a;
    a.L = L; a.obs_dim = v.obs_dim; a.batch = v.batch; a.steps = v.steps;
    a.obs = v.obs; a.tstep = v.tstep; a.returns = v.returns; a.perm = v.perm;
    a.lr = v.lr; a.reg = v.reg; a.beta1 = v.beta1; a.beta2 = v.beta2; a.eps = v.eps; a.step0 = v.step0;
    a.kw = kw; a.km = km; a.kv = kv; a.gpart = gpart; a.prof = g_vf_prof; a.consts = consts; a.red = nullptr;
    const size_t smem = ((size_t)L.total + 2 * L.per + CT * 8 + (size_t)(v.K + 2 * v.H1 + v.H2) * AP + 16) * 4;
    e = (C == 8) ? launch_cluster<8>(a, smem, s) : (C == 16 ? launch_cluster<16>(a, smem, s) : cudaErrorInvalidValue);
    if (e != cudaSuccess) return e;
    vf_relayout_kernel<<<(n_nat + 255) / 256, 256, 0, s>>>(v.w, v.m, v.v, kw, km, kv, L, n_nat, 0);
    return cudaGetLastError();
}

}  // namespace mjb

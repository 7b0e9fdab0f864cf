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

template <int C>
__global__ void __launch_bounds__(MT_, 1) vf_fit_mp_kernel(const MpArgs a) {
    cg::cluster_group cluster = cg::this_cluster();
    const int c = (int)cluster.block_rank();
    const int K = a.K, H1 = a.H1, H2 = a.H2, u1 = a.u1, u2 = a.u2;
    const int W2P = H1 + 4;                       // pitch of owned W2 rows
    const int tid = threadIdx.x;
    extern __shared__ __align__(16) float sm[];
    // ---- owned parameters + Adam state (persist for the whole epoch) ----
    float* W1o = sm;                              // [u1][K]
    float* b1o = W1o + u1 * K;                    // [u1]
    float* W2o = b1o + u1;                        // [u2][W2P]
    float* b2o = W2o + u2 * W2P;                  // [u2]
    float* W3o = b2o + u2;                        // [u2]
    float* b3r = W3o + u2;                        // [1] replicated in every CTA (identical updates)
    const int np = u1 * K + u1 + u2 * W2P + 2 * u2 + 1;
    const int npp = round_up(np, 4);
    float* Mo = sm + npp;                         // Adam exp_avg, same indexing as the parameter block
    float* Vo = Mo + npp;                         // Adam exp_avg_sq
    // ---- per-step buffers ----
    float* xT = Vo + npp;                         // [K][BP]      minibatch features (all 64 rows)
    float* h1f = xT + K * BP;                     // [H1][BP]     full h1, assembled from every CTA's slice
    float* h1o = h1f + H1 * BP;                   // [u1][BP]     owned slice of h1
    float* h2o = h1o + u1 * BP;                   // [u2][BP]     owned slice of h2, then delta2
    float* d1o = h2o + u2 * BP;                   // [u1][BP]     owned slice of delta1
    float* dg = d1o + u1 * BP;                    // [C][u1][BP]  partial dgrad slices received from every CTA
    float* yp = dg + C * u1 * BP;                 // [C][NB]      partial outputs received from every CTA
    float* red = yp + C * NB;                     // [MT_][4]     k-split partials
    float* tv = red + MT_ * 4;                    // [NB] targets
    float* dy = tv + NB;                          // [NB]
    __shared__ AdamP s_c;

    // ---- load owned parameters / moments from the natural layout ----
    const int oW1 = 0, ob1 = H1 * K, oW2 = ob1 + H1, ob2 = oW2 + H2 * H1, oW3 = ob2 + H2, ob3 = oW3 + H2;
    auto nat_index = [&](int p) -> int {          // owned-block index -> natural flat index
        if (p < u1 * K) return oW1 + (c * u1 + p / K) * K + p % K;
        p -= u1 * K;
        if (p < u1) return ob1 + c * u1 + p;
        p -= u1;
        if (p < u2 * W2P) { const int n = p / W2P, k = p % W2P; return k < H1 ? oW2 + (c * u2 + n) * H1 + k : -1; }
        p -= u2 * W2P;
        if (p < u2) return ob2 + c * u2 + p;
        p -= u2;
        if (p < u2) return oW3 + c * u2 + p;
        return ob3;
    };
    for (int p = tid; p < np; p += MT_) {
        const int j = nat_index(p);
        W1o[p] = j >= 0 ? a.w[j] : 0.0f;
        Mo[p] = j >= 0 ? a.m[j] : 0.0f;
        Vo[p] = j >= 0 ? a.v[j] : 0.0f;
    }
    for (int i = tid; i < (K + H1 + 3 * u1 + u2) * BP + C * u1 * BP - u1 * BP + C * NB + MT_ * 4 + 2 * NB; i += MT_) xT[i] = 0.0f;
    __syncthreads();

    // ---- minibatch gather pipeline: indices two steps ahead, feature rows one step ahead, in registers ----
    constexpr int NPF = 4;                         // NB*K <= NPF*MT_  (K <= 32)
    float pre_x[NPF];
    float pre_t = 0.f;
    int r_nxt[NPF], rt_nxt = 0;
    auto load_idx = [&](int s) {
        const int* pidx = a.perm + (size_t)s * NB;
#pragma unroll
        for (int u = 0; u < NPF; ++u) { const int e = tid + MT_ * u; r_nxt[u] = (e < NB * K) ? pidx[e / K] : 0; }
        rt_nxt = (tid < NB) ? pidx[tid] : 0;
    };
    auto load_vals = [&]() {
#pragma unroll
        for (int u = 0; u < NPF; ++u) {
            const int e = tid + MT_ * u;
            if (e < NB * K) pre_x[u] = a.feat[(size_t)r_nxt[u] * K + e % K];
        }
        if (tid < NB) pre_t = a.ret32[rt_nxt];
    };
    auto commit = [&]() {
#pragma unroll
        for (int u = 0; u < NPF; ++u) {
            const int e = tid + MT_ * u;
            if (e < NB * K) xT[(e % K) * BP + e / K] = pre_x[u];
        }
        if (tid < NB) tv[tid] = pre_t;
    };
    load_idx(0);
    load_vals();
    commit();
    if (a.steps > 1) load_idx(1);
    float2 cc_next = make_float2(1.f, 0.f);
    if (tid == 0) cc_next = a.consts[0];
    cluster_sync_relacq();                         // everybody's buffers are zeroed before any remote store

    long long t_last = clock64();
#define MP_PROF(i) do { if (a.prof && tid == 0 && c == 0) { const long long _t = clock64(); a.prof[i] += _t - t_last; t_last = _t; } } while (0)
    for (int s = 0; s < a.steps; ++s) {
        if (tid == 0) {
            const float2 cc = cc_next;
            s_c.one_m_b1 = (float)(1.0 - (double)a.beta1); s_c.b2 = a.beta2; s_c.one_m_b2 = (float)(1.0 - (double)a.beta2);
            s_c.bc2_sqrt = cc.x; s_c.eps = a.eps; s_c.neg_step = cc.y; s_c.reg = a.reg;
        }
        // ---- P1: owned slice of layer 1 ----
        for (int o = tid; o < u1 * 16; o += MT_) {
            const int n = o % u1, q = o / u1;
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
        // ---- E1: my h1 slice -> every CTA's full h1 ----
        for (int i = tid; i < u1 * 16 * C; i += MT_) {
            const int item = i % (u1 * 16), r = i / (u1 * 16);
            const int n = item % u1, q = item / u1;
            const float4 v = *reinterpret_cast<const float4*>(h1o + n * BP + 4 * q);
            float* dst = cluster.map_shared_rank(h1f, r);
            *reinterpret_cast<float4*>(dst + (c * u1 + n) * BP + 4 * q) = v;
        }
        MP_PROF(0);
        cluster_sync_relacq();                     // #1: full h1 everywhere
        MP_PROF(1);
        // ---- P2: owned slice of layer 2 (reduction over H1 split in 4) ----
        {
            const int items = u2 * 16, KS = MT_ / items, kr = H1 / KS;       // items*KS == MT_ for u2*16 | 512
            const int item = tid % items, ks = tid / items;
            const int n = item % u2, q = item / u2;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ks < KS) {
                const float* wr = W2o + n * W2P + ks * kr;
                const float* hp = h1f + (ks * kr) * BP + 4 * q;
#pragma unroll 8
                for (int i = 0; i < kr; ++i) {
                    const float w = wr[i];
                    const float4 x = *reinterpret_cast<const float4*>(hp + i * BP);
                    acc.x = fmaf(x.x, w, acc.x); acc.y = fmaf(x.y, w, acc.y); acc.z = fmaf(x.z, w, acc.z); acc.w = fmaf(x.w, w, acc.w);
                }
            }
            *reinterpret_cast<float4*>(red + tid * 4) = acc;
            __syncthreads();
            if (tid < items) {
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int k2 = 0; k2 < KS; ++k2) {
                    const float4 r = *reinterpret_cast<const float4*>(red + (k2 * items + tid) * 4);
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
            for (int n = 0; n < u2; ++n) t = fmaf(h2o[n * BP + b], W3o[n], t);
            cluster.map_shared_rank(yp, r)[c * NB + b] = t;
        }
        MP_PROF(2);
        cluster_sync_relacq();                     // #2: all partial outputs present
        MP_PROF(3);
        if (tid < NB) {
            float y = 0.0f;
            for (int r = 0; r < C; ++r) y += yp[r * NB + tid];
            dy[tid] = 2.0f * ((y + b3r[0]) - tv[tid]) / (float)NB;
        }
        __syncthreads();
        const AdamP ck = s_c;
        // ---- P3: delta2 (in place of h2), small gradients ----
        float g_small = 0.0f;                      // thread n<u2: gW3[n]; thread u2+n: gb2[n] (after delta2); thread 2*u2: gb3
        if (tid < u2) { for (int b = 0; b < NB; ++b) g_small = fmaf(dy[b], h2o[tid * BP + b], g_small); }
        else if (tid == 2 * u2) { for (int b = 0; b < NB; ++b) g_small += dy[b]; }
        __syncthreads();
        for (int o = tid; o < u2 * 16; o += MT_) {
            const int n = o % u2, q = o / u2;
            const float w3 = W3o[n];
            float4 h = *reinterpret_cast<const float4*>(h2o + n * BP + 4 * q);
            const float4 d = *reinterpret_cast<const float4*>(dy + 4 * q);
            h.x = h.x > 0.f ? d.x * w3 : 0.f; h.y = h.y > 0.f ? d.y * w3 : 0.f;
            h.z = h.z > 0.f ? d.z * w3 : 0.f; h.w = h.w > 0.f ? d.w * w3 : 0.f;
            *reinterpret_cast<float4*>(h2o + n * BP + 4 * q) = h;
        }
        __syncthreads();
        if (tid >= u2 && tid < 2 * u2) { for (int b = 0; b < NB; ++b) g_small += h2o[(tid - u2) * BP + b]; }
        // ---- partial dgrad over my units, scattered to the owners of each h1 unit (E3) ----
        for (int o = tid; o < H1 * 4; o += MT_) {
            const int k = o % H1, qg = o / H1;                 // 4 quads per thread
            float4 acc[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int n = 0; n < u2; ++n) {
                const float w = W2o[n * W2P + k];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 d = *reinterpret_cast<const float4*>(h2o + n * BP + 4 * (4 * qg + j));
                    acc[j].x = fmaf(d.x, w, acc[j].x); acc[j].y = fmaf(d.y, w, acc[j].y);
                    acc[j].z = fmaf(d.z, w, acc[j].z); acc[j].w = fmaf(d.w, w, acc[j].w);
                }
            }
            float* dst = cluster.map_shared_rank(dg, k / u1) + (c * u1 + k % u1) * BP;
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(dst + 4 * (4 * qg + j)) = acc[j];
        }
        // ---- wgrad of my W2 rows: g[n][k] = sum_b delta2[n][b] h1[k][b]  (kept in registers until the update) ----
        float gw2[4] = {0.f, 0.f, 0.f, 0.f};
        {
            const int k = tid % H1, ng = tid / H1, NG = MT_ / H1, per = (u2 + NG - 1) / NG;   // per <= 4
            for (int q = 0; q < 16; ++q) {
                const float4 h = *reinterpret_cast<const float4*>(h1f + k * BP + 4 * q);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int n = ng * per + j;
                    if (j < per && n < u2) {
                        const float4 d = *reinterpret_cast<const float4*>(h2o + n * BP + 4 * q);
                        gw2[j] = fmaf(d.x, h.x, gw2[j]); gw2[j] = fmaf(d.y, h.y, gw2[j]);
                        gw2[j] = fmaf(d.z, h.z, gw2[j]); gw2[j] = fmaf(d.w, h.w, gw2[j]);
                    }
                }
            }
        }
        MP_PROF(4);
        cluster_sync_relacq();                     // #3: all partial dgrad slices delivered; W2 reads are done
        MP_PROF(5);
        // Gather for the next step is issued HERE, after the last cluster barrier of this step: the barrier's release
        // fence (MEMBAR.ALL.GPU in SASS) waits for every outstanding load of the thread, so loads issued earlier would
        // put their HBM/L2 latency onto the barrier.  They complete under P4/P5 and are consumed by commit() below.
        if (s + 1 < a.steps) load_vals();
        if (s + 2 < a.steps) load_idx(s + 2);
        if (tid == 0 && s + 1 < a.steps) cc_next = a.consts[s + 1];
        // ---- P4: delta1 of my units (fixed-order sum over the C sources), then W1 / b1 gradients ----
        for (int o = tid; o < u1 * 16; o += MT_) {
            const int n = o % u1, q = o / u1;
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = 0; r < C; ++r) {
                const float4 d = *reinterpret_cast<const float4*>(dg + (r * u1 + n) * BP + 4 * q);
                t.x += d.x; t.y += d.y; t.z += d.z; t.w += d.w;
            }
            const float4 h = *reinterpret_cast<const float4*>(h1o + n * BP + 4 * q);
            t.x = h.x > 0.f ? t.x : 0.f; t.y = h.y > 0.f ? t.y : 0.f; t.z = h.z > 0.f ? t.z : 0.f; t.w = h.w > 0.f ? t.w : 0.f;
            *reinterpret_cast<float4*>(d1o + n * BP + 4 * q) = t;
        }
        __syncthreads();
        // ---- P5: Adam on everything I own ----
        {   // W2 rows
            const int k = tid % H1, ng = tid / H1, NG = MT_ / H1, per = (u2 + NG - 1) / NG;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = ng * per + j;
                if (j < per && n < u2) {
                    float* p = W2o + n * W2P + k;
                    *p = adam_apply(gw2[j], *p, Mo + (p - sm), Vo + (p - sm), ck);
                }
            }
        }
        for (int o = tid; o < u1 * K; o += MT_) {   // W1 rows
            const int n = o / K, k = o - n * K;
            float g = 0.0f;
            for (int q = 0; q < 16; ++q) {
                const float4 d = *reinterpret_cast<const float4*>(d1o + n * BP + 4 * q);
                const float4 x = *reinterpret_cast<const float4*>(xT + k * BP + 4 * q);
                g = fmaf(d.x, x.x, g); g = fmaf(d.y, x.y, g); g = fmaf(d.z, x.z, g); g = fmaf(d.w, x.w, g);
            }
            W1o[o] = adam_apply(g, W1o[o], Mo + o, Vo + o, ck);
        }
        if (tid < u1) {                            // b1
            float g = 0.0f;
            for (int b = 0; b < NB; ++b) g += d1o[tid * BP + b];
            float* p = b1o + tid;
            *p = adam_apply(g, *p, Mo + (p - sm), Vo + (p - sm), ck);
        }
        __syncthreads();                           // W3 / b2 / b3 are read above by other threads this step: update last
        if (tid < u2) { float* p = W3o + tid; *p = adam_apply(g_small, *p, Mo + (p - sm), Vo + (p - sm), ck); }
        else if (tid < 2 * u2) { float* p = b2o + (tid - u2); *p = adam_apply(g_small, *p, Mo + (p - sm), Vo + (p - sm), ck); }
        else if (tid == 2 * u2) { float* p = b3r; *p = adam_apply(g_small, *p, Mo + (p - sm), Vo + (p - sm), ck); }
        if (s + 1 < a.steps) commit();
        __syncthreads();
        MP_PROF(6);
    }
    // ---- write the owned parameters / moments back to the natural layout ----
    for (int p = tid; p < np; p += MT_) {
        const int j = nat_index(p);
        if (j >= 0 && (j != ob3 || c == 0)) { a.w[j] = W1o[p]; a.m[j] = Mo[p]; a.v[j] = Vo[p]; }
    }
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
    const int u1 = H1 / C, u2 = H2 / C, W2P = H1 + 4;
    const int np = u1 * K + u1 + u2 * W2P + 2 * u2 + 1, npp = round_up(np, 4);
    const size_t fl = 3 * (size_t)npp + (size_t)(K + H1 + 3 * u1 + u2) * BP + (size_t)(C - 1) * u1 * BP + (size_t)C * NB + MT_ * 4 + 2 * NB + 64;
    return fl * 4;
}

template <int C>
cudaError_t launch_mp(const MpArgs& a, size_t smem, cudaStream_t s) {
    auto kern = vf_fit_mp_kernel<C>;
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
    if (H1 % C || H2 % C || H1 > 256 || H2 > 256) return false;
    const int u1 = H1 / C, u2 = H2 / C;
    if (NB * K > 4 * MT_) return false;                       // gather prefetch registers
    if (MT_ % (u2 * 16) || (MT_ / (u2 * 16)) < 1 || H1 % (MT_ / (u2 * 16))) return false;   // layer-2 k-split mapping
    if (MT_ % H1 || (u2 + MT_ / H1 - 1) / (MT_ / H1) > 4) return false;                     // wgrad W2 mapping
    if (H1 * 4 % MT_ && H1 * 4 > MT_) return false;
    if (2 * u2 + 1 > MT_ || u1 < 1) return false;
    return mp_smem_bytes(K, H1, H2, C) <= 200 * 1024;
}

// Build the fp32 feature matrix + fp32 targets once per fit (all epochs reuse them).
cudaError_t vf_build_features(const VfFitArgs& v, float* feat, float* ret32, cudaStream_t s) {
    vf_features_kernel<<<148 * 8, 256, 0, s>>>(v.obs, v.tstep, v.returns, v.n, v.obs_dim, v.K, feat, ret32);
    return cudaGetLastError();
}

cudaError_t launch_vf_fit_mp(const VfFitArgs& v, const float* feat, const float* ret32, int C, cudaStream_t s) {
    static float2* consts = nullptr;
    static int consts_cap = 0;
    if (v.steps > consts_cap) {
        if (consts) cudaFree(consts);
        consts_cap = v.steps + 1024;
        cudaError_t ce = cudaMalloc(&consts, sizeof(float2) * consts_cap);
        if (ce != cudaSuccess) { consts = nullptr; consts_cap = 0; return ce; }
    }
    mp_adam_consts_kernel<<<(v.steps + 255) / 256, 256, 0, s>>>(consts, v.steps, v.step0, v.lr, v.beta1, v.beta2);
    MpArgs a;
    a.K = v.K; a.H1 = v.H1; a.H2 = v.H2; a.u1 = v.H1 / C; a.u2 = v.H2 / C; a.obs_dim = v.obs_dim; a.steps = v.steps;
    a.feat = feat; a.ret32 = ret32; a.perm = v.perm;
    a.lr = v.lr; a.reg = v.reg; a.beta1 = v.beta1; a.beta2 = v.beta2; a.eps = v.eps;
    a.w = v.w; a.m = v.m; a.v = v.v; a.consts = consts; a.prof = g_mp_prof;
    const size_t smem = mp_smem_bytes(v.K, v.H1, v.H2, C);
    cudaError_t e = (C == 8) ? launch_mp<8>(a, smem, s) : (C == 16 ? launch_mp<16>(a, smem, s) : cudaErrorInvalidValue);
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace mjb

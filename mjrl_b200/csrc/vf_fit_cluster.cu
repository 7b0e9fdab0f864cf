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

constexpr int CT = 256;        // threads per CTA
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

struct ClArgs {
    KL L;
    int obs_dim, batch, steps;
    const float* obs; const int* tstep; const double* returns; const int* perm;
    float lr, reg, beta1, beta2, eps; long long step0;
    float* kw; float* km; float* kv;      // kernel-layout weights / moments (global)
    float* gpart;                         // [C][total] gradient partials
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
    const int BL = a.batch / C, NQ = BL / 4;          // local rows, quads of rows
    extern __shared__ __align__(16) float sm[];
    float* sw = sm;                                   // [total] weights, kernel layout
    float* s_m = sw + L.total;                        // [per] Adam exp_avg of the owned slice
    float* s_v = s_m + L.per;                         // [per] Adam exp_avg_sq
    float* xT = s_v + L.per;                          // [K][AP]
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
    // element slot e = tid + CT*u (u < 2) of the local [BL][K] feature block.
    float pre_x[2] = {0.f, 0.f};
    double pre_t64 = 0.0;
    int r_nxt[2] = {0, 0};
    int rt_nxt = 0;
    auto load_idx = [&](int s) {
        const int* pidx = a.perm + (size_t)s * a.batch + c * BL;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int e = tid + CT * u;
            r_nxt[u] = (e < BL * K) ? pidx[e / K] : 0;
        }
        rt_nxt = (tid < BL) ? pidx[tid] : 0;
    };
    auto load_vals = [&]() {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int e = tid + CT * u;
            if (e < BL * K) {                            // raw loads only: nothing here consumes the value
                const int k = e % K;
                const long long r = r_nxt[u];
                pre_x[u] = (k < a.obs_dim) ? a.obs[r * a.obs_dim + k] : __int_as_float(a.tstep[r]);
            }
        }
        if (tid < BL) pre_t64 = a.returns[rt_nxt];
    };
    auto commit_prefetch = [&]() {                       // feature map of mlp_baseline.py:36-58 applied here
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int e = tid + CT * u;
            if (e < BL * K) {
                const int k = e % K;
                float val;
                if (k < a.obs_dim) {
                    double x = (double)pre_x[u];
                    x = fmin(fmax(x, -10.0), 10.0) / 10.0;
                    val = (float)x;
                } else {
                    const double tau = (double)__float_as_int(pre_x[u]) / 1000.0;
                    double p = tau;
                    for (int q = a.obs_dim; q < k; ++q) p *= tau;
                    val = (float)p;
                }
                xT[k * AP + e / K] = val;
            }
        }
        if (tid < BL) tv[tid] = (float)pre_t64;
    };
    load_idx(0);
    load_vals();
    commit_prefetch();
    if (a.steps > 1) load_idx(1);
    __syncthreads();

    for (int s = 0; s < a.steps; ++s) {
        if (tid == 0) {
            const double t = (double)(a.step0 + s + 1);
            const double bc1 = 1.0 - pow((double)a.beta1, t), bc2 = 1.0 - pow((double)a.beta2, t);
            s_c.one_m_b1 = (float)(1.0 - (double)a.beta1);
            s_c.b2 = a.beta2;
            s_c.one_m_b2 = (float)(1.0 - (double)a.beta2);
            s_c.bc2_sqrt = (float)sqrt(bc2);
            s_c.eps = a.eps;
            s_c.neg_step = (float)(-((double)a.lr / bc1));
            s_c.reg = a.reg;
        }
        if (s + 1 < a.steps) load_vals();              // rows of step s+1: HBM-latency loads in flight during this step
        if (s + 2 < a.steps) load_idx(s + 2);          // indices of step s+2
        // ---- forward layer 1: h1[n][q] = relu(sum_k x[k][q] W1T[k][n] + b1[n]) ----
        for (int o = tid; o < H1 * NQ; o += CT) {
            const int n = o % H1, q = o / H1;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
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
        // ---- forward layer 2 ----
        for (int o = tid; o < H2 * NQ; o += CT) {
            const int n = o % H2, q = o / H2;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
            for (int k = 0; k < H1; ++k) {
                const float w = sw[L.oW2T + k * P2 + n];
                const float4 x = *reinterpret_cast<const float4*>(h1T + k * AP + 4 * q);
                acc.x = fmaf(x.x, w, acc.x); acc.y = fmaf(x.y, w, acc.y); acc.z = fmaf(x.z, w, acc.z); acc.w = fmaf(x.w, w, acc.w);
            }
            const float bb = sw[L.ob2 + n];
            acc.x = fmaxf(acc.x + bb, 0.f); acc.y = fmaxf(acc.y + bb, 0.f); acc.z = fmaxf(acc.z + bb, 0.f); acc.w = fmaxf(acc.w + bb, 0.f);
            *reinterpret_cast<float4*>(h2T + n * AP + 4 * q) = acc;
        }
        __syncthreads();
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
        // ---- dgrad: delta1[k][q] = relu'(h1) * sum_n delta2[n][q] W2T[k][n] -> d1T ----
        for (int o = tid; o < H1 * NQ; o += CT) {
            const int k = o % H1, q = o / H1;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            const float* wrow = sw + L.oW2T + k * P2;
#pragma unroll 8
            for (int n = 0; n < H2; ++n) {
                const float w = wrow[n];
                const float4 d = *reinterpret_cast<const float4*>(h2T + n * AP + 4 * q);
                acc.x = fmaf(d.x, w, acc.x); acc.y = fmaf(d.y, w, acc.y); acc.z = fmaf(d.z, w, acc.z); acc.w = fmaf(d.w, w, acc.w);
            }
            const float4 h = *reinterpret_cast<const float4*>(h1T + k * AP + 4 * q);
            acc.x = h.x > 0.f ? acc.x : 0.f; acc.y = h.y > 0.f ? acc.y : 0.f; acc.z = h.z > 0.f ? acc.z : 0.f; acc.w = h.w > 0.f ? acc.w : 0.f;
            *reinterpret_cast<float4*>(d1T + k * AP + 4 * q) = acc;
        }
        // ---- wgrad W2: G2[n][k] = sum_b delta2[n][b] h1[k][b]; thread tile n = ng+16i, k = 8kg+j ----
        {
            const int ng = tid & 15, kg = tid >> 4;
            float g[8][8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) g[i][j] = 0.0f;
            for (int q = 0; q < NQ; ++q) {
                float4 dv[8], hv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int n = min(ng + 16 * i, H2 - 1);
                    dv[i] = *reinterpret_cast<const float4*>(h2T + n * AP + 4 * q);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int k = min(kg * 8 + j, H1 - 1);
                    hv[j] = *reinterpret_cast<const float4*>(h1T + k * AP + 4 * q);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float t = g[i][j];
                        t = fmaf(dv[i].x, hv[j].x, t); t = fmaf(dv[i].y, hv[j].y, t);
                        t = fmaf(dv[i].z, hv[j].z, t); t = fmaf(dv[i].w, hv[j].w, t);
                        g[i][j] = t;
                    }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int n = ng + 16 * i, k = kg * 8 + j;
                    if (n < H2 && k < H1) __stcg(gp + L.oW2T + k * P2 + n, g[i][j]);
                }
        }
        if (tid < H2) {
            float t = 0.0f;
            for (int b = 0; b < BL; ++b) t += h2T[tid * AP + b];
            __stcg(gp + L.ob2 + tid, t);
        }
        __syncthreads();                                  // d1T complete
        // ---- wgrad W1 / b1 ----
        for (int o = tid; o < H1 * K; o += CT) {
            const int k = o / H1, n = o - k * H1;
            float t = 0.0f;
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
        cluster.sync();                                   // #1: every CTA's partial is in L2
        // ---- owner: sum the C partials of my slice, Adam, publish ----
        for (int i = tid; i < L.per; i += CT) {
            const int p = own0 + i;
            float g = 0.0f;
#pragma unroll
            for (int cc = 0; cc < C; ++cc) g += __ldcg(a.gpart + (size_t)cc * L.total + p);
            const float w = sw[p];
            g = fmaf(ck.reg, w, g);
            const float mn = s_m[i] + ck.one_m_b1 * (g - s_m[i]);
            const float vn = fmaf(ck.one_m_b2 * g, g, s_v[i] * ck.b2);
            s_m[i] = mn; s_v[i] = vn;
            const float denom = sqrtf(vn) / ck.bc2_sqrt + ck.eps;
            __stcg(a.kw + p, fmaf(ck.neg_step, mn / denom, w));
        }
        cluster.sync();                                   // #2: new weights visible
        for (int i = tid * 4; i < L.total; i += CT * 4)
            *reinterpret_cast<float4*>(sw + i) = __ldcg(reinterpret_cast<const float4*>(a.kw + i));
        if (s + 1 < a.steps) commit_prefetch();
        __syncthreads();
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
    if (H1 > 128 || H2 > 128 || H1 < 16 || H2 < 16 || batch % (4 * C) != 0 || batch / C > 8) return false;
    const KL L = make_kl(K, H1, H2, C);
    if ((batch / C) * K > 2 * CT) return false;
    const size_t smem = ((size_t)L.total + 2 * L.per + (size_t)(K + 2 * H1 + H2) * AP + 16) * 4;
    return smem <= 200 * 1024;
}

cudaError_t launch_vf_fit_cluster(const VfFitArgs& v, float* scratch, int C, cudaStream_t s) {
    const KL L = make_kl(v.K, v.H1, v.H2, C);
    float* kw = scratch; float* km = kw + L.total; float* kv = km + L.total; float* gpart = kv + L.total;
    const int n_nat = v.H1 * v.K + v.H1 + v.H2 * v.H1 + v.H2 + v.H2 + 1;
    cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(float) * (size_t)(3 + C) * L.total, s);
    if (e != cudaSuccess) return e;
    vf_relayout_kernel<<<(n_nat + 255) / 256, 256, 0, s>>>(v.w, v.m, v.v, kw, km, kv, L, n_nat, 1);
    ClArgs a;
    a.L = L; a.obs_dim = v.obs_dim; a.batch = v.batch; a.steps = v.steps;
    a.obs = v.obs; a.tstep = v.tstep; a.returns = v.returns; a.perm = v.perm;
    a.lr = v.lr; a.reg = v.reg; a.beta1 = v.beta1; a.beta2 = v.beta2; a.eps = v.eps; a.step0 = v.step0;
    a.kw = kw; a.km = km; a.kv = kv; a.gpart = gpart;
    const size_t smem = ((size_t)L.total + 2 * L.per + (size_t)(v.K + 2 * v.H1 + v.H2) * AP + 16) * 4;
    e = (C == 8) ? launch_cluster<8>(a, smem, s) : (C == 16 ? launch_cluster<16>(a, smem, s) : cudaErrorInvalidValue);
    if (e != cudaSuccess) return e;
    vf_relayout_kernel<<<(n_nat + 255) / 256, 256, 0, s>>>(v.w, v.m, v.v, kw, km, kv, L, n_nat, 0);
    return cudaGetLastError();
}

}  // namespace mjb

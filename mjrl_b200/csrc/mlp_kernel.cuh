// Fused 2-hidden-layer MLP tile kernel (fp32 FMA path): forward, tangent-forward and backward of the
// Gaussian-MLP policy (and forward of the ReLU value net) for one tile of MT consecutive timesteps
// per iteration of a persistent CTA.  One template, four modes:
//
//   MODE_EVAL : fused forward + log-likelihood + likelihood-ratio + surrogate/KL partial sums
//               (policies/gaussian_mlp.py:99-145, algos/batch_reinforce.py:40-52)
//   MODE_VPG  : MODE_EVAL + backward of mean(LR*adv) -> flat gradient partials
//               (algos/batch_reinforce.py:54-58)
//   MODE_FVP  : forward + tangent-forward (J v) + backward (J^T W J v): the Fisher-vector product the
//               reference obtains by double backward through mean_kl (algos/npg_cg.py:62-81)
//   MODE_VF   : MLPBaseline features + forward (baselines/mlp_baseline.py:36-58,97-105)
//
// Data layout in shared memory: activations are *feature-major* ([feature][sample], row pitch
// LDM = MT+4 floats, so a float4 of 4 consecutive samples is one 16-byte access and rows of
// different features fall in different bank groups).  All GEMMs are then outer-product updates over
// the reduction index with both operand fragments read as float4:
//   forward/dgrad : acc[m][n] += A_T[r][m] * B[r][n]     (B = weight slice streamed from L2 by cp.async)
//   wgrad         : acc[a][b] += sum_m A_T[a][m] * B_T[b][m]
// Weights are tiny (<= 333 KB) and shared by every CTA, so they are streamed from L2 through a
// double-buffered 2 x 4 KB slice ring instead of being pinned in shared memory; that leaves room for
// three H x MT activation buffers (h1, h2, scratch) so no activation ever goes to HBM.
// Algorithmic HBM traffic is therefore obs (+act, adv) read once per launch.
#pragma once
#include "common.cuh"

namespace mjb {

enum { MODE_EVAL = 0, MODE_VPG = 1, MODE_FVP = 2, MODE_VF = 3 };
enum { OLD_READ = 1, OLD_WRITE = 2 };

struct MlpArgs {
    PrepLayout L;
    const float* P;            // prepped parameters the forward runs with
    const float* T;            // prepped tangent (FVP)
    const float* in_shift;     // [K0]  policy modes (fc_network.py:46)
    const float* in_scale;     // [K0]
    const float* out_shift;    // [A]
    const float* out_scale;    // [A]
    const float* obs;          // [rows][obs_dim] fp32 row-major
    int obs_dim;
    const float* act;          // [rows][A]
    const int* idx;            // optional gather list (hvp_sample_frac < 1)
    const int* tstep;          // VF: timestep index inside its path
    long long n;               // rows to process
    const float* weight;       // advantages (whitened, fp32) / DAPG weights
    float* ll_old;             // [rows]   log-likelihood under the old policy (cache)
    float* mu_old;             // [rows][A] mean under the old policy (cache)
    const float* old_log_std;  // [A]
    int old_flags;             // OLD_READ / OLD_WRITE
    double* eval_partial;      // [grid][2]  sum LR*w, sum KL
    float* vf_out;             // [rows]
    float* gpartial;           // [grid][gstride] per-CTA gradient partials, theta layout
    long long gstride;
};

template <int H, int MT>
struct MlpShape {
    static constexpr int LDM = MT + 4;
    static constexpr int SR = 1024 / H;      // weight rows per streamed slice (slice = 1024 floats)
    static constexpr int TMG = MT / 8;       // thread groups along the sample axis
    static constexpr int TN = H * MT / (kThreads * 8);
    static_assert(TMG * (H / TN) == kThreads, "thread tiling must cover the H x MT tile");
    static_assert(TN == 4 || TN == 8, "thread tile width");
};

__host__ __device__ inline size_t mlp_smem_bytes(int H, int MT, int mode, int YR) {
    int LDM = MT + 4;
    int nbuf = (mode == MODE_VPG || mode == MODE_FVP) ? 3 : 2;
    size_t fl = (size_t)nbuf * H * LDM + 2048 + (size_t)YR * LDM + 64;
    return fl * 4 + 32 * 8;
}

template <int ACT>
__device__ __forceinline__ float activate(float z) {
    return ACT == 0 ? tanhf(z) : fmaxf(z, 0.0f);
}

// acc[i][j] += sum_r A_T[r][m_i] * B[r][n_j]; B streamed from global in slices of SR rows x H cols.
template <int H, int MT, int TN>
__device__ __forceinline__ void gemm_stream(float (&acc)[8][TN], const float* __restrict__ AT,
                                            const float* __restrict__ Bg, int R, float* wbuf) {
    using S = MlpShape<H, MT>;
    const int tid = threadIdx.x, mg = tid % S::TMG, ng = tid / S::TMG;
    const int nsl = R / S::SR;
    __syncthreads();                       // A_T complete, slice ring free
    cp_async16(wbuf + tid * 4, Bg + tid * 4);
    cp_async_commit();
    for (int s = 0; s < nsl; ++s) {
        cp_async_wait<0>();
        __syncthreads();
        if (s + 1 < nsl) {
            cp_async16(wbuf + ((s + 1) & 1) * 1024 + tid * 4, Bg + (size_t)(s + 1) * 1024 + tid * 4);
            cp_async_commit();
        }
        const float* wb = wbuf + (s & 1) * 1024 + TN * ng;
        const float* ap = AT + (size_t)(s * S::SR) * S::LDM + 4 * mg;
#pragma unroll
        for (int r = 0; r < S::SR; ++r) {
            const float4 a0 = *reinterpret_cast<const float4*>(ap + r * S::LDM);
            const float4 a1 = *reinterpret_cast<const float4*>(ap + r * S::LDM + MT / 2);
            float b[TN];
#pragma unroll
            for (int j = 0; j < TN; j += 4) {
                const float4 t = *reinterpret_cast<const float4*>(wb + r * H + j);
                b[j] = t.x; b[j + 1] = t.y; b[j + 2] = t.z; b[j + 3] = t.w;
            }
            const float am[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(am[i], b[j], acc[i][j]);
        }
    }
}

// acc[i][j] += sum_m A[(a0+i*as)][m] * B[(b0+j*bs)][m]   (both feature-major, reduction over samples)
template <int MT, int TA, int TB>
__device__ __forceinline__ void wgrad_acc(float (&acc)[TA][TB], const float* __restrict__ A, int a0, int as,
                                          const float* __restrict__ B, int b0, int bs) {
    constexpr int LDM = MT + 4;
#pragma unroll 2
    for (int m = 0; m < MT; m += 4) {
        float4 a[TA], b[TB];
#pragma unroll
        for (int i = 0; i < TA; ++i) a[i] = *reinterpret_cast<const float4*>(A + (size_t)(a0 + i * as) * LDM + m);
#pragma unroll
        for (int j = 0; j < TB; ++j) b[j] = *reinterpret_cast<const float4*>(B + (size_t)(b0 + j * bs) * LDM + m);
#pragma unroll
        for (int i = 0; i < TA; ++i)
#pragma unroll
            for (int j = 0; j < TB; ++j) {
                float t = acc[i][j];
                t = fmaf(a[i].x, b[j].x, t); t = fmaf(a[i].y, b[j].y, t);
                t = fmaf(a[i].z, b[j].z, t); t = fmaf(a[i].w, b[j].w, t);
                acc[i][j] = t;
            }
    }
}

template <int TN>
__device__ __forceinline__ void zero_acc(float (&acc)[8][TN]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;
}

// Stage chunk c (<= 32 input features) of MT rows, transposed to feature-major, applying the input
// transform (policy) or building the baseline features.  Returns the row count padded to SR.
template <int H, int MT, int MODE>
__device__ __forceinline__ int load_chunk(float* xs, const MlpArgs& a, long long base, int c) {
    using S = MlpShape<H, MT>;
    const int K0 = a.L.K0;
    const int cw = min(kChunk, K0 - kChunk * c);
    const int rows_p = round_up(cw, S::SR);
    for (int f = threadIdx.x; f < MT * cw; f += kThreads) {
        const int m = f / cw, kk = f - m * cw, k = kChunk * c + kk;
        const long long row = base + m;
        float v = 0.0f;
        if (row < a.n) {
            const long long r = a.idx ? (long long)a.idx[row] : row;
            if (MODE == MODE_VF) {
                if (k < a.obs_dim) {
                    double x = (double)a.obs[r * a.obs_dim + k];
                    x = fmin(fmax(x, -10.0), 10.0) / 10.0;
                    v = (float)x;
                } else {
                    const double tau = (double)a.tstep[r] / 1000.0;
                    double p = tau;
                    for (int q = a.obs_dim; q < k; ++q) p *= tau;
                    v = (float)p;
                }
            } else {
                v = (a.obs[r * a.obs_dim + k] - a.in_shift[k]) / (a.in_scale[k] + 1e-8f);
            }
        }
        xs[kk * S::LDM + m] = v;
    }
    for (int f = threadIdx.x; f < (rows_p - cw) * MT; f += kThreads)
        xs[(cw + f / MT) * S::LDM + (f % MT)] = 0.0f;
    return rows_p;
}

template <int H, int MT, int MODE, int ACT>
__global__ void __launch_bounds__(kThreads, (H <= 64 ? 2 : 1)) mlp_kernel(const MlpArgs a) {
    using S = MlpShape<H, MT>;
    constexpr int LDM = S::LDM, TN = S::TN, TMG = S::TMG;
    constexpr bool BWD = (MODE == MODE_VPG || MODE == MODE_FVP);
    extern __shared__ __align__(16) float smem[];
    const PrepLayout& L = a.L;
    float* aT = smem;                                   // h1 (later delta1)
    float* bT = aT + H * LDM;                           // h2; aliases the input chunk staging
    float* cT = bT + H * LDM;                           // tangent / delta scratch (BWD modes only)
    float* wbuf = BWD ? cT + H * LDM : cT;
    float* ydT = wbuf + 2048;                           // [YR][LDM] output-layer rows (feature-major)
    float* s_gs = ydT + L.YR * LDM;                     // [32] log_std gradient accumulators
    double* s_red = reinterpret_cast<double*>(s_gs + 64);
    float* xs = bT;

    const int tid = threadIdx.x, mg = tid % TMG, ng = tid / TMG;
    const float* P = a.P;
    const int A = L.A, AP = L.AP;

    for (int i = tid; i < L.YR * LDM; i += kThreads) ydT[i] = 0.0f;
    if (tid < 32) s_gs[tid] = 0.0f;
    // log_std gradient: per-thread running sums over this CTA's tiles (thread = row of the tile), reduced once at the end
    // in a fixed order (warp xor-tree, then warps 0..7) -- no floating-point atomics, so the flat gradient is
    // bit-reproducible run to run.
    float gs_acc[32];
    if (MODE == MODE_VPG) {
#pragma unroll
        for (int j = 0; j < 32; ++j) gs_acc[j] = 0.0f;
    }
    double sum0 = 0.0, sum1 = 0.0;
    float sum_ls = 0.0f;
    if (MODE == MODE_EVAL || MODE == MODE_VPG)
        for (int j = 0; j < A; ++j) sum_ls += P[L.oLS + j];
    float* gp = BWD ? a.gpartial + (size_t)blockIdx.x * a.gstride : nullptr;

    const long long n_tiles = (a.n + MT - 1) / MT;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long base = tile * MT;
        float acc[8][TN];
        const int nchunk = L.K0P / kChunk;

        // ---------------- layer 1: z1 = x~ W1^T + b1, h1 = act(z1) -> aT ----------------
        zero_acc<TN>(acc);
        for (int c = 0; c < nchunk; ++c) {
            __syncthreads();
            const int rp = load_chunk<H, MT, MODE>(xs, a, base, c);
            gemm_stream<H, MT, TN>(acc, xs, P + L.oW1T + (size_t)c * kChunk * H, rp, wbuf);
        }
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = TN * ng + j;
            const float bias = P[L.ob1 + n];
            float4 v0, v1;
            v0.x = activate<ACT>(acc[0][j] + bias); v0.y = activate<ACT>(acc[1][j] + bias);
            v0.z = activate<ACT>(acc[2][j] + bias); v0.w = activate<ACT>(acc[3][j] + bias);
            v1.x = activate<ACT>(acc[4][j] + bias); v1.y = activate<ACT>(acc[5][j] + bias);
            v1.z = activate<ACT>(acc[6][j] + bias); v1.w = activate<ACT>(acc[7][j] + bias);
            *reinterpret_cast<float4*>(aT + n * LDM + 4 * mg) = v0;
            *reinterpret_cast<float4*>(aT + n * LDM + MT / 2 + 4 * mg) = v1;
        }
        // ---------------- layer 1 tangent: hd1 = (1-h1^2)(x~ V1^T + c1) -> cT ----------------
        if (MODE == MODE_FVP) {
            zero_acc<TN>(acc);
            for (int c = 0; c < nchunk; ++c) {
                int rp = round_up(min(kChunk, L.K0 - kChunk * c), S::SR);
                if (nchunk > 1) {
                    __syncthreads();
                    rp = load_chunk<H, MT, MODE>(xs, a, base, c);
                }
                gemm_stream<H, MT, TN>(acc, xs, a.T + L.oW1T + (size_t)c * kChunk * H, rp, wbuf);
            }
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const int n = TN * ng + j;
                const float bias = a.T[L.ob1 + n];
                const float4 h0 = *reinterpret_cast<const float4*>(aT + n * LDM + 4 * mg);
                const float4 h1 = *reinterpret_cast<const float4*>(aT + n * LDM + MT / 2 + 4 * mg);
                float4 v0, v1;
                v0.x = (1.0f - h0.x * h0.x) * (acc[0][j] + bias); v0.y = (1.0f - h0.y * h0.y) * (acc[1][j] + bias);
                v0.z = (1.0f - h0.z * h0.z) * (acc[2][j] + bias); v0.w = (1.0f - h0.w * h0.w) * (acc[3][j] + bias);
                v1.x = (1.0f - h1.x * h1.x) * (acc[4][j] + bias); v1.y = (1.0f - h1.y * h1.y) * (acc[5][j] + bias);
                v1.z = (1.0f - h1.z * h1.z) * (acc[6][j] + bias); v1.w = (1.0f - h1.w * h1.w) * (acc[7][j] + bias);
                *reinterpret_cast<float4*>(cT + n * LDM + 4 * mg) = v0;
                *reinterpret_cast<float4*>(cT + n * LDM + MT / 2 + 4 * mg) = v1;
            }
        }
        // ---------------- layer 2: h2 = act(h1 W2^T + b2) -> bT ----------------
        zero_acc<TN>(acc);
        gemm_stream<H, MT, TN>(acc, aT, P + L.oW2T, H, wbuf);
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = TN * ng + j;
            const float bias = P[L.ob2 + n];
            float4 v0, v1;
            v0.x = activate<ACT>(acc[0][j] + bias); v0.y = activate<ACT>(acc[1][j] + bias);
            v0.z = activate<ACT>(acc[2][j] + bias); v0.w = activate<ACT>(acc[3][j] + bias);
            v1.x = activate<ACT>(acc[4][j] + bias); v1.y = activate<ACT>(acc[5][j] + bias);
            v1.z = activate<ACT>(acc[6][j] + bias); v1.w = activate<ACT>(acc[7][j] + bias);
            *reinterpret_cast<float4*>(bT + n * LDM + 4 * mg) = v0;
            *reinterpret_cast<float4*>(bT + n * LDM + MT / 2 + 4 * mg) = v1;
        }
        // ---------------- layer 2 tangent: hd2 = (1-h2^2)(hd1 W2^T + h1 V2^T + c2) -> cT ----------------
        if (MODE == MODE_FVP) {
            zero_acc<TN>(acc);
            gemm_stream<H, MT, TN>(acc, cT, P + L.oW2T, H, wbuf);
            gemm_stream<H, MT, TN>(acc, aT, a.T + L.oW2T, H, wbuf);
            __syncthreads();               // every read of hd1 (cT) is done
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const int n = TN * ng + j;
                const float bias = a.T[L.ob2 + n];
                const float4 h0 = *reinterpret_cast<const float4*>(bT + n * LDM + 4 * mg);
                const float4 h1 = *reinterpret_cast<const float4*>(bT + n * LDM + MT / 2 + 4 * mg);
                float4 v0, v1;
                v0.x = (1.0f - h0.x * h0.x) * (acc[0][j] + bias); v0.y = (1.0f - h0.y * h0.y) * (acc[1][j] + bias);
                v0.z = (1.0f - h0.z * h0.z) * (acc[2][j] + bias); v0.w = (1.0f - h0.w * h0.w) * (acc[3][j] + bias);
                v1.x = (1.0f - h1.x * h1.x) * (acc[4][j] + bias); v1.y = (1.0f - h1.y * h1.y) * (acc[5][j] + bias);
                v1.z = (1.0f - h1.z * h1.z) * (acc[6][j] + bias); v1.w = (1.0f - h1.w * h1.w) * (acc[7][j] + bias);
                *reinterpret_cast<float4*>(cT + n * LDM + 4 * mg) = v0;
                *reinterpret_cast<float4*>(cT + n * LDM + MT / 2 + 4 * mg) = v1;
            }
        }
        __syncthreads();                   // h2 (and hd2) visible to everyone

        // ---------------- output layer -> ydT[a][m] ----------------
        {
            constexpr int MQ = MT / 4, NAG = kThreads / MQ;
            const int mq = tid % MQ, ag = tid / MQ;
            for (int o = ag; o < AP; o += NAG) {
                float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
                const float* w3 = P + L.oW3T + o;
#pragma unroll 4
                for (int k = 0; k < H; ++k) {
                    const float4 h = *reinterpret_cast<const float4*>((MODE == MODE_FVP ? cT : bT) + k * LDM + 4 * mq);
                    const float w = __ldg(w3 + k * AP);
                    y.x = fmaf(h.x, w, y.x); y.y = fmaf(h.y, w, y.y); y.z = fmaf(h.z, w, y.z); y.w = fmaf(h.w, w, y.w);
                }
                if (MODE == MODE_FVP) {
                    const float* v3 = a.T + L.oW3T + o;
#pragma unroll 4
                    for (int k = 0; k < H; ++k) {
                        const float4 h = *reinterpret_cast<const float4*>(bT + k * LDM + 4 * mq);
                        const float w = __ldg(v3 + k * AP);
                        y.x = fmaf(h.x, w, y.x); y.y = fmaf(h.y, w, y.y); y.z = fmaf(h.z, w, y.z); y.w = fmaf(h.w, w, y.w);
                    }
                    // delta_y = out_scale^2 * 2/(2 sigma^2 + 1e-8) * ydot  (1/N applied by the finalizer)
                    float f = 0.0f;
                    if (o < A) {
                        const float sd = expf(P[L.oLS + o]);
                        const float os = a.out_scale[o];
                        f = os * os * (2.0f / (2.0f * sd * sd + 1e-8f));
                    }
                    const float c3 = a.T[L.ob3 + o];
                    const long long r0 = base + 4 * mq;
                    y.x = (r0 + 0 < a.n) ? f * (y.x + c3) : 0.0f;
                    y.y = (r0 + 1 < a.n) ? f * (y.y + c3) : 0.0f;
                    y.z = (r0 + 2 < a.n) ? f * (y.z + c3) : 0.0f;
                    y.w = (r0 + 3 < a.n) ? f * (y.w + c3) : 0.0f;
                } else {
                    const float b3 = P[L.ob3 + o];
                    y.x += b3; y.y += b3; y.z += b3; y.w += b3;
                }
                *reinterpret_cast<float4*>(ydT + o * LDM + 4 * mq) = y;
            }
        }
        __syncthreads();

        // ---------------- per-sample likelihood / KL / cotangent ----------------
        if (MODE == MODE_VF) {
            if (tid < MT && base + tid < a.n) a.vf_out[base + tid] = ydT[tid];
        }
        if (MODE == MODE_EVAL || MODE == MODE_VPG) {
            if (tid < MT) {
                const long long row = base + tid;
                const bool valid = row < a.n;
                if (valid) {
                    const float w = a.weight ? a.weight[row] : 0.0f;
                    float z2 = 0.0f, kl = 0.0f;
                    float zz[32];
#pragma unroll 1
                    for (int j = 0; j < A; ++j) {
                        const float mu = ydT[j * LDM + tid] * a.out_scale[j] + a.out_shift[j];
                        const float s = P[L.oLS + j];
                        const float sd = expf(s);
                        const float z = (a.act[row * A + j] - mu) / sd;
                        z2 += z * z;
                        if (MODE == MODE_VPG) { zz[j] = z; ydT[j * LDM + tid] = z / sd * a.out_scale[j]; }
                        if (a.old_flags & OLD_WRITE) a.mu_old[row * A + j] = mu;
                        if (a.old_flags & OLD_READ) {
                            const float so = a.old_log_std[j];
                            const float sdo = expf(so);
                            const float dm = a.mu_old[row * A + j] - mu;
                            const float nr = dm * dm + sdo * sdo - sd * sd;
                            const float dr = 2.0f * sd * sd + 1e-8f;
                            kl += nr / dr + s - so;
                        }
                    }
                    const float ll = -0.5f * z2 - sum_ls - 0.5f * (float)A * 1.8378770664093453f;
                    if (a.old_flags & OLD_WRITE) a.ll_old[row] = ll;
                    const float lr = (a.old_flags & OLD_READ) ? expf(ll - a.ll_old[row]) : 1.0f;
                    sum0 += (double)(lr * w);
                    sum1 += (double)kl;
                    if (MODE == MODE_VPG) {
                        const float coef = lr * w;
#pragma unroll 1
                        for (int j = 0; j < A; ++j) {
                            ydT[j * LDM + tid] *= coef;
                            gs_acc[j] += coef * (zz[j] * zz[j] - 1.0f);
                        }
                    }
                } else if (MODE == MODE_VPG) {
                    for (int j = 0; j < A; ++j) ydT[j * LDM + tid] = 0.0f;
                }
            }
        }
        if (!BWD) continue;

        // =============================== backward ===============================
        __syncthreads();                   // delta_y complete
        // G3[a][k] += sum_m dy[a][m] h2[k][m];  gb3[a] += sum_m dy[a][m]
        for (int o = tid; o < A * H; o += kThreads) {
            const int aa = o / H, k = o - aa * H;
            float t = 0.0f;
            for (int m = 0; m < MT; m += 4) {
                const float4 d = *reinterpret_cast<const float4*>(ydT + aa * LDM + m);
                const float4 h = *reinterpret_cast<const float4*>(bT + k * LDM + m);
                t = fmaf(d.x, h.x, t); t = fmaf(d.y, h.y, t); t = fmaf(d.z, h.z, t); t = fmaf(d.w, h.w, t);
            }
            if (k < L.h2) gp[L.tW3 + aa * L.h2 + k] += t;
        }
        if (tid < A) {
            float t = 0.0f;
            for (int m = 0; m < MT; ++m) t += ydT[tid * LDM + m];
            gp[L.tb3 + tid] += t;
        }
        // delta2 = (dy W3) * act'(h2) -> cT
        zero_acc<TN>(acc);
        gemm_stream<H, MT, TN>(acc, ydT, P + L.oW3N, L.YR, wbuf);
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = TN * ng + j;
            const float4 h0 = *reinterpret_cast<const float4*>(bT + n * LDM + 4 * mg);
            const float4 h1 = *reinterpret_cast<const float4*>(bT + n * LDM + MT / 2 + 4 * mg);
            float4 v0, v1;
            v0.x = (1.0f - h0.x * h0.x) * acc[0][j]; v0.y = (1.0f - h0.y * h0.y) * acc[1][j];
            v0.z = (1.0f - h0.z * h0.z) * acc[2][j]; v0.w = (1.0f - h0.w * h0.w) * acc[3][j];
            v1.x = (1.0f - h1.x * h1.x) * acc[4][j]; v1.y = (1.0f - h1.y * h1.y) * acc[5][j];
            v1.z = (1.0f - h1.z * h1.z) * acc[6][j]; v1.w = (1.0f - h1.w * h1.w) * acc[7][j];
            *reinterpret_cast<float4*>(cT + n * LDM + 4 * mg) = v0;
            *reinterpret_cast<float4*>(cT + n * LDM + MT / 2 + 4 * mg) = v1;
        }
        __syncthreads();                   // delta2 complete
        // G2[n][k] += sum_m delta2[n][m] h1[k][m]
        {
            constexpr int KG = (H == 256) ? 32 : 16, TB = H / KG, NGT = kThreads / KG;
            constexpr int TA = (H / NGT < 8) ? H / NGT : 8, PASSES = H / (NGT * TA);
            const int kg = tid % KG, ngw = tid / KG;
#pragma unroll 1
            for (int p = 0; p < PASSES; ++p) {
                float g[TA][TB];
#pragma unroll
                for (int i = 0; i < TA; ++i)
#pragma unroll
                    for (int j = 0; j < TB; ++j) g[i][j] = 0.0f;
                const int n0 = p * NGT * TA + ngw * TA;
                wgrad_acc<MT, TA, TB>(g, cT, n0, 1, aT, kg, KG);
#pragma unroll
                for (int i = 0; i < TA; ++i)
#pragma unroll
                    for (int j = 0; j < TB; ++j) {
                        const int n = n0 + i, k = kg + KG * j;
                        if (n < L.h2 && k < L.h1) gp[L.tW2 + n * L.h1 + k] += g[i][j];
                    }
            }
        }
        // gb2[n] += sum_m delta2[n][m]   (one warp per row, 8 rows in flight)
        for (int n = tid >> 5; n < L.h2; n += kThreads / 32) {
            float t = 0.0f;
            for (int m = (tid & 31) * 4; m < MT; m += 128) {
                const float4 d = *reinterpret_cast<const float4*>(cT + n * LDM + m);
                t += d.x + d.y + d.z + d.w;
            }
            t = warp_sum(t);
            if ((tid & 31) == 0) gp[L.tb2 + n] += t;
        }
        // delta1 = (delta2 W2) * act'(h1) -> aT (in place)
        zero_acc<TN>(acc);
        gemm_stream<H, MT, TN>(acc, cT, P + L.oW2N, H, wbuf);
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = TN * ng + j;
            const float4 h0 = *reinterpret_cast<const float4*>(aT + n * LDM + 4 * mg);
            const float4 h1 = *reinterpret_cast<const float4*>(aT + n * LDM + MT / 2 + 4 * mg);
            float4 v0, v1;
            v0.x = (1.0f - h0.x * h0.x) * acc[0][j]; v0.y = (1.0f - h0.y * h0.y) * acc[1][j];
            v0.z = (1.0f - h0.z * h0.z) * acc[2][j]; v0.w = (1.0f - h0.w * h0.w) * acc[3][j];
            v1.x = (1.0f - h1.x * h1.x) * acc[4][j]; v1.y = (1.0f - h1.y * h1.y) * acc[5][j];
            v1.z = (1.0f - h1.z * h1.z) * acc[6][j]; v1.w = (1.0f - h1.w * h1.w) * acc[7][j];
            *reinterpret_cast<float4*>(aT + n * LDM + 4 * mg) = v0;
            *reinterpret_cast<float4*>(aT + n * LDM + MT / 2 + 4 * mg) = v1;
        }
        // G1[n][k] += sum_m delta1[n][m] x~[k][m], chunk by chunk (x re-staged from L2 into the dead h2 buffer)
        for (int c = 0; c < nchunk; ++c) {
            __syncthreads();
            load_chunk<H, MT, MODE>(xs, a, base, c);
            __syncthreads();
            constexpr int KG = 16, TB = 2, NGT = kThreads / KG;
            constexpr int TA = (H / NGT < 8) ? H / NGT : 8, PASSES = H / (NGT * TA);
            const int kg = tid % KG, ngw = tid / KG;
#pragma unroll 1
            for (int p = 0; p < PASSES; ++p) {
                float g[TA][TB];
#pragma unroll
                for (int i = 0; i < TA; ++i) { g[i][0] = 0.0f; g[i][1] = 0.0f; }
                const int n0 = p * NGT * TA + ngw * TA;
                wgrad_acc<MT, TA, TB>(g, aT, n0, 1, xs, kg, KG);
#pragma unroll
                for (int i = 0; i < TA; ++i)
#pragma unroll
                    for (int j = 0; j < TB; ++j) {
                        const int n = n0 + i, k = kChunk * c + kg + KG * j;
                        if (n < L.h1 && k < L.K0) gp[L.tW1 + n * L.K0 + k] += g[i][j];
                    }
            }
            if (c == 0) {
                for (int n = tid >> 5; n < L.h1; n += kThreads / 32) {
                    float t = 0.0f;
                    for (int m = (tid & 31) * 4; m < MT; m += 128) {
                        const float4 d = *reinterpret_cast<const float4*>(aT + n * LDM + m);
                        t += d.x + d.y + d.z + d.w;
                    }
                    t = warp_sum(t);
                    if ((tid & 31) == 0) gp[L.tb1 + n] += t;
                }
            }
        }
    }

    if (MODE == MODE_EVAL || MODE == MODE_VPG) {
        const double t0 = block_sum(sum0, s_red);
        const double t1 = block_sum(sum1, s_red);
        if (tid == 0) { a.eval_partial[2 * blockIdx.x] = t0; a.eval_partial[2 * blockIdx.x + 1] = t1; }
    }
    if (MODE == MODE_VPG) {
        __syncthreads();                                   // ydT is free now: [warps][32] scratch
#pragma unroll 1
        for (int j = 0; j < A; ++j) {
            const float t = warp_sum(gs_acc[j]);
            if ((tid & 31) == 0) ydT[(tid >> 5) * 32 + j] = t;
        }
        __syncthreads();
        if (tid < A) {
            float t = 0.0f;
            for (int w = 0; w < kThreads / 32; ++w) t += ydT[w * 32 + tid];
            gp[L.tLS + tid] += t;
        }
    }
}

// host-side launcher, one translation unit per hidden width (mlp_h*.cu)
struct MlpLaunch {
    int grid;
    size_t smem;
};
int mlp_max_grid(int H, int mode, int YR, int num_sms);
cudaError_t launch_mlp(int H, int mode, const MlpArgs& args, int grid, cudaStream_t stream);
int mlp_tile_rows(int H);

}  // namespace mjb

// Gaussian *linear* policy (policies/gaussian_linear.py: FCNetwork(hidden_sizes=())) tile kernel:
//   mu = (x~ W^T + b) * out_scale + out_shift
// Same three modes as the MLP kernel (EVAL / VPG / FVP).  The path is HBM-bound (4*obs_dim bytes vs
// 4*obs*act flops per sample for the FVP): x is streamed once from HBM in 32-feature chunks for the
// forward product and re-read (L2-resident: one tile per CTA) for the weight-gradient product.
#include "kernels.h"

namespace mjb {

constexpr int LMT = 128;           // samples per tile
constexpr int LLDM = LMT + 4;

__device__ __forceinline__ int lin_load_chunk(float* xs, const LinArgs& a, long long base, int c) {
    const int cw = min(kChunk, a.L.K0 - kChunk * c);
    for (int f = threadIdx.x; f < LMT * cw; f += kThreads) {
        const int m = f / cw, kk = f - m * cw, k = kChunk * c + kk;
        const long long row = base + m;
        float v = 0.0f;
        if (row < a.n) {
            const long long r = a.idx ? (long long)a.idx[row] : row;
            v = (a.obs[r * a.L.K0 + k] - a.in_shift[k]) / (a.in_scale[k] + 1e-8f);
        }
        xs[kk * LLDM + m] = v;
    }
    for (int f = threadIdx.x; f < (kChunk - cw) * LMT; f += kThreads) xs[(cw + f / LMT) * LLDM + (f % LMT)] = 0.0f;
    return cw;
}

template <int AG, int MODE>
__global__ void __launch_bounds__(kThreads, 2) linear_kernel(const LinArgs a) {
    constexpr int AP = 8 * AG;
    __shared__ __align__(16) float xs[kChunk * LLDM];
    __shared__ __align__(16) float ydT[AP * LLDM];
    __shared__ float s_gs[32];
    __shared__ double s_red[32];
    const LinLayout& L = a.L;
    const int tid = threadIdx.x, A = L.A;
    const float* P = a.P;
    const int mq = tid % 32, ag = tid / 32;              // forward: 4 samples x AG outputs (a = ag + 8 i)
    const int nchunk = L.K0P / kChunk;
    if (tid < 32) s_gs[tid] = 0.0f;
    // log_std gradient: per-thread running sums, reduced once at the end in a fixed order (no floating-point atomics)
    float gs_acc[32];
    if (MODE == MODE_VPG) {
#pragma unroll
        for (int j = 0; j < 32; ++j) gs_acc[j] = 0.0f;
    }
    double sum0 = 0.0, sum1 = 0.0;
    float sum_ls = 0.0f;
    for (int j = 0; j < A; ++j) sum_ls += P[L.oLS + j];
    float* gp = (MODE != MODE_EVAL) ? a.gpartial + (size_t)blockIdx.x * a.gstride : nullptr;
    const float* Wf = (MODE == MODE_FVP) ? a.T : P;     // FVP forward runs with the tangent weights

    const long long n_tiles = (a.n + LMT - 1) / LMT;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long base = tile * LMT;
        float4 y[AG];
#pragma unroll
        for (int i = 0; i < AG; ++i) y[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = 0; c < nchunk; ++c) {
            __syncthreads();
            const int cw = lin_load_chunk(xs, a, base, c);
            __syncthreads();
            const float* w = Wf + L.oWT + (size_t)(kChunk * c) * AP + ag;
#pragma unroll 4
            for (int kk = 0; kk < cw; ++kk) {
                const float4 h = *reinterpret_cast<const float4*>(xs + kk * LLDM + 4 * mq);
#pragma unroll
                for (int i = 0; i < AG; ++i) {
                    const float wv = __ldg(w + kk * AP + 8 * i);
                    y[i].x = fmaf(h.x, wv, y[i].x); y[i].y = fmaf(h.y, wv, y[i].y);
                    y[i].z = fmaf(h.z, wv, y[i].z); y[i].w = fmaf(h.w, wv, y[i].w);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < AG; ++i) {
            const int o = ag + 8 * i;
            const float b = Wf[L.ob + o];
            float4 v = y[i];
            if (MODE == MODE_FVP) {
                float f = 0.0f;
                if (o < A) {
                    const float sd = expf(P[L.oLS + o]);
                    const float os = a.out_scale[o];
                    f = os * os * (2.0f / (2.0f * sd * sd + 1e-8f));
                }
                const long long r0 = base + 4 * mq;
                v.x = (r0 + 0 < a.n) ? f * (v.x + b) : 0.0f; v.y = (r0 + 1 < a.n) ? f * (v.y + b) : 0.0f;
                v.z = (r0 + 2 < a.n) ? f * (v.z + b) : 0.0f; v.w = (r0 + 3 < a.n) ? f * (v.w + b) : 0.0f;
            } else {
                v.x += b; v.y += b; v.z += b; v.w += b;
            }
            *reinterpret_cast<float4*>(ydT + o * LLDM + 4 * mq) = v;
        }
        __syncthreads();
        if (MODE == MODE_EVAL || MODE == MODE_VPG) {
            if (tid < LMT) {
                const long long row = base + tid;
                if (row < a.n) {
                    const float w = a.weight ? a.weight[row] : 0.0f;
                    float z2 = 0.0f, kl = 0.0f;
                    float zz[32];
#pragma unroll 1
                    for (int j = 0; j < A; ++j) {
                        const float mu = ydT[j * LLDM + tid] * a.out_scale[j] + a.out_shift[j];
                        const float s = P[L.oLS + j];
                        const float sd = expf(s);
                        const float z = (a.act[row * A + j] - mu) / sd;
                        z2 += z * z;
                        if (MODE == MODE_VPG) { zz[j] = z; ydT[j * LLDM + tid] = z / sd * a.out_scale[j]; }
                        if (a.old_flags & OLD_WRITE) a.mu_old[row * A + j] = mu;
                        if (a.old_flags & OLD_READ) {
                            const float so = a.old_log_std[j];
                            const float sdo = expf(so);
                            const float dm = a.mu_old[row * A + j] - mu;
                            kl += (dm * dm + sdo * sdo - sd * sd) / (2.0f * sd * sd + 1e-8f) + s - so;
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
                            ydT[j * LLDM + tid] *= coef;
                            gs_acc[j] += coef * (zz[j] * zz[j] - 1.0f);
                        }
                    }
                } else if (MODE == MODE_VPG) {
                    for (int j = 0; j < A; ++j) ydT[j * LLDM + tid] = 0.0f;
                }
            }
        }
        if (MODE == MODE_EVAL) continue;
        __syncthreads();
        // G[a][k] += sum_m dy[a][m] x~[k][m]  (x chunk re-staged; L2 hit), gb[a] += sum_m dy[a][m]
        if (tid < A) {
            float t = 0.0f;
            for (int m = 0; m < LMT; ++m) t += ydT[tid * LLDM + m];
            gp[L.tb + tid] += t;
        }
        const int kg = tid % 32, ngw = tid / 32;
        for (int c = 0; c < nchunk; ++c) {
            __syncthreads();
            lin_load_chunk(xs, a, base, c);
            __syncthreads();
            float g[AG][1];
#pragma unroll
            for (int i = 0; i < AG; ++i) g[i][0] = 0.0f;
            wgrad_acc<LMT, AG, 1>(g, ydT, ngw, 8, xs, kg, 0);
            const int k = kChunk * c + kg;
#pragma unroll
            for (int i = 0; i < AG; ++i) {
                const int o = ngw + 8 * i;
                if (o < A && k < L.K0) gp[L.tW + o * L.K0 + k] += g[i][0];
            }
        }
    }
    if (MODE == MODE_EVAL || MODE == MODE_VPG) {
        const double t0 = block_sum(sum0, s_red);
        const double t1 = block_sum(sum1, s_red);
        if (tid == 0) { a.eval_partial[2 * blockIdx.x] = t0; a.eval_partial[2 * blockIdx.x + 1] = t1; }
    }
    if (MODE == MODE_VPG) {
        __syncthreads();                                   // xs is free now: [warps][32] scratch
#pragma unroll 1
        for (int j = 0; j < A; ++j) {
            const float t = warp_sum(gs_acc[j]);
            if ((tid & 31) == 0) xs[(tid >> 5) * 32 + j] = t;
        }
        __syncthreads();
        if (tid < A) {
            float t = 0.0f;
            for (int w = 0; w < kThreads / 32; ++w) t += xs[w * 32 + tid];
            gp[L.tLS + tid] += t;
        }
    }
}

template <int AG>
static cudaError_t launch_lin_ag(int mode, const LinArgs& args, int grid, cudaStream_t s) {
    switch (mode) {
        case MODE_EVAL: linear_kernel<AG, MODE_EVAL><<<grid, kThreads, 0, s>>>(args); break;
        case MODE_VPG: linear_kernel<AG, MODE_VPG><<<grid, kThreads, 0, s>>>(args); break;
        case MODE_FVP: linear_kernel<AG, MODE_FVP><<<grid, kThreads, 0, s>>>(args); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_linear(int mode, const LinArgs& args, int grid, cudaStream_t s) {
    switch (args.L.AP / 8) {
        case 1: return launch_lin_ag<1>(mode, args, grid, s);
        case 2: return launch_lin_ag<2>(mode, args, grid, s);
        case 3: return launch_lin_ag<3>(mode, args, grid, s);
        case 4: return launch_lin_ag<4>(mode, args, grid, s);
    }
    return cudaErrorInvalidValue;
}

}  // namespace mjb

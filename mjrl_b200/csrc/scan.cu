// Discounted-sum scans over ragged trajectories and the batch statistics around them.
//   compute_returns / discount_sum   utils/process_samples.py:3-5,37-44
//   compute_advantages (GAE)         utils/process_samples.py:7-35
//   whitening                        algos/batch_reinforce.py:185
// The recurrences y_t = x_t + g*y_{t+1} are evaluated *sequentially per path in fp64 with separate
// multiply and add* (no FMA contraction), i.e. in exactly the reference's operation order, so returns and
// advantages are bit-identical to numpy given the same inputs.  Parallelism is across paths (lane = path); the data
// moves through shared-memory tiles with coalesced 256-byte segments (see "tiled, coalesced scans" below).
#include "kernels.h"

namespace mjb {

__global__ void f64_to_f32_kernel(const double* __restrict__ src, float* __restrict__ dst, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = (float)src[i];
}
void launch_f64_to_f32(const double* src, float* dst, long long n, cudaStream_t s) {
    if (n <= 0) return;
    int grid = (int)min((long long)148 * 8, (n + 255) / 256);
    f64_to_f32_kernel<<<grid, 256, 0, s>>>(src, dst, n);
}

// Device-resident rollouts [n_traj][H][width] (model-based rollouts, model_accel/sampling.py:16-90) -> packed rows of the
// valid prefixes in path order: trajectory i contributes len[i] rows at row offset path_off[i].  One CTA per trajectory.
template <typename T, typename D>
__global__ void pack_rollouts_kernel(const T* __restrict__ src, int H, int width, const int* __restrict__ path_off, int n_traj,
                                     D* __restrict__ dst) {
    for (int p = blockIdx.x; p < n_traj; p += gridDim.x) {
        const long long o = path_off[p], n = (long long)(path_off[p + 1] - path_off[p]) * width;
        const T* s = src + (size_t)p * H * width;
        D* d = dst + o * width;
        for (long long i = threadIdx.x; i < n; i += blockDim.x) d[i] = (D)s[i];
    }
}
void launch_pack_rollouts(const void* src, int is_f64, int H, int width, const int* path_off, int n_traj, void* dst, int dst_f64,
                          cudaStream_t s) {
    if (n_traj <= 0) return;
    const int grid = min(n_traj, 148 * 8);
    if (is_f64 && dst_f64) pack_rollouts_kernel<double, double><<<grid, 256, 0, s>>>((const double*)src, H, width, path_off, n_traj, (double*)dst);
    else if (is_f64) pack_rollouts_kernel<double, float><<<grid, 256, 0, s>>>((const double*)src, H, width, path_off, n_traj, (float*)dst);
    else if (dst_f64) pack_rollouts_kernel<float, double><<<grid, 256, 0, s>>>((const float*)src, H, width, path_off, n_traj, (double*)dst);
    else pack_rollouts_kernel<float, float><<<grid, 256, 0, s>>>((const float*)src, H, width, path_off, n_traj, (float*)dst);
}

__global__ void tstep_kernel(const int* __restrict__ path_off, int n_paths, int* __restrict__ tstep) {
    for (int p = blockIdx.x; p < n_paths; p += gridDim.x) {
        const int o = path_off[p], T = path_off[p + 1] - o;
        for (int t = threadIdx.x; t < T; t += blockDim.x) tstep[o + t] = t;
    }
}
void launch_tstep(const int* path_off, int n_paths, int* tstep, cudaStream_t s) {
    if (n_paths <= 0) return;
    tstep_kernel<<<min(n_paths, 148 * 8), 128, 0, s>>>(path_off, n_paths, tstep);
}

// ---- tiled, coalesced scans ------------------------------------------------------------------------------------------
// One warp owns 32 consecutive paths (lane = path) and walks them backwards in chunks of 32 steps.  A chunk is staged
// through shared memory: for every path the 32 lanes load its 32 consecutive doubles (one coalesced 256-byte segment),
// lane j then runs path j's recurrence over the staged row -- sequentially, fp64, separate multiply and add, i.e. the
// reference's operation order -- and the results leave through the same tile with coalesced stores.  The loads of the
// NEXT chunk are issued into registers before the recurrence runs, so their latency hides under the dependent chain
// (32 steps x DMUL+DADD).  The chain itself is the floor: bit-exact parity forbids re-associating it, so a path costs
// T x ~20 cycles no matter how it is fed -- coalescing removes the 32-byte-sector waste of one-thread-per-path loads
// (16 MB in 167 us before), it cannot beat the recurrence (1000 steps ~ 10 us; a batch of few very long paths stays
// serial in its paths).
constexpr int SC = 32;                       // steps per staged chunk
constexpr unsigned FULL = 0xffffffffu;

// ragged geometry of lane `lane`'s own path and of path j of this warp (broadcast)
struct WarpPaths {
    int o, T, nchunk;
    __device__ __forceinline__ void init(const int* __restrict__ path_off, int n_paths, int p) {
        o = p < n_paths ? path_off[p] : 0;
        T = p < n_paths ? path_off[p + 1] - o : 0;
        int tm = T;
        for (int d = 16; d > 0; d >>= 1) tm = max(tm, __shfl_xor_sync(FULL, tm, d));
        nchunk = (tm + SC - 1) / SC;
    }
};

// returns: reverse scan y_t = r_t + gamma*y_{t+1}
__global__ void __launch_bounds__(32) returns_kernel(const double* __restrict__ rew, const int* __restrict__ path_off,
                                                     int n_paths, double gamma, double* __restrict__ ret) {
    __shared__ double tile[32][SC + 1];
    const int lane = threadIdx.x;
    WarpPaths w;
    w.init(path_off, n_paths, blockIdx.x * 32 + lane);
    // chunk c of path j covers steps [T_j - (c+1) SC, T_j - c SC) clipped at 0; lane l holds step T_j - (c+1) SC + l
    auto load_chunk = [&](int c, double (&buf)[32]) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int oj = __shfl_sync(FULL, w.o, j), Tj = __shfl_sync(FULL, w.T, j);
            const int t = Tj - (c + 1) * SC + lane;
            buf[j] = (t >= 0 && Tj - c * SC > 0) ? rew[oj + t] : 0.0;
        }
    };
    double buf[32];
    if (w.nchunk > 0) load_chunk(0, buf);
    double run = 0.0;
    for (int c = 0; c < w.nchunk; ++c) {
#pragma unroll
        for (int j = 0; j < 32; ++j) tile[j][lane] = buf[j];
        __syncwarp();
        if (c + 1 < w.nchunk) load_chunk(c + 1, buf);            // in flight under the recurrence below
        const int hi = w.T - c * SC;
        if (hi > 0) {
            const int first = max(SC - hi, 0);                    // tile index of step 0 when the chunk is clipped
            for (int k = SC - 1; k >= first; --k) {
                run = __dadd_rn(tile[lane][k], __dmul_rn(gamma, run));
                tile[lane][k] = run;
            }
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int oj = __shfl_sync(FULL, w.o, j), Tj = __shfl_sync(FULL, w.T, j);
            const int t = Tj - (c + 1) * SC + lane;
            if (t >= 0 && Tj - c * SC > 0) ret[oj + t] = tile[j][lane];
        }
        __syncwarp();
    }
}
void launch_returns(const double* rew, const int* path_off, int n_paths, double gamma, double* ret, cudaStream_t s) {
    if (n_paths <= 0) return;
    returns_kernel<<<(n_paths + 31) / 32, 32, 0, s>>>(rew, path_off, n_paths, gamma, ret);
}

// per-path undiscounted return: forward sum in Python's sum() order (batch_reinforce.py:188), same staging (forward)
__global__ void __launch_bounds__(32) path_sums_kernel(const double* __restrict__ rew, const int* __restrict__ path_off,
                                                       int n_paths, double* __restrict__ path_ret) {
    __shared__ double tile[32][SC + 1];
    const int lane = threadIdx.x, p = blockIdx.x * 32 + lane;
    WarpPaths w;
    w.init(path_off, n_paths, p);
    double tot = 0.0;
    for (int c = 0; c < w.nchunk; ++c) {
#pragma unroll 8
        for (int j = 0; j < 32; ++j) {
            const int oj = __shfl_sync(FULL, w.o, j), Tj = __shfl_sync(FULL, w.T, j);
            const int t = c * SC + lane;
            tile[j][lane] = t < Tj ? rew[oj + t] : 0.0;
        }
        __syncwarp();
        const int cnt = min(SC, w.T - c * SC);
        for (int k = 0; k < cnt; ++k) tot = __dadd_rn(tot, tile[lane][k]);
        __syncwarp();
    }
    if (p < n_paths) path_ret[p] = tot;
}
void launch_path_sums(const double* rew, const int* path_off, int n_paths, double* path_ret, cudaStream_t s) {
    if (n_paths <= 0) return;
    path_sums_kernel<<<(n_paths + 31) / 32, 32, 0, s>>>(rew, path_off, n_paths, path_ret);
}

// GAE (process_samples.py:7-35) incl. the fp32 / fp64 bootstrap asymmetry (A5, A6); same staging as the returns
__global__ void __launch_bounds__(32) advantages_kernel(const double* __restrict__ rew, const float* __restrict__ base,
                                                        const double* __restrict__ ret, const int* __restrict__ path_off,
                                                        const unsigned char* __restrict__ terminated, int n_paths,
                                                        double gamma, double gamma_lam, int use_gae,
                                                        double* __restrict__ adv) {
    __shared__ double tile[32][SC + 1];
    __shared__ float tb[32][SC + 1];
    const int lane = threadIdx.x, p = blockIdx.x * 32 + lane;
    WarpPaths w;
    w.init(path_off, n_paths, p);
    const bool term = p < n_paths ? terminated[p] != 0 : false;
    const float gf = (float)gamma;
    double run = 0.0;
    float b_next = w.T > 0 ? base[w.o + w.T - 1] : 0.0f;             // bootstrap with the last *visited* state (:25)
    for (int c = 0; c < w.nchunk; ++c) {
#pragma unroll 8
        for (int j = 0; j < 32; ++j) {
            const int oj = __shfl_sync(FULL, w.o, j), Tj = __shfl_sync(FULL, w.T, j);
            const int t = Tj - (c + 1) * SC + lane;
            const bool in = t >= 0 && Tj - c * SC > 0;
            tile[j][lane] = in ? (use_gae ? rew[oj + t] : ret[oj + t]) : 0.0;
            tb[j][lane] = in ? base[oj + t] : 0.0f;
        }
        __syncwarp();
        const int hi = w.T - c * SC;
        if (hi > 0) {
            const int first = max(SC - hi, 0);
            for (int k = SC - 1; k >= first; --k) {
                const float b = tb[lane][k];
                if (!use_gae) {                                       // returns - baseline (process_samples.py:11-13)
                    tile[lane][k] = __dsub_rn(tile[lane][k], (double)b);
                    continue;
                }
                const bool last = (c == 0 && k == SC - 1);           // t == T - 1
                double nxt;
                if (term) nxt = last ? __dmul_rn(gamma, 0.0) : __dmul_rn(gamma, (double)b_next);   // b1 promoted to fp64
                else      nxt = (double)__fmul_rn(gf, b_next);                                     // b1 stays fp32
                const double td = __dsub_rn(__dadd_rn(tile[lane][k], nxt), (double)b);
                run = __dadd_rn(td, __dmul_rn(gamma_lam, run));
                tile[lane][k] = run;
                b_next = b;
            }
        }
        __syncwarp();
#pragma unroll 8
        for (int j = 0; j < 32; ++j) {
            const int oj = __shfl_sync(FULL, w.o, j), Tj = __shfl_sync(FULL, w.T, j);
            const int t = Tj - (c + 1) * SC + lane;
            if (t >= 0 && Tj - c * SC > 0) adv[oj + t] = tile[j][lane];
        }
        __syncwarp();
    }
}
void launch_advantages(const double* rew, const float* base, const double* ret, const int* path_off,
                       const unsigned char* terminated, int n_paths, double gamma, double gamma_lam,
                       int use_gae, double* adv, cudaStream_t s) {
    if (n_paths <= 0) return;
    advantages_kernel<<<(n_paths + 31) / 32, 32, 0, s>>>(rew, base, ret, path_off, terminated, n_paths, gamma,
                                                        gamma_lam, use_gae, adv);
}

// ---- deterministic two-stage moments: scratch[grid][2], out[2] -----------------------------------------
constexpr int kMomGrid = 296;
__global__ void moments_stage1(const double* __restrict__ x, long long n, const double* __restrict__ shift_dev,
                               double* __restrict__ scratch) {
    __shared__ double red[32];
    const double shift = shift_dev ? *shift_dev : 0.0;
    double s1 = 0.0, s2 = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double d = x[i] - shift;
        s1 += d; s2 += d * d;
    }
    s1 = block_sum(s1, red);
    s2 = block_sum(s2, red);
    if (threadIdx.x == 0) { scratch[2 * blockIdx.x] = s1; scratch[2 * blockIdx.x + 1] = s2; }
}
__global__ void moments_stage2(const double* __restrict__ scratch, int grid, double* __restrict__ out2) {
    __shared__ double red[32];
    double s1 = 0.0, s2 = 0.0;
    for (int i = threadIdx.x; i < grid; i += blockDim.x) { s1 += scratch[2 * i]; s2 += scratch[2 * i + 1]; }
    s1 = block_sum(s1, red);
    s2 = block_sum(s2, red);
    if (threadIdx.x == 0) { out2[0] = s1; out2[1] = s2; }
}
void launch_moments(const double* x, long long n, const double* shift_dev, double* scratch, double* out2,
                    cudaStream_t s) {
    moments_stage1<<<kMomGrid, 256, 0, s>>>(x, n, shift_dev, scratch);
    moments_stage2<<<1, 256, 0, s>>>(scratch, kMomGrid, out2);
}

// ---- batch statistics finalised ON the device (mjb_process_paths makes one host round trip, at its end) -----------
// mode 0: stats[0] = mom[0] * inv_n (mean);  mode 1: stats[1] = sqrt(mom[1] * inv_n) (population std about that mean)
__global__ void stats_finalize_kernel(const double* __restrict__ mom, double inv_n, double* __restrict__ stats, int mode) {
    if (mode == 0) stats[0] = mom[0] * inv_n;
    else stats[1] = sqrt(mom[1] * inv_n);
}
void launch_stats_finalize(const double* mom2, double inv_n, double* stats2, int mode, cudaStream_t s) {
    stats_finalize_kernel<<<1, 1, 0, s>>>(mom2, inv_n, stats2, mode);
}
// pass 0: out[0] = sum, out[1] = -min, out[2] = max of the per-path returns (local paths);
// pass 1: out[3] = sum (r - mean)^2 with mean = out[0] * inv_paths_global (out[0] all-reduced in between)
__global__ void path_stats_kernel(const double* __restrict__ pr, int n_paths, double inv_paths_global, double* __restrict__ out,
                                  int pass) {
    __shared__ double red[32];
    __shared__ double redm[64];
    const double mean = pass ? out[0] * inv_paths_global : 0.0;
    double s = 0.0, mn = INFINITY, mx = -INFINITY;
    for (int i = threadIdx.x; i < n_paths; i += blockDim.x) {
        const double v = pr[i];
        if (pass) { s += (v - mean) * (v - mean); }
        else { s += v; mn = fmin(mn, v); mx = fmax(mx, v); }
    }
    s = block_sum(s, red);
    if (pass) { if (threadIdx.x == 0) out[3] = s; return; }
    for (int d = 16; d > 0; d >>= 1) { mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, d)); mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, d)); }
    if ((threadIdx.x & 31) == 0) { redm[threadIdx.x >> 5] = mn; redm[32 + (threadIdx.x >> 5)] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { mn = fmin(mn, redm[w]); mx = fmax(mx, redm[32 + w]); }
        out[0] = s; out[1] = -mn; out[2] = mx;
    }
}
void launch_path_stats(const double* path_ret, int n_paths, double inv_paths_global, double* out4, int pass, cudaStream_t s) {
    path_stats_kernel<<<1, 256, 0, s>>>(path_ret, n_paths, inv_paths_global, out4, pass);
}

__global__ void whiten_kernel(const double* __restrict__ adv, long long n, const double* __restrict__ stats,
                              float* __restrict__ white) {
    const double mean = stats[0], denom = stats[1] + 1e-6;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        white[i] = (float)((adv[i] - mean) / denom);
}
void launch_whiten(const double* adv, long long n, const double* stats, float* white, cudaStream_t s) {
    if (n <= 0) return;
    whiten_kernel<<<(int)min((long long)148 * 8, (n + 255) / 256), 256, 0, s>>>(adv, n, stats, white);
}

__global__ void dapg_weights_kernel(const double* __restrict__ adv, long long n, long long n_demo,
                                    const double* __restrict__ stats, double lam, float* __restrict__ w) {
    const double mean = stats[0], sd = stats[1], denom = sd + 1e-6;
    const double sw = sd / denom + 1e-8;              // std of the whitened advantages + 1e-8 (dapg.py:74)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n + n_demo; i += (long long)gridDim.x * blockDim.x)
        w[i] = (i < n) ? (float)(1e-2 * (((adv[i] - mean) / denom) / sw)) : (float)(1e-2 * lam);
}
void launch_dapg_weights(const double* adv, long long n, long long n_demo, const double* stats, double lam,
                         float* w, cudaStream_t s) {
    dapg_weights_kernel<<<(int)min((long long)148 * 8, (n + n_demo + 255) / 256), 256, 0, s>>>(adv, n, n_demo, stats, lam, w);
}

// ---- VF error pieces in the reference's dtypes: errors = float32(returns) - pred ------------------------
__global__ void vf_error_stage1(const double* __restrict__ ret, const float* __restrict__ pred, long long n,
                                double* __restrict__ scratch) {
    __shared__ double red[32];
    double s1 = 0.0, s2 = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float r = (float)ret[i];
        const float e = r - pred[i];
        s1 += (double)(e * e); s2 += (double)(r * r);
    }
    s1 = block_sum(s1, red);
    s2 = block_sum(s2, red);
    if (threadIdx.x == 0) { scratch[2 * blockIdx.x] = s1; scratch[2 * blockIdx.x + 1] = s2; }
}
void launch_vf_error(const double* ret, const float* pred, long long n, double* scratch, double* out2, cudaStream_t s) {
    vf_error_stage1<<<kMomGrid, 256, 0, s>>>(ret, pred, n, scratch);
    moments_stage2<<<1, 256, 0, s>>>(scratch, kMomGrid, out2);
}

}  // namespace mjb

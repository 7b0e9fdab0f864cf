// Discounted-sum scans over ragged trajectories and the batch statistics around them.
//   compute_returns / discount_sum   utils/process_samples.py:3-5,37-44
//   compute_advantages (GAE)         utils/process_samples.py:7-35
//   whitening                        algos/batch_reinforce.py:185
// The recurrences y_t = x_t + g*y_{t+1} are evaluated *sequentially per path in fp64 with separate
// multiply and add* (no FMA contraction), i.e. in exactly the reference's operation order, so returns and
// advantages are bit-identical to numpy given the same inputs.  Parallelism is across paths: one thread per
// path; a warp's 32 paths each stream their own 128-byte lines, which stay L1-resident for 16 steps.
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

// returns: reverse scan y_t = r_t + gamma*y_{t+1}
__global__ void returns_kernel(const double* __restrict__ rew, const int* __restrict__ path_off, int n_paths,
                               double gamma, double* __restrict__ ret) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_paths) return;
    const int o = path_off[p], T = path_off[p + 1] - o;
    double run = 0.0;
    for (int t = T - 1; t >= 0; --t) {
        run = __dadd_rn(rew[o + t], __dmul_rn(gamma, run));
        ret[o + t] = run;
    }
}
void launch_returns(const double* rew, const int* path_off, int n_paths, double gamma, double* ret, cudaStream_t s) {
    if (n_paths <= 0) return;
    returns_kernel<<<(n_paths + 31) / 32, 32, 0, s>>>(rew, path_off, n_paths, gamma, ret);
}

// per-path undiscounted return: forward sum in Python's sum() order (batch_reinforce.py:188)
__global__ void path_sums_kernel(const double* __restrict__ rew, const int* __restrict__ path_off, int n_paths,
                                 double* __restrict__ path_ret) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_paths) return;
    const int o = path_off[p], T = path_off[p + 1] - o;
    double tot = 0.0;
    for (int t = 0; t < T; ++t) tot = __dadd_rn(tot, rew[o + t]);
    path_ret[p] = tot;
}
void launch_path_sums(const double* rew, const int* path_off, int n_paths, double* path_ret, cudaStream_t s) {
    if (n_paths <= 0) return;
    path_sums_kernel<<<(n_paths + 31) / 32, 32, 0, s>>>(rew, path_off, n_paths, path_ret);
}

__global__ void advantages_kernel(const double* __restrict__ rew, const float* __restrict__ base,
                                  const double* __restrict__ ret, const int* __restrict__ path_off,
                                  const unsigned char* __restrict__ terminated, int n_paths, double gamma,
                                  double gamma_lam, int use_gae, double* __restrict__ adv) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_paths) return;
    const int o = path_off[p], T = path_off[p + 1] - o;
    if (T <= 0) return;
    if (!use_gae) {                                   // returns - baseline (process_samples.py:11-13)
        for (int t = 0; t < T; ++t) adv[o + t] = __dsub_rn(ret[o + t], (double)base[o + t]);
        return;
    }
    const bool term = terminated[p] != 0;
    const float gf = (float)gamma;
    double run = 0.0;
    float b_next = base[o + T - 1];                   // bootstrap with the last *visited* state (:25)
    for (int t = T - 1; t >= 0; --t) {
        const float b = base[o + t];
        double nxt;
        if (term) nxt = (t == T - 1) ? __dmul_rn(gamma, 0.0) : __dmul_rn(gamma, (double)b_next);   // b1 promoted to fp64
        else      nxt = (double)__fmul_rn(gf, b_next);                                             // b1 stays fp32
        const double td = __dsub_rn(__dadd_rn(rew[o + t], nxt), (double)b);
        run = __dadd_rn(td, __dmul_rn(gamma_lam, run));
        adv[o + t] = run;
        b_next = b;
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

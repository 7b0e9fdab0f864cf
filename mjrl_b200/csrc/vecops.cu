// Small vector kernels around the tile kernels: parameter re-layout, deterministic reduction of the
// per-CTA gradient partials, and the device-resident conjugate-gradient state machine
// (utils/cg_solve.py:3-22).  d <= ~83k floats, so every vector op here is latency-bound; the CG update
// is one single-CTA kernel per iteration that fuses both dot products with the x/r/p updates.
#include "kernels.h"

namespace mjb {

// theta (reference flat layout) -> padded / transposed kernel layout (see PrepLayout)
__global__ void prep_mlp_kernel(const float* __restrict__ th, const PrepLayout L, float* __restrict__ out) {
    const int H = L.H;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L.total; i += gridDim.x * blockDim.x) {
        float v = 0.0f;
        if (i < L.ob1) {                                   // W1T [K0P][H]
            const int k = (i - L.oW1T) / H, n = (i - L.oW1T) % H;
            if (k < L.K0 && n < L.h1) v = th[L.tW1 + n * L.K0 + k];
        } else if (i < L.oW2T) {
            const int n = i - L.ob1;
            if (n < L.h1) v = th[L.tb1 + n];
        } else if (i < L.ob2) {                            // W2T [H(k)][H(n)]
            const int k = (i - L.oW2T) / H, n = (i - L.oW2T) % H;
            if (k < L.h1 && n < L.h2) v = th[L.tW2 + n * L.h1 + k];
        } else if (i < L.oW3T) {
            const int n = i - L.ob2;
            if (n < L.h2) v = th[L.tb2 + n];
        } else if (i < L.ob3) {                            // W3T [H(k)][AP]
            const int k = (i - L.oW3T) / L.AP, o = (i - L.oW3T) % L.AP;
            if (k < L.h2 && o < L.A) v = th[L.tW3 + o * L.h2 + k];
        } else if (i < L.oW2N) {
            const int o = i - L.ob3;
            if (o < L.A) v = th[L.tb3 + o];
        } else if (i < L.oW3N) {                           // W2N [H(n)][H(k)]
            const int n = (i - L.oW2N) / H, k = (i - L.oW2N) % H;
            if (n < L.h2 && k < L.h1) v = th[L.tW2 + n * L.h1 + k];
        } else if (i < L.oLS) {                            // W3N [YR][H(k)]
            const int o = (i - L.oW3N) / H, k = (i - L.oW3N) % H;
            if (o < L.A && k < L.h2) v = th[L.tW3 + o * L.h2 + k];
        } else if (i < L.oLS + L.AP) {
            const int o = i - L.oLS;
            if (o < L.A && L.d > L.tLS) v = th[L.tLS + o];
        }
        out[i] = v;
    }
}
void launch_prep_mlp(const float* theta, const PrepLayout& L, float* out, cudaStream_t s) {
    prep_mlp_kernel<<<(L.total + 255) / 256, 256, 0, s>>>(theta, L, out);
}

__global__ void prep_linear_kernel(const float* __restrict__ th, const LinLayout L, float* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L.total; i += gridDim.x * blockDim.x) {
        float v = 0.0f;
        if (i < L.ob) {
            const int k = i / L.AP, o = i % L.AP;
            if (k < L.K0 && o < L.A) v = th[L.tW + o * L.K0 + k];
        } else if (i < L.oLS) {
            const int o = i - L.ob;
            if (o < L.A) v = th[L.tb + o];
        } else if (i < L.oLS + L.AP) {
            const int o = i - L.oLS;
            if (o < L.A) v = th[L.tLS + o];
        }
        out[i] = v;
    }
}
void launch_prep_linear(const float* theta, const LinLayout& L, float* out, cudaStream_t s) {
    prep_linear_kernel<<<(L.total + 255) / 256, 256, 0, s>>>(theta, L, out);
}

__global__ void clamp_tail_kernel(float* th, int d, int A, float lo) {
    const int i = threadIdx.x;
    if (i < A) th[d - A + i] = fmaxf(th[d - A + i], lo);
}
void launch_clamp_tail(float* theta, int d, int A, float lo, cudaStream_t s) {
    clamp_tail_kernel<<<1, 32, 0, s>>>(theta, d, A, lo);
}

// out[i] = scale * sum_c partial[c][i]; FVP adds the data-free log_std block c(sigma) * v
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int grid, long long stride, int d,
                                       const double* __restrict__ scale_dev, float* __restrict__ out,
                                       const float* __restrict__ theta, const float* __restrict__ v, int tLS,
                                       int fvp_ls_block, const float* __restrict__ vscale) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d) return;
    // tensor-core FVP: the tangent was multiplied by the power of two vscale[0]; vscale[1] = 1/vscale[0] undoes it exactly
    const double sc0 = vscale ? scale_dev[0] * (double)vscale[1] : scale_dev[0];
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int c = 0;
    for (; c + 3 < grid; c += 4) {
        s0 += partial[(size_t)c * stride + i];
        s1 += partial[(size_t)(c + 1) * stride + i];
        s2 += partial[(size_t)(c + 2) * stride + i];
        s3 += partial[(size_t)(c + 3) * stride + i];
    }
    for (; c < grid; ++c) s0 += partial[(size_t)c * stride + i];
    float r = (float)((double)((s0 + s1) + (s2 + s3)) * sc0);
    if (fvp_ls_block && i >= tLS) {
        const float u = expf(2.0f * theta[i]);
        const float den = 2.0f * u + 1e-8f;
        r = v[i] * ((8.0f * u * u - 4.0f * u * 1e-8f) / (den * den));
        // under data parallelism every rank adds the same block once: scaled so the all-reduce sum is exact
        r = (float)((double)r * scale_dev[1]);
    }
    out[i] = r;
}
void launch_reduce_partials(const float* partial, int grid, long long stride, int d, const double* scale_dev,
                            float* out, const float* theta, const float* v, int tLS, int fvp_ls_block,
                            const float* vscale, cudaStream_t s) {
    reduce_partials_kernel<<<(d + 127) / 128, 128, 0, s>>>(partial, grid, stride, d, scale_dev, out, theta, v, tLS,
                                                         fvp_ls_block, vscale);
}

__global__ void reduce_eval_kernel(const double* __restrict__ partial, int grid, double* __restrict__ out2) {
    __shared__ double red[32];
    double s0 = 0.0, s1 = 0.0;
    for (int i = threadIdx.x; i < grid; i += blockDim.x) { s0 += partial[2 * i]; s1 += partial[2 * i + 1]; }
    s0 = block_sum(s0, red);
    s1 = block_sum(s1, red);
    if (threadIdx.x == 0) { out2[0] = s0; out2[1] = s1; }
}
void launch_reduce_eval(const double* partial, int grid, double* out2, cudaStream_t s) {
    reduce_eval_kernel<<<1, 256, 0, s>>>(partial, grid, out2);
}

// ---- conjugate gradient, state on device: st[0]=r.r  st[1]=done  st[2]=FVPs consumed  st[3]=scratch ----
constexpr int kCgThreads = 1024;
// Block-wide max of |v| -> the power of two that brings it into [1, 2): out[0] = scale, out[1] = 1/scale.  The
// tensor-core FVP multiplies its tangent by out[0] (two-term fp16 operands want O(1) magnitudes) and the reduction
// multiplies by out[1]; both are exact.  Fused here so that the CG iteration has no separate scale kernel.
__device__ __forceinline__ void block_pow2_scale(float mx, float* out, float* red) {
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    __syncthreads();
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

__global__ void cg_init_kernel(const float* __restrict__ b, float* x, float* r, float* p, int d, double* st,
                               float* vscale) {
    __shared__ double red[32];
    __shared__ float redf[32];
    double s = 0.0;
    float mx = 0.0f;
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
        const float v = b[i];
        x[i] = 0.0f; r[i] = v; p[i] = v;
        s += (double)v * (double)v;
        mx = fmaxf(mx, fabsf(v));
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) { st[0] = (double)(float)s; st[1] = 0.0; st[2] = 0.0; }
    if (vscale) block_pow2_scale(mx, vscale, redf);
}
void launch_cg_init(const float* b, float* x, float* r, float* p, int d, double* st, float* vscale, cudaStream_t s) {
    cg_init_kernel<<<1, kCgThreads, 0, s>>>(b, x, r, p, d, st, vscale);
}

// One iteration of cg_solve.py:10-20 given Fp (already all-reduced, undamped): z = Fp + damping*p, ...
// Scalars are rounded to fp32 where the reference's numpy arithmetic is fp32 (A2).
__global__ void cg_update_kernel(const float* __restrict__ Fp, float damping, float tol, float* x, float* r,
                                 float* p, int d, double* st, float* vscale) {
    __shared__ double red[32];
    __shared__ float redf[32];
    __shared__ float s_alpha, s_mu;
    if (st[1] != 0.0) return;                           // converged earlier: x frozen (break at :19-20)
    double pz = 0.0;
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
        const float z = fmaf(damping, p[i], Fp[i]);
        pz += (double)p[i] * (double)z;
    }
    pz = block_sum(pz, red);
    if (threadIdx.x == 0) s_alpha = (float)st[0] / (float)pz;
    __syncthreads();
    const float alpha = s_alpha;
    double rr = 0.0;
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
        const float z = fmaf(damping, p[i], Fp[i]);
        x[i] = fmaf(alpha, p[i], x[i]);
        const float rn = fmaf(-alpha, z, r[i]);
        r[i] = rn;
        rr += (double)rn * (double)rn;
    }
    rr = block_sum(rr, red);
    if (threadIdx.x == 0) s_mu = (float)rr / (float)st[0];
    __syncthreads();
    const float mu = s_mu;
    float mx = 0.0f;
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
        const float pn = fmaf(mu, p[i], r[i]);
        p[i] = pn;
        mx = fmaxf(mx, fabsf(pn));
    }
    if (threadIdx.x == 0) {
        st[0] = (double)(float)rr;
        st[2] += 1.0;
        if ((float)rr < tol) st[1] = 1.0;
    }
    if (vscale) block_pow2_scale(mx, vscale, redf);     // scale of the NEXT Fisher product's tangent
}
void launch_cg_update(const float* Fp, float damping, float tol, float* x, float* r, float* p, int d, double* st,
                      float* vscale, cudaStream_t s) {
    cg_update_kernel<<<1, kCgThreads, 0, s>>>(Fp, damping, tol, x, r, p, d, st, vscale);
}

__global__ void dot_kernel(const float* __restrict__ a, const float* __restrict__ b, int d, double* out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < d; i += blockDim.x) s += (double)a[i] * (double)b[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) *out = s;
}
void launch_dot(const float* a, const float* b, int d, double* out, cudaStream_t s) {
    dot_kernel<<<1, kCgThreads, 0, s>>>(a, b, d, out);
}

// out = clamp_logstd(theta + float(alpha*scale) * x)          (npg_cg.py:137-139, gaussian_mlp.py:73-75)
__global__ void axpy_clamp_kernel(const float* __restrict__ theta, const float* __restrict__ x,
                                  const double* __restrict__ alpha_dev, double alpha_scale, int d, int A, float lo,
                                  float* __restrict__ out) {
    const float al = (float)((*alpha_dev) * alpha_scale);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d; i += gridDim.x * blockDim.x) {
        float v = fmaf(al, x[i], theta[i]);
        if (i >= d - A) v = fmaxf(v, lo);
        out[i] = v;
    }
}
void launch_axpy_clamp(const float* theta, const float* x, const double* alpha_dev, double alpha_scale, int d,
                       int A, float lo, float* out, cudaStream_t s) {
    axpy_clamp_kernel<<<(d + 255) / 256, 256, 0, s>>>(theta, x, alpha_dev, alpha_scale, d, A, lo, out);
}

__global__ void scale_kernel(float* x, int d, float f) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d; i += gridDim.x * blockDim.x) x[i] *= f;
}
void launch_scale(float* x, int d, float f, cudaStream_t s) { scale_kernel<<<(d + 255) / 256, 256, 0, s>>>(x, d, f); }

}  // namespace mjb

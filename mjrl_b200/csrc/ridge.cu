// Ridge-regression baselines on the device-resident batch: the reference's LinearBaseline / QuadraticBaseline
// (baselines/linear_baseline.py:11-60, baselines/quadratic_baseline.py:11-68) build an [N x K] float64 feature matrix on the
// host and solve (F^T F + reg I) c = F^T y.  Everything that scales with N runs here, in float64 like the reference:
//
//   * ridge_scale_kernel writes T = {clip(o)/10 ..., 1, al, al^2, al^3, al^4, y, 0} per sample (obs_dim + 7 float64 columns);
//   * ridge_gram_kernel accumulates the Gram matrix of the AUGMENTED matrix [F | y], i.e. F^T F, F^T y and y^T y together.
//     The feature matrix is never materialised: every feature is the product of two columns of T (linear feature =
//     column x 1, quadratic feature = column x column, bias = 1 x 1), looked up through a per-feature (a, b) table.  Output
//     blocks of 64 x 64 (upper triangle) x sample splits; the splits are summed in a fixed order (bit-identical repeats).
//     Up to 32 columns: a warp-per-sample-stream kernel without any CTA-wide staging.
//   * ridge_predict_kernel: predictions F c for every resident sample into the engine's baseline buffer (read by the GAE
//     kernel), plus sum (y - F c)^2 for the reference's error_before / error_after.
//
// The K x K solve itself (np.linalg.lstsq with the reference's retry-on-NaN loop) stays on the host: it is N-independent.
#include <cuda_runtime.h>

#include "kernels.h"

namespace mjb {

namespace {

constexpr int RB = 64;          // output block edge
constexpr int RCH = 32;         // samples per staged chunk

// T[row][c], c < tile_cols: {clip(o)/10 (obs_dim) | 1 | al al^2 al^3 al^4 | y | 0} in float64 -- the scaled observations, NOT the
// feature matrix (which for the quadratic baseline is n/2 times wider).  Written once per Gram / predict call; every feature
// is the product of two of its columns.  (Computing these columns inside the Gram kernel made it 10x slower: the float64
// division by 10 -- kept because the reference divides -- was redone for every block pair.)
__global__ void ridge_scale_kernel(const RidgeArgs a, double* __restrict__ T) {
    const int TC = a.tile_cols, od = a.obs_dim;
    const long long total = a.n * TC;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long row = idx / TC;
        const int c = (int)(idx - row * TC);
        double v = 0.0;
        if (c < od) {
            v = (double)a.obs[row * od + c];
            v = fmin(fmax(v, -10.0), 10.0) / 10.0;
        } else if (c == od) {
            v = 1.0;
        } else if (c <= od + 4) {
            const double al = (double)a.tstep[row] / 1000.0;
            v = al;
            for (int q = 1; q < c - od; ++q) v *= al;                // al^k, k = 1..4 (linear_baseline.py:31-35)
        } else if (c == od + 5) {
            v = a.ret ? a.ret[row] : 0.0;
        }                                                            // c == od + 6: the zero column (padding features)
        T[idx] = v;
    }
}

__global__ void __launch_bounds__(256, 2) ridge_gram_kernel(const RidgeArgs a, const double* __restrict__ T) {
    __shared__ __align__(16) double phi_i[RCH * RB];
    __shared__ __align__(16) double phi_j[RCH * RB];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    // block pair (bi <= bj) of this CTA
    int bi = 0, p = blockIdx.x;
    while (p >= a.nb - bi) { p -= a.nb - bi; ++bi; }
    const int bj = bi + p;
    // feature staging: thread -> feature f = tid & 63 of block bi and of block bj, samples (tid >> 6) + 4 k of the chunk
    const int sf = tid & 63, ss = tid >> 6;
    const short zc = (short)(a.obs_dim + 6);                         // the zero column: features beyond K
    const short2 pi = RB * bi + sf < a.K ? a.ab[RB * bi + sf] : make_short2(zc, zc);
    const short2 pj = RB * bj + sf < a.K ? a.ab[RB * bj + sf] : make_short2(zc, zc);
    double acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
    const long long lo = a.n * (long long)blockIdx.y / a.splits, hi = a.n * (long long)(blockIdx.y + 1) / a.splits;
    const int TC = a.tile_cols;
    for (long long n0 = lo; n0 < hi; n0 += RCH) {
        double vi[RCH / 4], vj[RCH / 4];
#pragma unroll
        for (int k = 0; k < RCH / 4; ++k) {                          // every feature value of the chunk is computed once
            const long long row = n0 + ss + 4 * k;
            const double* t = T + row * TC;
            const bool in = row < hi;
            vi[k] = in ? t[pi.x] * t[pi.y] : 0.0;
            vj[k] = in ? t[pj.x] * t[pj.y] : 0.0;
        }
        __syncthreads();                                             // the previous chunk has been consumed
#pragma unroll
        for (int k = 0; k < RCH / 4; ++k) {
            phi_i[(ss + 4 * k) * RB + sf] = vi[k];
            phi_j[(ss + 4 * k) * RB + sf] = vj[k];
        }
        __syncthreads();
#pragma unroll 4
        for (int s = 0; s < RCH; ++s) {
            const double2 i0 = *reinterpret_cast<const double2*>(phi_i + s * RB + 4 * ty);
            const double2 i1 = *reinterpret_cast<const double2*>(phi_i + s * RB + 4 * ty + 2);
            const double2 j0 = *reinterpret_cast<const double2*>(phi_j + s * RB + 4 * tx);
            const double2 j1 = *reinterpret_cast<const double2*>(phi_j + s * RB + 4 * tx + 2);
            const double fi[4] = {i0.x, i0.y, i1.x, i1.y}, fj[4] = {j0.x, j0.y, j1.x, j1.y};
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = fma(fi[r], fj[c], acc[r][c]);
        }
    }
    double* out = a.partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (RB * RB);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) out[(4 * ty + r) * RB + 4 * tx + c] = acc[r][c];
}

// Up to 32 (augmented) features: no CTA-wide staging at all.  Every warp streams its own samples; lane l computes feature l
// of the sample from two columns of the T row, the 32 values go through a per-warp shared-memory line, and lane l
// accumulates row l of the 32 x 32 Gram block (32 fp64 registers) from broadcast reads.  One [32][32] partial per warp,
// summed afterwards in warp order.
__global__ void __launch_bounds__(256) ridge_gram_small_kernel(const RidgeArgs a, const double* __restrict__ T) {
    __shared__ __align__(16) double phi[8][2][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int TC = a.tile_cols;
    const int gw = blockIdx.x * 8 + warp, nw = gridDim.x * 8;
    const short2 ab = lane < a.K ? a.ab[lane] : make_short2((short)(a.obs_dim + 6), (short)(a.obs_dim + 6));
    double acc[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = 0.0;
    const long long lo = a.n * (long long)gw / nw, hi = a.n * (long long)(gw + 1) / nw;
    double f_next = lo < hi ? T[lo * TC + ab.x] * T[lo * TC + ab.y] : 0.0;
    for (long long row = lo; row < hi; ++row) {
        const double f = f_next;
        if (row + 1 < hi) f_next = T[(row + 1) * TC + ab.x] * T[(row + 1) * TC + ab.y];     // in flight under the 32 FMAs below
        double* line = phi[warp][row & 1];                           // two lines: no second __syncwarp per sample
        line[lane] = f;
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
            const double2 p = *reinterpret_cast<const double2*>(line + c);
            acc[c] = fma(f, p.x, acc[c]);
            acc[c + 1] = fma(f, p.y, acc[c + 1]);
        }
    }
    double* out = a.partial + (size_t)gw * 1024 + lane * 32;
#pragma unroll
    for (int c = 0; c < 32; ++c) out[c] = acc[c];
}

// one CTA per output element: 128 threads sum strided slices of the per-warp partials, then a fixed-order tree
__global__ void __launch_bounds__(128) ridge_gram_small_reduce_kernel(const double* __restrict__ partial, int nw, int K, double* __restrict__ G) {
    __shared__ double red[128];
    const int i = blockIdx.x / K, j = blockIdx.x - i * K;
    double s = 0.0;
    for (int w = threadIdx.x; w < nw; w += 128) s += partial[(size_t)w * 1024 + i * 32 + j];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) G[blockIdx.x] = red[0];
}

// G[i][j] = sum over the splits (fixed order) of the block partials; both triangles are written
__global__ void ridge_gram_reduce_kernel(const double* __restrict__ partial, int nb, int npairs, int splits, int K, double* __restrict__ G) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)K * K) return;
    const int i = (int)(idx / K), j = (int)(idx - (long long)i * K);
    if (j < i) return;
    const int bi = i / RB, bj = j / RB;
    int p = 0;
    for (int q = 0; q < bi; ++q) p += nb - q;
    p += bj - bi;
    const double* src = partial + (size_t)p * (RB * RB) + (i - RB * bi) * RB + (j - RB * bj);
    double s = 0.0;
    for (int sp = 0; sp < splits; ++sp) s += src[(size_t)sp * npairs * (RB * RB)];
    G[(size_t)i * K + j] = s;
    G[(size_t)j * K + i] = s;
}

// one warp per sample: pred = sum_f coeff[f] * T[a_f] * T[b_f]; base[row] = (float)pred; err partial = sum (y - pred)^2
__global__ void __launch_bounds__(256) ridge_predict_kernel(const RidgeArgs a, const double* __restrict__ T, const double* __restrict__ coeff,
                                                           int Kfeat, float* __restrict__ base, double* __restrict__ err_partial) {
    __shared__ double red[8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int TC = a.tile_cols;
    double err = 0.0;
    for (long long row = (long long)blockIdx.x * 8 + warp; row < a.n; row += (long long)gridDim.x * 8) {
        const double* t = T + row * TC;
        double s = 0.0;
        for (int f = lane; f < Kfeat; f += 32) {
            const short2 ab = a.ab[f];
            s = fma(coeff[f], t[ab.x] * t[ab.y], s);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) {
            base[row] = (float)s;
            if (a.ret) { const double d = a.ret[row] - s; err = fma(d, d, err); }
        }
    }
    if (lane == 0) red[warp] = err;
    __syncthreads();
    if (tid == 0 && err_partial) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[w];
        err_partial[blockIdx.x] = s;
    }
}

}  // namespace

int ridge_blocks(int K) { return (K + RB - 1) / RB; }

cudaError_t launch_ridge_scale(const RidgeArgs& a, double* T, cudaStream_t s) {
    ridge_scale_kernel<<<148 * 8, 256, 0, s>>>(a, T);
    return cudaGetLastError();
}

cudaError_t launch_ridge_gram(const RidgeArgs& a, const double* T, double* G, cudaStream_t s) {
    if (a.K <= 32) {                                                 // splits = number of CTAs of 8 warps
        ridge_gram_small_kernel<<<a.splits, 256, 0, s>>>(a, T);
        ridge_gram_small_reduce_kernel<<<a.K * a.K, 128, 0, s>>>(a.partial, a.splits * 8, a.K, G);
        return cudaGetLastError();
    }
    const int npairs = a.nb * (a.nb + 1) / 2;
    ridge_gram_kernel<<<dim3(npairs, a.splits), 256, 0, s>>>(a, T);
    const long long tot = (long long)a.K * a.K;
    ridge_gram_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(a.partial, a.nb, npairs, a.splits, a.K, G);
    return cudaGetLastError();
}

cudaError_t launch_ridge_predict(const RidgeArgs& a, const double* T, const double* coeff, int Kfeat, float* base, double* err_partial,
                                 int grid, cudaStream_t s) {
    ridge_predict_kernel<<<grid, 256, 0, s>>>(a, T, coeff, Kfeat, base, err_partial);
    return cudaGetLastError();
}

}  // namespace mjb

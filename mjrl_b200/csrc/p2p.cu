// Fused "sum of the per-CTA gradient partials" + all-reduce over NVLink peer memory (every Fisher-vector product
// ends in an all-reduce of d floats -- 20 KB .. 330 KB, latency-bound -- ten times per train step).
//
// One kernel does both steps.  Thread i (128 per CTA) reduces element i of this rank's per-CTA partials exactly as
// reduce_partials_kernel does, then STORES it into slot [rank] of every rank's exchange buffer (peer-mapped pointers, opened
// with cudaIpcOpenMemHandle at communicator setup: plain NVLink stores, no copy engine, no proxy thread) as ONE 64-bit word
// {value, call number}.  The call number travels with the value, so there are no flags and no fences: the same thread
// polls element i of every rank's slot in its OWN memory until the word carries this call's number and adds the values in
// rank order -- every rank adds the same numbers in the same order, so the replicated CG states stay bit-identical.  (A
// first version with separate flags needed a system-scope release fence after the scatter and an acquire fence after the
// poll: 37 us per all-reduce at 2 GPUs against ncclAllReduce's 29 us.)  No grid-wide or CTA-wide synchronisation: thread i
// only ever depends on thread i of the peers.
//
// Reuse of the slots across calls: double-buffered by the parity of a per-CTA call counter kept in device memory (kernel
// arguments are frozen inside the captured CG graph).  A rank can start call k+2 only after call k+1 has completed, for
// which every peer must have sent its k+1 values -- which a peer does only after its kernel of call k has finished
// reading: two parities are enough.
#include <cuda_runtime.h>

#include "kernels.h"

namespace mjb {

namespace {

__device__ __forceinline__ void st_relaxed_sys64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.b64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.b64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(128) reduce_allreduce_p2p_kernel(const P2PReduceArgs a) {
    const int tid = threadIdx.x, cta = blockIdx.x;
    const int i = cta * 128 + tid;
    const unsigned seq = (unsigned)a.cta_seq[cta] + 1u;              // this CTA's call number (same on every rank, never 0)
    const unsigned par = seq & 1u;
    if (i < a.d) {
        // ---- local reduction (identical to reduce_partials_kernel) ----
        const double sc0 = a.vscale ? a.scale_dev[0] * (double)a.vscale[1] : a.scale_dev[0];
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int c = 0;
        for (; c + 3 < a.grid; c += 4) {
            s0 += a.partial[(size_t)c * a.stride + i];
            s1 += a.partial[(size_t)(c + 1) * a.stride + i];
            s2 += a.partial[(size_t)(c + 2) * a.stride + i];
            s3 += a.partial[(size_t)(c + 3) * a.stride + i];
        }
        for (; c < a.grid; ++c) s0 += a.partial[(size_t)c * a.stride + i];
        float r = (float)((double)((s0 + s1) + (s2 + s3)) * sc0);
        if (a.fvp_ls_block && i >= a.tLS) {
            const float u = expf(2.0f * a.theta[i]);
            const float den = 2.0f * u + 1e-8f;
            r = a.v[i] * ((8.0f * u * u - 4.0f * u * 1e-8f) / (den * den));
            r = (float)((double)r * a.scale_dev[1]);                 // every rank adds the block once: 1/world each
        }
        // ---- scatter {value, call number} into slot [par][rank] of every peer ----
        const unsigned long long word = ((unsigned long long)seq << 32) | (unsigned long long)__float_as_uint(r);
        const size_t mine = ((size_t)par * a.world + a.rank) * a.slot_words + (size_t)i;
        for (int p = 0; p < a.world; ++p)
            if (p != a.rank) st_relaxed_sys64(a.peers[p] + mine, word);
        // ---- gather in rank order: poll until the word of this call has landed ----
        const unsigned long long* own = a.peers[a.rank] + (size_t)par * a.world * a.slot_words + (size_t)i;
        float s = 0.0f;
        const long long t0 = clock64();
        for (int q = 0; q < a.world; ++q) {
            if (q == a.rank) { s += r; continue; }
            unsigned long long w;
            while ((unsigned)((w = ld_relaxed_sys64(own + (size_t)q * a.slot_words)) >> 32) != seq) {
                if (clock64() - t0 > 8000000000ll) __trap();        // a peer that never arrives: fail the launch, do not hang
            }
            s += __uint_as_float((unsigned)w);
        }
        a.out[i] = s;
    }
    __syncthreads();
    if (tid == 0) a.cta_seq[cta] = (int)seq;
}

}  // namespace

cudaError_t launch_reduce_allreduce_p2p(const P2PReduceArgs& a, cudaStream_t s) {
    reduce_allreduce_p2p_kernel<<<(a.d + 127) / 128, 128, 0, s>>>(a);
    return cudaGetLastError();
}

}  // namespace mjb

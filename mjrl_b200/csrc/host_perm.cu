// Host-side helper of the baseline fit: np.random.permutation(n) of the GLOBAL legacy RandomState, bit for bit.
//
// MLPBaseline.fit draws one permutation of the batch per epoch from numpy's global RNG (utils/optimize_model.py:22).
// At 1e6 timesteps that draw costs 18 ms of host time on the critical path of a 150 ms step (the fit kernel cannot
// start before it has its minibatch order).  numpy's implementation interleaves MT19937 generation, rejection
// sampling and the Fisher-Yates swap element by element; here the same three steps run as tight batched loops:
//   1. j_i = random_interval(i) for i = n-1 .. 1   (legacy-distributions: mask-and-reject on 32-bit MT19937 draws)
//   2. swap(x[i], x[j_i]) in the same order
// The caller passes numpy's MT19937 state (624 words + position, np.random.get_state()) and writes the advanced state
// back with np.random.set_state(), so every later draw of the program is unchanged as well.
#include <stdint.h>
#include <stdlib.h>

namespace {

constexpr int MT_N = 624, MT_M = 397;
constexpr uint32_t MATRIX_A = 0x9908b0dfu, UPPER_MASK = 0x80000000u, LOWER_MASK = 0x7fffffffu;

struct Mt { uint32_t* key; int pos; };

inline void mt_gen(Mt& s) {
    uint32_t* key = s.key;
    int kk = 0;
    uint32_t y;
    for (; kk < MT_N - MT_M; ++kk) {
        y = (key[kk] & UPPER_MASK) | (key[kk + 1] & LOWER_MASK);
        key[kk] = key[kk + MT_M] ^ (y >> 1) ^ ((0u - (y & 1u)) & MATRIX_A);
    }
    for (; kk < MT_N - 1; ++kk) {
        y = (key[kk] & UPPER_MASK) | (key[kk + 1] & LOWER_MASK);
        key[kk] = key[kk + (MT_M - MT_N)] ^ (y >> 1) ^ ((0u - (y & 1u)) & MATRIX_A);
    }
    y = (key[MT_N - 1] & UPPER_MASK) | (key[0] & LOWER_MASK);
    key[MT_N - 1] = key[MT_M - 1] ^ (y >> 1) ^ ((0u - (y & 1u)) & MATRIX_A);
    s.pos = 0;
}

inline uint32_t temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

}  // namespace

extern "C" int mjb_host_permutation(uint32_t* mt_key624, int32_t* mt_pos, int64_t n, int32_t* out) {
    if (!mt_key624 || !mt_pos || !out || n < 0 || n > 0x7fffffffLL || *mt_pos < 0 || *mt_pos > MT_N) return -1;
    for (int64_t i = 0; i < n; ++i) out[i] = (int32_t)i;
    if (n < 2) return 0;
    Mt s{mt_key624, *mt_pos};
    uint32_t* js = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(n + 1));
    if (!js) return -2;
    // 1. the indices (numpy: random_interval(bitgen, i) for i = n-1 .. 1; max <= 0xffffffff -> one 32-bit MT19937 draw
    //    per trial, masked to the smallest 2^k - 1 >= i, rejected while > i).  Every trial consumes exactly one draw, so
    //    the draws are tempered a block at a time and compacted without branches: store, then advance only if accepted.
    uint32_t buf[MT_N];
    int have = 0, used = 0;                              // tempered outputs of the current block: buf[used .. have)
    int block_pos = s.pos;                               // position inside the current key block (numpy's `pos`)
    uint32_t mask = (uint32_t)(n - 1);
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    int64_t i = n - 1;
    while (i >= 1) {
        if (used == have) {                              // refill: the rest of the current block, or a fresh block
            if (block_pos == MT_N) { mt_gen(s); block_pos = 0; }
            have = MT_N - block_pos;
            for (int k = 0; k < have; ++k) buf[k] = temper(s.key[block_pos + k]);
            used = 0;
        }
        const int64_t lo = (int64_t)(mask >> 1);         // this mask serves i in (lo, mask]
        int k = used;
        while (k < have && i > lo) {
            const uint32_t v = buf[k++] & mask;
            js[i] = v;                                   // overwritten by the next trial if rejected
            i -= (v <= (uint32_t)i);
        }
        block_pos += k - used;
        used = k;
        if (i <= lo) mask >>= 1;
    }
    s.pos = block_pos;
    // 2. the swaps
    for (int64_t i = n - 1; i >= 1; --i) {
        if (i > 48) __builtin_prefetch(&out[js[i - 48]], 1, 1);    // the partner of a later swap: random access in a 4 MB array
        const uint32_t j = js[i];
        const int32_t t = out[i];
        out[i] = out[j];
        out[j] = t;
    }
    free(js);
    *mt_pos = s.pos;
    return 0;
}

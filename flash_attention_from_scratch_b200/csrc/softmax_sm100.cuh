// Per-row softmax building blocks of the attention kernel (one thread owns one row of a 128-wide
// S block held in registers as 4 fragments of 32 fp32 values).  Shared by the production kernel
// (fa_fwd_sm100.cuh) and the timing probe (tools/softmax_probe.cu) so that tuning experiments
// measure exactly the code that ships.
//
// Arithmetic follows the reference (/root/reference/src/include/softmax.cuh:15-105):
// m = max(m, rowmax S); P = exp2(S*c - m*c); l += rowsum(P) in fp32; P rounded RN to 16 bit.
#pragma once
#include <cstdint>

#include "ptx_sm100.cuh"

namespace fa {

// evenly spread `n` polynomial-exp2 pairs over the 16 (p0,p1) pairs of a 32-column fragment
__host__ __device__ constexpr bool emulate_pair(int pair, int n) {
    return n > 0 && ((pair * n) % 16) < n;
}

// Row max of 128 values: 8 independent chains (a serial chain would expose 43 x FMNMX latency).
__device__ __forceinline__ float row_max_128(const uint32_t (&sr)[4][32]) {
    float mxs[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) mxs[u] = __uint_as_float(sr[u >> 1][(u & 1) * 16]);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int i = 1; i < 16; ++i) {
            mxs[2 * q] = fmaxf(mxs[2 * q], __uint_as_float(sr[q][i]));
            mxs[2 * q + 1] = fmaxf(mxs[2 * q + 1], __uint_as_float(sr[q][16 + i]));
        }
    }
    return fmaxf(fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3])),
                 fmaxf(fmaxf(mxs[4], mxs[5]), fmaxf(mxs[6], mxs[7])));
}

// Row max of fragments [kFirst, kFirst + kCount) of a row (two chains per fragment).
template <int kFirst, int kCount>
__device__ __forceinline__ float row_max_frags(const uint32_t (&sr)[4][32]) {
    float mxs[2 * kCount];
#pragma unroll
    for (int u = 0; u < 2 * kCount; ++u) mxs[u] = __uint_as_float(sr[kFirst + (u >> 1)][(u & 1) * 16]);
#pragma unroll
    for (int q = 0; q < kCount; ++q) {
#pragma unroll
        for (int i = 1; i < 16; ++i) {
            mxs[2 * q] = fmaxf(mxs[2 * q], __uint_as_float(sr[kFirst + q][i]));
            mxs[2 * q + 1] = fmaxf(mxs[2 * q + 1], __uint_as_float(sr[kFirst + q][16 + i]));
        }
    }
    float m = mxs[0];
#pragma unroll
    for (int u = 1; u < 2 * kCount; ++u) m = fmaxf(m, mxs[u]);
    return m;
}

// One 32-column fragment: x = S*c - m*c (FFMA2), p = 2^x (MUFU or FMA-pipe polynomial for `kEmu`
// of the 16 pairs), two fp32 partial row sums (FADD2), 16 packed 16-bit pairs for tcgen05.st.
//   kVariant 0: one fused loop per pair (the compiler interleaves freely)
//   kVariant 1: three explicit stages (all FFMA2, then all exp2, then sum + pack)
template <bool kBF16, int kEmu, int kVariant>
__device__ __forceinline__ void exp_fragment(const uint32_t (&sr)[32], float2 c2, float2 nm2,
                                             float2& sum_a, float2& sum_b, uint32_t (&pk)[16]) {
    if constexpr (kVariant == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float2 x = __ffma2_rn(
                make_float2(__uint_as_float(sr[2 * i]), __uint_as_float(sr[2 * i + 1])), c2, nm2);
            float2 p;
            if (emulate_pair(i, kEmu)) {
                p = ex2_emulated_x2(x);
            } else {
                p.x = ex2_approx(x.x);
                p.y = ex2_approx(x.y);
            }
            if (i & 1) sum_a = __fadd2_rn(sum_a, p);
            else sum_b = __fadd2_rn(sum_b, p);
            pk[i] = pack_16x2<kBF16>(p.x, p.y);
        }
    } else {
        float2 x[16];
#pragma unroll
        for (int i = 0; i < 16; ++i)
            x[i] = __ffma2_rn(
                make_float2(__uint_as_float(sr[2 * i]), __uint_as_float(sr[2 * i + 1])), c2, nm2);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (emulate_pair(i, kEmu)) {
                x[i] = ex2_emulated_x2(x[i]);
            } else {
                x[i].x = ex2_approx(x[i].x);
                x[i].y = ex2_approx(x[i].y);
            }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (i & 1) sum_a = __fadd2_rn(sum_a, x[i]);
            else sum_b = __fadd2_rn(sum_b, x[i]);
            pk[i] = pack_16x2<kBF16>(x[i].x, x[i].y);
        }
    }
}

}  // namespace fa

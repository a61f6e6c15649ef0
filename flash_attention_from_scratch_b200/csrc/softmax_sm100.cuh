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

// Like exp_fragment, but additionally tracks the max of the inputs of the polynomial pairs.
template <bool kBF16, int kEmu>
__device__ __forceinline__ void exp_fragment_track(const uint32_t (&sr)[32], float2 c2, float2 nm2,
                                                   float2& sum_a, float2& sum_b, float& xmax,
                                                   uint32_t (&pk)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float2 x = __ffma2_rn(
            make_float2(__uint_as_float(sr[2 * i]), __uint_as_float(sr[2 * i + 1])), c2, nm2);
        float2 p;
        if (emulate_pair(i, kEmu)) {
            p = ex2_emulated_x2_track(x, xmax);
        } else {
            p.x = ex2_approx(x.x);
            p.y = ex2_approx(x.y);
        }
        if (i & 1) sum_a = __fadd2_rn(sum_a, p);
        else sum_b = __fadd2_rn(sum_b, p);
        pk[i] = pack_16x2<kBF16>(p.x, p.y);
    }
}

// Max of one 32-value fragment (2 chains).
__device__ __forceinline__ float frag_max_32(const uint32_t (&sr)[32]) {
    float a = __uint_as_float(sr[0]), b = __uint_as_float(sr[16]);
#pragma unroll
    for (int i = 1; i < 16; ++i) {
        a = fmaxf(a, __uint_as_float(sr[i]));
        b = fmaxf(b, __uint_as_float(sr[16 + i]));
    }
    return fmaxf(a, b);
}

// One KV block (not the first of a tile) of online softmax WITHOUT computing the block's row max
// up front.  The lazy-rescale invariant only needs P <= 2^threshold against the running (stale) max
// m_run, and that can be verified after the fact at almost no cost:
//   * MUFU elements of fragments 0-2: if one exceeded 2^threshold the fp32 row sum does too
//     (inf/NaN included: the test is written so that NaN fails it),
//   * polynomial elements of fragments 0-2 (the emulation is only valid for inputs <= 127): a running
//     max of their inputs (one FMNMX3 per pair),
//   * fragment 3, which is published separately after the first three: its 32-value max, checked
//     before anything is published.
// If every row of the warp passes, P is published as computed -- the serialised row-max phase
// (~270-400 clk per block, profiles/r01_v4_trace_notes.md) is gone.  Otherwise (rare: first blocks,
// adversarial data) the exact path runs: full row max, failing rows adopt it, O and l are rescaled,
// fragments 0-2 are recomputed.  Both paths give online softmax with a stale max, exact up to
// rounding, like the lazy rescale of the plain path.
//   store_p(q, pk), arrive_part(last), rescale_o(alpha): as in the kernel.
template <bool kBF16, int kEmu, int kEmuLast, class StoreP, class ArrivePart, class RescaleO>
__device__ __forceinline__ void softmax_block_nomax(const uint32_t (&sr)[4][32], float c,
                                                    float threshold, float& m_run, float& l_run,
                                                    StoreP&& store_p, ArrivePart&& arrive_part,
                                                    RescaleO&& rescale_o) {
    const float2 c2 = make_float2(c, c);
    float2 sum_a = make_float2(0.f, 0.f), sum_b = make_float2(0.f, 0.f);
    float alpha = 1.f;
    float xmax = -INFINITY;
    {
        const float neg_mc = -m_run * c;
        const float2 nm2 = make_float2(neg_mc, neg_mc);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            uint32_t pk[16];
            exp_fragment_track<kBF16, kEmu>(sr[q], c2, nm2, sum_a, sum_b, xmax, pk);
            store_p(q, pk);
        }
    }
    const float limit = exp2f(threshold);
    const float s012 = (sum_a.x + sum_a.y) + (sum_b.x + sum_b.y);
    const float m3 = frag_max_32(sr[3]);
    // written so that NaN (from inf - inf, garbage polynomial output, NaN inputs) fails the test
    const bool ok = (s012 <= limit) && (xmax <= threshold) && ((m3 - m_run) * c <= threshold);
    if (!__all_sync(0xffffffffu, ok)) {
        // exact path for this warp: failing rows adopt the true block max
        const float mx = fmaxf(row_max_128(sr), m_run);
        if (!ok) {
            alpha = ex2_approx((m_run - mx) * c);
            m_run = mx;
        }
        rescale_o(alpha);
        sum_a = make_float2(0.f, 0.f);
        sum_b = make_float2(0.f, 0.f);
        const float neg_mc = -m_run * c;
        const float2 nm2 = make_float2(neg_mc, neg_mc);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            uint32_t pk[16];
            exp_fragment<kBF16, kEmu, 0>(sr[q], c2, nm2, sum_a, sum_b, pk);
            store_p(q, pk);
        }
    }
    arrive_part(false);
    {
        const float neg_mc = -m_run * c;
        const float2 nm2 = make_float2(neg_mc, neg_mc);
        uint32_t pk[16];
        exp_fragment<kBF16, kEmuLast, 0>(sr[3], c2, nm2, sum_a, sum_b, pk);
        store_p(3, pk);
    }
    arrive_part(true);
    l_run = l_run * alpha + ((sum_a.x + sum_a.y) + (sum_b.x + sum_b.y));
}

}  // namespace fa

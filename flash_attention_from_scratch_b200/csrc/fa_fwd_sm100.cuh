// Fused attention forward for NVIDIA B200 (sm_100a):  O = softmax(Q K^T / sqrt(d)) V
//
// Replaces the reference's Ampere kernel `flash::flash_forward_kernel`
// (/root/reference/src/include/forward_kernel.cuh:85-204) and its device primitives
// (gemm.cuh / softmax.cuh / load_store.cuh).  Same arithmetic contract
// (/root/reference/src/include/softmax.cuh:15-128, forward_kernel.cuh:150-152):
//     c = log2(e)/sqrt(d);  m = running row max of raw S;  P = exp2(S*c - m*c) in fp32;
//     l += rowsum(P) (un-rounded fp32);  O += rn16(P) V (fp32 accumulate);  out = rn16(O / l)
// but a different machine mapping -- nothing of the mma.sync / ldmatrix / cp.async path is kept.
//
// Kernel generation 3 (see profiles/r01_v2_ncu_notes.md for why):
//   * one CTA = one Q tile of 128 rows of one (batch, head); 1 CTA per SM, 384 threads
//   * TMA (cp.async.bulk.tensor, 128B swizzle) stages Q once and K/V blocks of 128 rows through a
//     ring of 6 shared-memory slots guarded by full/empty mbarriers
//   * one thread issues tcgen05.mma.  S is DOUBLE-BUFFERED in tensor memory across consecutive KV
//     blocks, so S(j+2) = Q K_{j+2}^T is computed while the softmax of block j+1 runs and the
//     softmax never waits for the tensor pipe (generation 2 aliased P onto the only S buffer of a
//     tile, which made the loop softmax -> PV -> next S latency bound at 62 % tensor activity):
//       TMEM columns [0,128) S[0], [128,256) S[1], [256,384) O_0, [384,512) O_1
//   * the two softmax warpgroups split every KV block by KEYS: warpgroup g owns key columns
//     [64g, 64g+64) of each block, with its own running max m_g, row sum l_g and accumulator O_g
//     (O_g += P_g V[64g:64g+64]); no cross-warpgroup exchange inside the loop.  One thread per row:
//     tcgen05.ld S, fp32 row max, exp2 (MUFU + a tunable share on the FMA pipe), fp32 row sum,
//     P written back as packed 16-bit over the first 32 columns of its own S half, lazy rescale
//     of O_g (only when the max grew by more than 2^8).
//   * epilogue: the two partial results are merged like split-KV attention,
//       m = max(m_0, m_1), a_g = 2^((m_g - m) c), out = (a_0 O_0 + a_1 O_1) / (a_0 l_0 + a_1 l_1),
//     rounded to 16 bit, staged in swizzled shared memory and written with TMA.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#include "ptx_sm100.cuh"

namespace fa {

constexpr int kBlockM = 128;   // query rows per CTA == TMEM lanes
constexpr int kBlockN = 128;   // key/value rows per block
constexpr int kHeadDim = 128;  // d_head (the only one the reference supports, README.md:9-15)
constexpr int kKeySplit = 2;   // softmax warpgroups; each owns kBlockN / kKeySplit keys of a block
constexpr int kKeysPerWG = kBlockN / kKeySplit;
constexpr int kKVStages = 6;   // K/V ring slots (each slot holds one K block or one V block)
constexpr int kTileBytes = kBlockN * kHeadDim * 2;  // 32 KiB: one 128x128 16-bit tile
constexpr int kHalfBytes = kTileBytes / 2;          // one TMA box: 128 rows x 64 cols (128 B rows)
constexpr int kNumThreads = 384;                    // 2 softmax warpgroups + 1 control warpgroup
constexpr int kTmemCols = 512;

constexpr int kSmemQ = 0;
constexpr int kSmemKV = kSmemQ + kTileBytes;
constexpr int kSmemBar = kSmemKV + kKVStages * kTileBytes;
constexpr int kNumBarriers = 1 + 2 * kKVStages + 2 + 4 + 2;
constexpr int kSmemTmemPtr = kSmemBar + kNumBarriers * 8;
constexpr int kSmemTotal = kSmemTmemPtr + 16;
constexpr int kSmemLaunchBytes = kSmemTotal + 1024;  // slack for manual 1024 B alignment
static_assert(kSmemLaunchBytes <= 232448, "exceeds the 227 KiB opt-in shared memory of sm_100");

// Tunables (overridable with -D at build time; tools/build_variants.py sweeps them).
#ifndef FA_EMU_PAIRS
#define FA_EMU_PAIRS 4  // of every 16 (p0,p1) pairs, how many use the FMA-pipe exp2 (rest: MUFU)
#endif
constexpr int kEmuPairs = FA_EMU_PAIRS;
// evenly spread `n` emulated pairs over the 16 pairs of a 32-column fragment
__host__ __device__ constexpr bool emulate_pair(int pair, int n) {
    return n > 0 && ((pair * n) % 16) < n;
}

// Lazy rescale threshold in log2 units: O_g and l_g are only rescaled when the running max grows
// by more than this; until then P is computed against the stale max, i.e. P <= 2^8 (exact in
// fp32, representable in bf16/fp16).  The final normalisation is unaffected.
constexpr float kRescaleThreshold = 8.0f;

// Bring-up hooks (only read by the kDebug instantiation; see tools/gpu_bringup.py).
struct FwdDebug {
    float* dump;      // level 2: raw smem Q | K0 (2 x 8192 words); level >= 3: S(block 0) [128][128],
                      // then m_g [2][128], l_g [2][128] of CTA 0
    uint32_t level;   // 1: setup/teardown only, 2: + TMA Q,K0, 3: + S = QK^T, >= 4: everything
    uint32_t* diag;   // host-mapped diagnostics ring (hang-guard builds)
};

struct FwdParams {
    int batch;
    int seq_len;
    int n_heads;
    int n_kv_blocks;   // seq_len / 128
    int n_q_tiles;     // seq_len / 128: CTAs per (batch, head)
    float scale_log2;  // log2(e) / sqrt(d_head)
};

template <bool kBF16, bool kDebug>
__global__ void __launch_bounds__(kNumThreads, 1)
fa_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
              const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o,
              const FwdParams prm, const FwdDebug dbg) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int wg = warp >> 2;

    // barrier addresses
    const uint32_t bar0 = smem_base + kSmemBar;
    const uint32_t q_full = bar0;
    auto kv_full = [&](int i) { return bar0 + 8u * (1 + i); };
    auto kv_empty = [&](int i) { return bar0 + 8u * (1 + kKVStages + i); };
    auto s_full = [&](int x) { return bar0 + 8u * (1 + 2 * kKVStages + x); };
    auto p_full = [&](int x, int g) { return bar0 + 8u * (3 + 2 * kKVStages + 2 * x + g); };
    auto pv_done = [&](int g) { return bar0 + 8u * (7 + 2 * kKVStages + g); };
    const uint32_t tmem_ptr_smem = smem_base + kSmemTmemPtr;

    // tile coordinates: q tile fastest so co-resident CTAs share K/V of one (b, h) in L2
    const int tile = blockIdx.x;
    const int qtile = tile % prm.n_q_tiles;
    const int bh = tile / prm.n_q_tiles;
    const int head = bh % prm.n_heads;
    const int batch = bh / prm.n_heads;
    const int n_blocks = prm.n_kv_blocks;

    const uint32_t level = kDebug ? dbg.level : 4u;
#if FA_HANG_GUARD
    if constexpr (kDebug) {
        if (threadIdx.x == 0) g_fa_diag = dbg.diag;
    }
#endif
    if (warp == 8) {
        if (lane == 0) {
            mbar_init(q_full, 1);
            for (int i = 0; i < kKVStages; ++i) {
                mbar_init(kv_full(i), 1);
                mbar_init(kv_empty(i), 1);
            }
            for (int x = 0; x < 2; ++x) {
                mbar_init(s_full(x), 1);
                mbar_init(pv_done(x), 1);
                mbar_init(p_full(x, 0), 4);  // one elected arrive per softmax warp
                mbar_init(p_full(x, 1), 4);
            }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(tmem_ptr_smem, kTmemCols);
        tmem_relinquish();
    } else if (warp == 9 && lane == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_k);
        tma_prefetch_desc(&tm_v);
        tma_prefetch_desc(&tm_o);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + kSmemTmemPtr);
    const uint32_t tmem_s0 = tmem_base;                  // S[x] at + x * 128
    const uint32_t tmem_o0 = tmem_base + 2 * kBlockN;    // O_g at + g * 128

    if (wg == 2) {
        if (warp == 9) {
            // ================================ TMA producer ================================
            if (lane == 0) {
                auto load_tile = [&](const CUtensorMap* map, uint32_t dst, uint32_t bar, int row0) {
                    mbar_arrive_expect_tx(bar, kTileBytes);
                    tma_load_4d(dst, map, bar, 0, head, row0, batch);
                    tma_load_4d(dst + kHalfBytes, map, bar, 64, head, row0, batch);
                };
                int item = 0;  // ring order: K0, K1, V0, K2, V1, K3, ... (the MMA warp's order)
                auto load_kv = [&](const CUtensorMap* map, int blk) {
                    const int slot = item % kKVStages;
                    const uint32_t use = item / kKVStages;
                    mbar_wait(kv_empty(slot), (use & 1u) ^ 1u, 100 + slot);
                    load_tile(map, smem_base + kSmemKV + slot * kTileBytes, kv_full(slot),
                              blk * kBlockN);
                    ++item;
                };
                if (level >= 2) {
                    load_tile(&tm_q, smem_base + kSmemQ, q_full, qtile * kBlockM);
                    load_kv(&tm_k, 0);
                }
                if (level >= 4) {
                    if (n_blocks > 1) load_kv(&tm_k, 1);
                    for (int j = 0; j < n_blocks; ++j) {
                        load_kv(&tm_v, j);
                        if (j + 2 < n_blocks) load_kv(&tm_k, j + 2);
                    }
                }
            }
        } else if (warp == 8) {
            // ================================ MMA issuer ==================================
            if (lane == 0) {
                constexpr uint32_t idesc_qk = umma_idesc_f16(kBF16, kBlockM, kBlockN, false);
                constexpr uint32_t idesc_pv = umma_idesc_f16(kBF16, kBlockM, kHeadDim, true);
                // Q/K tiles: K-major, 8-row x 128 B swizzle atoms 1024 B apart (SBO); LBO unused.
                // V tiles: MN-major; next 64-wide d chunk 16 KiB away (LBO), next 8 kv rows 1 KiB
                // (SBO); one k-step = 16 kv rows = 2 KiB.
                const uint64_t q_desc = umma_smem_desc_sw128(smem_base + kSmemQ, 16, 1024);
                auto issue_qk = [&](int x, int slot) {
                    const uint64_t b0 =
                        umma_smem_desc_sw128(smem_base + kSmemKV + slot * kTileBytes, 16, 1024);
#pragma unroll
                    for (int k = 0; k < kHeadDim / 16; ++k) {
                        const uint32_t off = ((k >> 2) * kHalfBytes + (k & 3) * 32) >> 4;
                        umma_ss(tmem_s0 + x * kBlockN, q_desc + off, b0 + off, idesc_qk, k > 0);
                    }
                };
                auto issue_pv = [&](int g, int x, int slot, bool accumulate) {
                    const uint64_t b0 = umma_smem_desc_sw128(
                        smem_base + kSmemKV + slot * kTileBytes + g * (kKeysPerWG / 16) * 2048,
                        kHalfBytes, 1024);
#pragma unroll
                    for (int k = 0; k < kKeysPerWG / 16; ++k) {
                        umma_ts(tmem_o0 + g * kHeadDim,
                                tmem_s0 + x * kBlockN + g * kKeysPerWG + k * 8,
                                b0 + ((k * 2048) >> 4), idesc_pv, (accumulate || k > 0) ? 1u : 0u);
                    }
                };
                int item = 0;
                auto slot_of = [&](int it) { return it % kKVStages; };
                auto wait_item = [&](int it, int tag) {
                    mbar_wait(kv_full(slot_of(it)), (uint32_t)((it / kKVStages) & 1), tag);
                    tc_fence_after();
                };
                if constexpr (kDebug) {
                    if (level == 2) {  // raw smem images of Q and K0 as TMA wrote them
                        mbar_wait(kv_full(0), 0, 200);
                        mbar_wait(q_full, 0, 210);
                        if (dbg.dump != nullptr && blockIdx.x == 0) {
                            const uint32_t* qs = reinterpret_cast<const uint32_t*>(smem_gen + kSmemQ);
                            const uint32_t* ks = reinterpret_cast<const uint32_t*>(smem_gen + kSmemKV);
                            uint32_t* out = reinterpret_cast<uint32_t*>(dbg.dump);
                            for (int i = 0; i < kTileBytes / 4; ++i) out[i] = qs[i];
                            for (int i = 0; i < kTileBytes / 4; ++i) out[kTileBytes / 4 + i] = ks[i];
                        }
                    }
                }
                if (level >= 3) {
                    // prologue: S[0] = Q K_0^T, S[1] = Q K_1^T
                    mbar_wait(q_full, 0, 210);
                    const int n_pro = (level >= 4 && n_blocks > 1) ? 2 : 1;
                    for (int x = 0; x < n_pro; ++x) {
                        wait_item(item, 200 + x);
                        issue_qk(x, slot_of(item));
                        umma_commit(s_full(x));
                        umma_commit(kv_empty(slot_of(item)));
                        ++item;
                    }
                }
                for (int j = 0; level >= 4 && j < n_blocks; ++j) {
                    const int x = j & 1;
                    const uint32_t par = (uint32_t)((j >> 1) & 1);
                    const int it_v = item++;
                    wait_item(it_v, 220);
#pragma unroll
                    for (int g = 0; g < kKeySplit; ++g) {
                        mbar_wait(p_full(x, g), par, 230 + g);  // P_g(j) stored, O_g rescaled
                        tc_fence_after();
                        issue_pv(g, x, slot_of(it_v), j > 0);
                        umma_commit(pv_done(g));
                    }
                    umma_commit(kv_empty(slot_of(it_v)));
                    if (j + 2 < n_blocks) {
                        // both P halves of S[x] are consumed (in-order tensor pipe): refill S[x]
                        const int it_k = item++;
                        wait_item(it_k, 240);
                        issue_qk(x, slot_of(it_k));
                        umma_commit(s_full(x));
                        umma_commit(kv_empty(slot_of(it_k)));
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ==================================== softmax =====================================
        const int g = wg;                    // key half owned by this warpgroup
        const int row = threadIdx.x & 127;   // query row == TMEM lane
        const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const uint32_t t_s = tmem_s0 + lane_sel + g * kKeysPerWG;  // + x * 128
        const uint32_t t_o = tmem_o0 + lane_sel + g * kHeadDim;    // own accumulator O_g
        const float c = prm.scale_log2;

        float m_run = -INFINITY;  // running (possibly stale) max over this warpgroup's keys
        float l_run = 0.f;        // running row sum of exp2 over this warpgroup's keys

        const int n_iter = (level >= 4) ? n_blocks : (level == 3 ? 1 : 0);
        for (int j = 0; j < n_iter; ++j) {
            const int x = j & 1;
            mbar_wait(s_full(x), (uint32_t)((j >> 1) & 1), 300 + g);
            tc_fence_after();
            uint32_t sr[2][32];
            tmem_ld_32x32b_x32(t_s + x * kBlockN, sr[0]);
            tmem_ld_32x32b_x32(t_s + x * kBlockN + 32, sr[1]);
            tmem_wait_ld();

            if constexpr (kDebug) {
                if (dbg.dump != nullptr && blockIdx.x == 0 && j == 0) {
                    for (int q = 0; q < 2; ++q)
                        for (int i = 0; i < 32; ++i)
                            dbg.dump[row * 128 + g * kKeysPerWG + q * 32 + i] =
                                __uint_as_float(sr[q][i]);
                }
                if (level == 3) break;
            }
            // row max: 4 independent chains
            float mxs[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) mxs[u] = __uint_as_float(sr[u >> 1][(u & 1) * 16]);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
#pragma unroll
                for (int i = 1; i < 16; ++i) {
                    mxs[2 * q] = fmaxf(mxs[2 * q], __uint_as_float(sr[q][i]));
                    mxs[2 * q + 1] = fmaxf(mxs[2 * q + 1], __uint_as_float(sr[q][16 + i]));
                }
            }
            float mx = fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3]));
            mx = fmaxf(mx, m_run);
            float alpha = 1.f;
            if (j == 0) {
                m_run = mx;
            } else {
                const float delta = (mx - m_run) * c;  // >= 0
                const bool need = delta > kRescaleThreshold;
                if (__any_sync(0xffffffffu, need)) {
                    if (need) {
                        alpha = ex2_approx(-delta);
                        m_run = mx;
                    }
                    // O_g must be quiescent: wait until PV_g(j-1) retired.
                    mbar_wait(pv_done(g), (uint32_t)((j - 1) & 1), 320 + g);
                    tc_fence_after();
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t o[32];
                        tmem_ld_32x32b_x32(t_o + q * 32, o);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st_32x32b_x32(t_o + q * 32, o);
                    }
                }
            }
            const float neg_mc = -m_run * c;
            const float2 c2 = make_float2(c, c);
            const float2 nm2 = make_float2(neg_mc, neg_mc);
            float2 sum_a = make_float2(0.f, 0.f), sum_b = make_float2(0.f, 0.f);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float2 xv = __ffma2_rn(
                        make_float2(__uint_as_float(sr[q][2 * i]), __uint_as_float(sr[q][2 * i + 1])),
                        c2, nm2);
                    float2 p;
                    if (emulate_pair(i, kEmuPairs)) {
                        p = ex2_emulated_x2(xv);
                    } else {
                        p.x = ex2_approx(xv.x);
                        p.y = ex2_approx(xv.y);
                    }
                    if (i & 1) sum_a = __fadd2_rn(sum_a, p);
                    else sum_b = __fadd2_rn(sum_b, p);
                    pk[i] = pack_16x2<kBF16>(p.x, p.y);
                }
                // P_g(j): 16-bit pairs over the first 32 columns of this warpgroup's S half
                tmem_st_32x32b_x16(t_s + x * kBlockN + q * 16, pk);
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full(x, g));
            l_run = l_run * alpha + ((sum_a.x + sum_a.y) + (sum_b.x + sum_b.y));
        }

        // --------------------------------- epilogue --------------------------------------
        if (level >= 4) {
            // every MMA retired => both accumulators final, all smem tiles dead
            mbar_wait(pv_done(0), (uint32_t)((n_blocks - 1) & 1), 310);
            mbar_wait(pv_done(1), (uint32_t)((n_blocks - 1) & 1), 311);
            tc_fence_after();
            // exchange (m_g, l_g) through the (dead) first K/V slot
            float2* stats = reinterpret_cast<float2*>(smem_gen + kSmemKV);
            stats[g * kBlockM + row] = make_float2(m_run, l_run);
            if constexpr (kDebug) {
                if (dbg.dump != nullptr && blockIdx.x == 0) {
                    dbg.dump[128 * 128 + g * 128 + row] = m_run;
                    dbg.dump[128 * 128 + 256 + g * 128 + row] = l_run;
                }
            }
            named_bar_sync(1, 2 * 128);
            const float2 st0 = stats[row];
            const float2 st1 = stats[kBlockM + row];
            const float m_all = fmaxf(st0.x, st1.x);
            const float a0 = ex2_approx((st0.x - m_all) * c);
            const float a1 = ex2_approx((st1.x - m_all) * c);
            const float inv = 1.0f / (a0 * st0.y + a1 * st1.y);
            const float w0 = a0 * inv, w1 = a1 * inv;
            // warpgroup g produces d columns [64g, 64g+64) == TMA box g, staged in the (dead) Q
            // tile with the 128B swizzle: 16-byte chunk c of row r lives at chunk (c ^ (r & 7)).
            uint8_t* stage = smem_gen + kSmemQ + g * kHalfBytes + row * 128;
            const uint32_t t_o0 = tmem_o0 + lane_sel + g * 64;
            const uint32_t t_o1 = t_o0 + kHeadDim;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                uint32_t o0[32], o1[32];
                tmem_ld_32x32b_x32(t_o0 + q * 32, o0);
                tmem_ld_32x32b_x32(t_o1 + q * 32, o1);
                tmem_wait_ld();
#pragma unroll
                for (int cidx = 0; cidx < 4; ++cidx) {
                    float f[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        f[e] = fmaf(w1, __uint_as_float(o1[cidx * 8 + e]),
                                    w0 * __uint_as_float(o0[cidx * 8 + e]));
                    uint4 v;
                    v.x = pack_16x2<kBF16>(f[0], f[1]);
                    v.y = pack_16x2<kBF16>(f[2], f[3]);
                    v.z = pack_16x2<kBF16>(f[4], f[5]);
                    v.w = pack_16x2<kBF16>(f[6], f[7]);
                    const int chunk = (q * 4 + cidx) ^ (row & 7);
                    *reinterpret_cast<uint4*>(stage + chunk * 16) = v;
                }
            }
            fence_proxy_async_smem();
            named_bar_sync(2 + g, 128);
            if (row == 0) {
                tma_store_4d(&tm_o, smem_base + kSmemQ + g * kHalfBytes, 64 * g, head,
                             qtile * kBlockM, batch);
                tma_store_commit();
                tma_store_wait_read<0>();
            }
        }
    }

    // ------------------------------------ teardown ---------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace fa

// Generation 13 of the B200 attention forward: "KV ping-pong" (CTA pairs only).
//
// Replaces the reference's `flash::flash_forward_kernel` (/root/reference/src/include/forward_kernel.cuh:85-204)
// like the kernels in fa_fwd_sm100.cuh, same arithmetic contract (softmax.cuh:15-128), different machine
// mapping.  Why it exists: in generation 9 both Q tiles of a CTA share ONE S accumulator (tensor memory is
// full: P 128 + S 128 + O 256 columns), so every KV block pays the serial chain
//     tcgen05.ld S -> s_free -> QK^T of the other tile on a cold pipe -> s_full        (~1345 clk, twice),
// 2690 clk per block against 2048 clk of MMA work (profiles/r01_g9_notes.md, section 6).
//
// Here a CTA owns ONE Q tile of 128 rows (a pair: 256 rows, tcgen05 cta_group::2, M = 256), which frees
// 128 tensor-memory columns:
//     [0,128) S_a   [128,256) S_b   [256,320) P_a   [320,384) P_b   [384,512) O
// The two softmax warpgroups work on the SAME rows and alternate KV blocks of the CTA's block stream
// (block G -> warpgroup G & 1, accumulator S_{G&1}, probabilities P_{G&1}); both feed the same O.
// Nothing links consecutive blocks any more except the data itself:
//     S(G+3) is issued as soon as warpgroup (G+1)&1 has read S(G+1) out, ~1000 clk before it is needed;
//     P(G+2) overwrites P(G) only after PV(G) retired (long ago).
// The price: a K/V block now serves 256 instead of 512 query rows, i.e. twice the L2 -> shared-memory
// traffic (32 B/clk/SM; the generation-9 capture shows this path at 10 % of its peak) and 94 instead of
// 80 B/clk of shared-memory operand reads.
//
// Shared row state.  Both warpgroups scale P against the same reference maximum m (the one O is
// accumulated against).  It lives in shared memory (one word per row) and is handed from the owner of block
// G-1 to the owner of block G -- sibling warps on the same SM sub-partition -- behind a 64-thread named
// barrier: the owner of G reads it after its own row max is known, decides whether the lazy rescale is due (max grew by more than 2^8,
// same rule as generation 9), rescales O if so (after PV(G-1) retired), publishes the new m and only then
// spends ~1500 clk on the exponentials -- so the hand-over is ~1500 clk ahead of the consumer.  Each
// warpgroup keeps a partial row sum l_w relative to the m it last saw and re-bases it when m moved; at the
// end of a tile it hands (l_w, m_w) to a separate EPILOGUE warpgroup (512 threads in all), which adds the two
// partial sums, waits for the last PV, scales and stores O -- the softmax warpgroups go straight on to the
// next tile and are never re-synchronised (a shared epilogue put them back in lock step every tile:
// profiles/r02_pp_notes.md).
//
// Two issuing warps (see the control warpgroup): one streams S(G) = Q K_G^T as soon as S_{G&1} was read out, the
// other O += P(G) V_G as soon as P(G) is stored; blocks G = 0, 1, ... run across work tiles, the first PV of a
// tile waits for the previous tile's epilogue to have read O (`o_free`).  K and V travel in separate rings.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#include "fa_fwd_sm100.cuh"

namespace fa {
namespace pp {

constexpr int kKStages = 4, kVStages = 4;         // K ring and V ring slots ...
constexpr int kStages = kKStages + kVStages;
constexpr int kSlotBytes = kTileBytes / 2;        // ... of 16 KiB: 64 keys x 128 d (K) or 128 keys x 64 d (V)
constexpr int kKBoxBytes = kSlotBytes / 2;        // one K TMA box: 64 keys x 64 d columns
constexpr int kSmemQ = 0;                                     // Q tile, double-buffered across work tiles
constexpr int kSmemK = kSmemQ + 2 * kTileBytes;               // K ring
constexpr int kSmemV = kSmemK + kKStages * kSlotBytes;        // V ring
constexpr int kSmemStage = kSmemV + kVStages * kSlotBytes;    // O staging: two TMA boxes of 64 d columns
constexpr int kSmemM = kSmemStage + 2 * kHalfBytes;           // float m[128]: reference max of O, per row
constexpr int kSmemL = kSmemM + 128 * 4;                      // float2 lm[2][128]: (partial row sum, its reference max)
constexpr int kSmemBar = kSmemL + 2 * 128 * 8;                //   of either softmax warpgroup, for the epilogue
constexpr int kNumBarriers = 4 + 2 * kStages + 11 + 8;
constexpr int kSmemTmemPtr = kSmemBar + kNumBarriers * 8;
constexpr int kSmemTotal = kSmemTmemPtr + 16;
constexpr int kSmemLaunchBytes = kSmemTotal;  // the base is 1024-byte aligned by declaration (checked at run time)
constexpr int kThreads = 512;  // warps 0-3 / 4-7 softmax, 8 MMA, 9 TMA, 10-11 idle, 12-15 epilogue
#ifndef FA_PP_REGS_SOFTMAX
#define FA_PP_REGS_SOFTMAX 192
#endif
#ifndef FA_PP_REGS_EPI
#define FA_PP_REGS_EPI 64
#endif
#ifndef FA_PP_REGS_CTRL
#define FA_PP_REGS_CTRL 56
#endif
static_assert(256 * FA_PP_REGS_SOFTMAX + 128 * FA_PP_REGS_EPI + 128 * FA_PP_REGS_CTRL <= 65536, "register pool");
#ifndef FA_PP_PROBE
#define FA_PP_PROBE 1     // test the barriers a softmax warp needs later in a block early (see the softmax loop)
#endif
constexpr bool kProbe = FA_PP_PROBE != 0;
#ifndef FA_PP_SKEW_NS
#define FA_PP_SKEW_NS 500  // one-time delay of warpgroup 1 so the two warpgroups start half a block apart
#endif
static_assert(kSmemLaunchBytes <= 232448, "exceeds the 227 KiB opt-in shared memory of sm_100");

constexpr uint32_t kColS = 0, kColP = 256, kColO = 384;  // S_w at kColS + 128 w, P_w at kColP + 64 w

#ifndef FA_PP_TOKEN
#define FA_PP_TOKEN 0     // 1: the exp2 phases of the two warps that share an SM sub-partition (warp q of either
                          // warpgroup) strictly alternate in block order, so each runs with the MUFU to itself
                          // (16 ex2/clk/SM: ~840 clk per block when alone, twice that when both are in it) and
                          // P(G) completions -- hence PV / S issue -- are spaced evenly instead of in bursts
#endif
constexpr bool kExpToken = FA_PP_TOKEN != 0;
template <bool kBF16, bool kDebug, bool kRagged>
__device__ __forceinline__ void fa_fwd_body_pp(const CUtensorMap& tm_q, const CUtensorMap& tm_k,
                                               const CUtensorMap& tm_v, const CUtensorMap& tm_o,
                                               const FwdParams& prm, const FwdDebug& dbg) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = smem_u32(smem_raw);
    uint8_t* smem_gen = smem_raw;
    if ((smem_base & 1023u) != 0u) {
        if (threadIdx.x == 0) printf("[fa] dynamic shared memory is not 1024-byte aligned\n");
        __trap();
    }
    // warp-uniform warp index (see generation 9 in fa_fwd_sm100.cuh: keeps the MMA warp in uniform registers)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int wg = warp >> 2;
    const uint32_t rank = cluster_ctarank();
    const bool is_leader = rank == 0;
    const int cta_lin = (int)(blockIdx.x >> 1);
    const int n_cta = (int)(gridDim.x >> 1);

    // Barriers (every CTA holds a full set; "leader" = used in the even CTA's copy only, the odd CTA's
    // threads arrive remotely; "both" = signalled in both copies by one multicast tcgen05.commit).
    const uint32_t bar0 = smem_base + kSmemBar;
    auto q_full = [&](int b) { return bar0 + 8u * b; };                       // leader, tx bytes
    auto q_empty = [&](int b) { return bar0 + 8u * (2 + b); };                // both
    auto k_full = [&](int i) { return bar0 + 8u * (4 + i); };                 // leader, tx bytes
    auto k_empty = [&](int i) { return bar0 + 8u * (4 + kKStages + i); };     // both
    auto v_full = [&](int i) { return bar0 + 8u * (4 + 2 * kKStages + i); };              // leader, tx bytes
    auto v_empty = [&](int i) { return bar0 + 8u * (4 + 2 * kKStages + kVStages + i); };  // both
    constexpr int kB = 4 + 2 * kStages;
    auto s_full = [&](int w) { return bar0 + 8u * (kB + w); };                // both
    auto s_free = [&](int w) { return bar0 + 8u * (kB + 2 + w); };            // leader, 8 warps
    auto p_full = [&](int w) { return bar0 + 8u * (kB + 4 + w); };            // leader, 8 warps
    auto p_last = [&](int w) { return bar0 + 8u * (kB + 6 + w); };            // leader, 8 warps
    const uint32_t o_free = bar0 + 8u * (kB + 8);                             // leader, 8 warps (epilogue)
    auto pv_done = [&](int w) { return bar0 + 8u * (kB + 9 + w); };           // both
    auto exp_done = [&](int w, int q) { return bar0 + 8u * (kB + 11 + 4 * w + q); };  // local, 1 warp
    // Hand-overs through generic shared memory use NAMED barriers (bar.arrive by the writer, bar.sync by the reader:
    // cheap, and compute-sanitizer's racecheck understands them, which it does not for mbarrier-ordered accesses):
    //   2 + 4 w + q   reference max of a block published by warp q of warpgroup w        (64 threads)
    //   10 + w        (l, m) of a tile handed from softmax warpgroup w to the epilogue   (256 threads)
    //   12 + w        ... and read by the epilogue: warpgroup w may overwrite its pair    (256 threads)
    //   1             epilogue warpgroup: staging buffer written / reusable               (128 threads)
    auto bar_m = [](int w, int q) { return 2u + 4u * (uint32_t)w + (uint32_t)q; };
    static_assert(kB + 19 == kNumBarriers, "barrier count");

    auto wait = [&](uint32_t bar, uint32_t parity, int tag) { mbar_wait(bar, parity, tag); };
    auto arrive_leader = [&](uint32_t bar) { mbar_arrive_cluster(mapa_shared(bar, 0)); };
    auto commit = [&](uint32_t bar) { umma_commit_2cta(bar, 3); };

    const int n_blocks = prm.n_kv_blocks;
    const int n_local = cta_lin < prm.n_tiles ? (prm.n_tiles - cta_lin + n_cta - 1) / n_cta : 0;
    struct TileCoord {
        int q_row0, head, batch;
    };
    auto coord_of = [&](int it) {  // it-th work tile of this CTA pair
        const int tile = cta_lin + it * n_cta;
        TileCoord tc;
        const int qgroup = tile % prm.n_q_groups;
        const int bh = tile / prm.n_q_groups;
        tc.q_row0 = qgroup * (2 * kBlockM) + (int)rank * kBlockM;
        tc.head = bh % prm.n_heads;
        tc.batch = bh / prm.n_heads;
        return tc;
    };

#if FA_HANG_GUARD
    if constexpr (kDebug) {
        if (threadIdx.x == 0) g_fa_diag = dbg.diag;
    }
#endif
    if (warp == 8) {
        if (lane == 0) {
            for (int b = 0; b < 2; ++b) {
                mbar_init(q_full(b), 1);
                mbar_init(q_empty(b), 1);
                mbar_init(s_full(b), 1);
                mbar_init(s_free(b), 8);
                mbar_init(p_full(b), 8);
                mbar_init(p_last(b), 8);
                mbar_init(pv_done(b), 1);
                for (int q = 0; q < 4; ++q) mbar_init(exp_done(b, q), 1);
            }
            mbar_init(o_free, 8);
            for (int i = 0; i < kKStages; ++i) {
                mbar_init(k_full(i), 1);
                mbar_init(k_empty(i), 1);
            }
            for (int i = 0; i < kVStages; ++i) {
                mbar_init(v_full(i), 1);
                mbar_init(v_empty(i), 1);
            }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc_2cta(smem_base + kSmemTmemPtr, kTmemCols);
        tmem_relinquish_2cta();
    } else if (warp == 9 && lane == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_k);
        tma_prefetch_desc(&tm_v);
        tma_prefetch_desc(&tm_o);
    }
    tc_fence_before();
    cluster_sync();   // the peer's barriers exist before anything targets them
    __syncthreads();  // (orders the TMEM-address word for compute-sanitizer's racecheck, which does not model
                      // barrier.cluster; costs one CTA barrier per launch)
    tc_fence_after();
    if constexpr (kDebug) {  // (racecheck cannot order tcgen05.alloc's write of this word across the CTA pair)
        if (*reinterpret_cast<volatile uint32_t*>(smem_gen + kSmemTmemPtr) != 0u) {
            if (threadIdx.x == 0) printf("[fa] unexpected TMEM base address\n");
            __trap();
        }
    }
    constexpr uint32_t tmem_base = 0u;  // all 512 columns are ours: literal keeps tcgen05 operands uniform
    const bool run = !(kDebug && dbg.level == 1u);  // bring-up level 1: setup and teardown only
    // Cycle trace (debug level 5; tools/gpu_pp_trace.py): %clock stamps of CTA 0 for its SECOND work tile
    // (steady state), as 32-bit words behind kTraceBase in the dump buffer:
    //   [0,256)   softmax [w][k < 16][8]: 0 before s_full wait, 1 S ready, 2 S in registers (s_free), 3 m published,
    //             4 exp2 token held, 5 P_w free, 6 p_full, 7 p_last
    //   [256,384) PV issue [j < 32][4]: 0 at the waits, 1 V + P (96 keys) seen, 2 P (all) seen, 3 committed
    //   [384,512) S issue  [j < 32][4]: 0 at the waits, 2 K landed + accumulator free, 3 committed
    const bool tracing = kDebug && dbg.level == 5u && dbg.dump != nullptr && blockIdx.x == 0;
    uint32_t* const trace = tracing ? reinterpret_cast<uint32_t*>(dbg.dump) + kTraceBase : nullptr;
    constexpr int kTraceTile = 1;

    if (wg == 2) {
        setmaxnreg_dec<FA_PP_REGS_CTRL>();
        // Control warpgroup: FOUR independent in-order streams.  A cycle trace of the first version (one producer,
        // one issuing warp; profiles/r02_pp_notes.md) showed the single issuing warp needing ~1400 clk of its own
        // instruction time per KV block -- five mbarrier waits at ~90 clk each even when long complete, 16
        // tcgen05.mma at ~25-30 clk, five commits -- against 1024 clk of tensor work: the issuing warp, not the
        // tensor pipe and not the softmax, set the pace.  Split in two, each stream needs ~500 clk per block.
        //   warp 8  issues S(x) = Q K^T   as soon as K(x) landed and S_{x&1} was read out
        //   warp 10 issues O += P(g) V    as soon as V(g) landed and P(g) is stored
        //   warp 9  loads Q and K,  warp 11 loads V  (own rings, so neither stream ever waits for the other's slot)
        // Nothing orders S against PV except the data (different tensor-memory columns); every tcgen05.commit tracks
        // the MMAs of its own thread.
        const int total = n_local * n_blocks;
        auto load_box = [&](uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int row0, const TileCoord& tc) {
            tma_load_4d_2cta(dst, map, mapa_shared(bar, 0), c0, tc.head, row0, tc.batch);
        };
        if (warp == 9 && run) {
            // ============================ TMA producer: Q and K ============================
            // every CTA loads its own Q rows and keys [64 rank, +64) of a K block; bytes of both CTAs are
            // credited to the leader's barrier
            int it = 0, j = 0;
            TileCoord tc = coord_of(0);
            for (int x = 0; x < total; ++x) {
                if (j == 0) {
                    tc = coord_of(it);
                    const int b = it & 1;
                    wait(q_empty(b), (uint32_t)(((it >> 1) & 1) ^ 1), 110 + b);
                    if (elect_one()) {
                        const uint32_t dst = smem_base + kSmemQ + b * kTileBytes;
                        if (is_leader) mbar_arrive_expect_tx(q_full(b), 2 * kTileBytes);
                        load_box(dst, &tm_q, q_full(b), 0, tc.q_row0, tc);
                        load_box(dst + kHalfBytes, &tm_q, q_full(b), 64, tc.q_row0, tc);
                    }
                    __syncwarp();
                }
                const int slot = x % kKStages;
                wait(k_empty(slot), (uint32_t)(((x / kKStages) & 1) ^ 1), 100 + slot);
                if (elect_one()) {
                    const uint32_t dst = smem_base + kSmemK + slot * kSlotBytes;
                    const int row0 = j * kBlockN + 64 * (int)rank;
                    if (is_leader) mbar_arrive_expect_tx(k_full(slot), kTileBytes);
                    load_box(dst, &tm_k, k_full(slot), 0, row0, tc);
                    load_box(dst + kKBoxBytes, &tm_k, k_full(slot), 64, row0, tc);
                }
                __syncwarp();
                if (++j == n_blocks) {
                    j = 0;
                    ++it;
                }
            }
        } else if (warp == 11 && run) {
            // ============================== TMA producer: V ================================
            // d columns [64 rank, +64) of every V block
            int it = 0, j = 0;
            TileCoord tc = coord_of(0);
            for (int g = 0; g < total; ++g) {
                if (j == 0) tc = coord_of(it);
                const int slot = g % kVStages;
                wait(v_empty(slot), (uint32_t)(((g / kVStages) & 1) ^ 1), 120 + slot);
                if (elect_one()) {
                    const uint32_t dst = smem_base + kSmemV + slot * kSlotBytes;
                    if (is_leader) mbar_arrive_expect_tx(v_full(slot), kTileBytes);
                    load_box(dst, &tm_v, v_full(slot), 64 * (int)rank, j * kBlockN, tc);
                }
                __syncwarp();
                if (++j == n_blocks) {
                    j = 0;
                    ++it;
                }
            }
        } else if (warp == 8 && is_leader && run) {
            // ============================== MMA issuer: S = Q K^T ===========================
            // Convergent warp, one elected lane issues (see fa_fwd_sm100.cuh).  M = 256 across the pair.
            constexpr uint32_t idesc_qk = umma_idesc_f16(kBF16, 2 * kBlockM, kBlockN, false);
            int it = 0, j = 0;
            for (int x = 0; x < total; ++x) {
                const int w = x & 1;
                const int slot = x % kKStages;
                const uint64_t b0 = umma_smem_desc_sw128(smem_base + kSmemK + slot * kSlotBytes, 16, 1024);
                const uint64_t a0 = umma_smem_desc_sw128(smem_base + kSmemQ + (it & 1) * kTileBytes, 16, 1024);
                uint32_t* tr = nullptr;
                if constexpr (kDebug) {
                    if (trace != nullptr && it == kTraceTile && j < 32 && lane == 0) tr = trace + 384 + j * 4;
                    if (tr) tr[0] = clk32();
                }
                if (j == 0) wait(q_full(it & 1), (uint32_t)((it >> 1) & 1), 210);
                // K(x) landed; warpgroup w has read S(x-2) out of S_w (fresh barrier: parity 1 passes)
                mbar_wait2(k_full(slot), (uint32_t)((x / kKStages) & 1), s_free(w), (uint32_t)(((x >> 1) & 1) ^ 1), 200);
                if constexpr (kDebug) {
                    if (tr) tr[2] = clk32();
                }
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < kHeadDim / 16; ++k) {
                        const uint32_t a_off = ((k >> 2) * kHalfBytes + (k & 3) * 32) >> 4;
                        const uint32_t b_off = ((k >> 2) * kKBoxBytes + (k & 3) * 32) >> 4;
                        umma_ss_2cta(tmem_base + kColS + 128u * w, a0 + a_off, b0 + b_off, idesc_qk, k > 0);
                    }
                    commit(s_full(w));
                    commit(k_empty(slot));
                    if (j + 1 == n_blocks) commit(q_empty(it & 1));  // last use of this Q buffer
                }
                __syncwarp();
                if constexpr (kDebug) {
                    if (tr) tr[3] = clk32();
                }
                if (++j == n_blocks) {
                    j = 0;
                    ++it;
                }
            }
        } else if (warp == 10 && is_leader && run) {
            // ============================== MMA issuer: O += P V ============================
            constexpr uint32_t idesc_pv = umma_idesc_f16(kBF16, 2 * kBlockM, kHeadDim, true);
            int it = 0, j = 0;
            for (int g = 0; g < total; ++g) {
                const int w = g & 1;
                const int slot = g % kVStages;
                const uint64_t b0 = umma_smem_desc_sw128(smem_base + kSmemV + slot * kSlotBytes, kHalfBytes, 1024);
                const uint32_t par = (uint32_t)((g >> 1) & 1);
                uint32_t* tr = nullptr;
                if constexpr (kDebug) {
                    if (trace != nullptr && it == kTraceTile && j < 32 && lane == 0) tr = trace + 256 + j * 4;
                    if (tr) tr[0] = clk32();
                }
                // V(g) landed; first 96 keys of P(g) stored, O rescaled if due
                mbar_wait2(v_full(slot), (uint32_t)((g / kVStages) & 1), p_full(w), par, 220);
                if (j == 0) wait(o_free, (uint32_t)((it & 1) ^ 1), 260);  // previous epilogue has read O
                if constexpr (kDebug) {
                    if (tr) tr[1] = clk32();
                }
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 6; ++k)
                        umma_ts_2cta(tmem_base + kColO, tmem_base + kColP + 64u * w + k * 8, b0 + ((k * 2048) >> 4),
                                     idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
                }
                __syncwarp();
                wait(p_last(w), par, 250 + w);  // last 32 keys
                if constexpr (kDebug) {
                    if (tr) tr[2] = clk32();
                }
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int k = 6; k < 8; ++k)
                        umma_ts_2cta(tmem_base + kColO, tmem_base + kColP + 64u * w + k * 8, b0 + ((k * 2048) >> 4),
                                     idesc_pv, 1u);
                    commit(pv_done(w));
                    commit(v_empty(slot));
                }
                __syncwarp();
                if constexpr (kDebug) {
                    if (tr) tr[3] = clk32();
                }
                if (++j == n_blocks) {
                    j = 0;
                    ++it;
                }
            }
        }
        __syncwarp();
    } else if (wg < 2 && run) {
        // ==================================== softmax =====================================
        setmaxnreg_inc<FA_PP_REGS_SOFTMAX>();
        const int w = wg;                    // this warpgroup owns blocks with (G & 1) == w
        const int row = threadIdx.x & 127;   // row of the Q tile == TMEM lane
        const int wq = warp & 3;
        const uint32_t lane_sel = static_cast<uint32_t>(wq * 32) << 16;
        const uint32_t t_s = tmem_base + lane_sel + kColS + 128u * w;
        const uint32_t t_p = tmem_base + lane_sel + kColP + 64u * w;
        const uint32_t t_o = tmem_base + lane_sel + kColO;
        const float c = prm.scale_log2;
        const int kv_tail = prm.seq_len & (kBlockN - 1);
        const uint32_t slot_m = smem_base + kSmemM + 4u * row;  // this row's shared state word
        float2* sm_lm = reinterpret_cast<float2*>(smem_gen + kSmemL);
        if (FA_PP_SKEW_NS > 0 && w == 1) __nanosleep(FA_PP_SKEW_NS);

        bool s_probed = false;  // s_full of the next block already seen complete
        for (int it = 0; it < n_local; ++it) {
            const int g0 = it * n_blocks;  // stream index of this tile's first block
            float l_run = 0.f;             // partial row sum of this warpgroup, relative to m_known
            float m_known = 0.f;           // the reference max this warpgroup last saw
            bool have_l = false;           // l_run holds something (this warpgroup did a block of this tile)
            for (int j = (g0 + w) & 1; j < n_blocks; j += 2) {
                const int g = g0 + j;
                const uint32_t par = (uint32_t)((g >> 1) & 1);
                uint32_t sr[4][32];
                uint32_t* tr = nullptr;
                if constexpr (kDebug) {
                    if (trace != nullptr && it == kTraceTile && (j >> 1) < 16 && wq == 0 && lane == 0)
                        tr = trace + (w * 16 + (j >> 1)) * 8;
                    if (tr) tr[0] = clk32();
                }
                if (!s_probed) wait(s_full(w), par, 300 + w);
                if constexpr (kDebug) {
                    if (tr) tr[1] = clk32();
                }
                tc_fence_after();
                float m_lo = -INFINITY;
                // Probes: a mbarrier.try_wait costs ~90 clk even on a phase that completed long ago, and this
                // warpgroup's own time per block is what paces the kernel.  The barriers needed later in the block
                // (m of block g-1 published, PV(g-2) retired) are tested here, under the tensor-memory load, and
                // only waited for at the point of use if the probe failed.
                bool p_probed = false;
                float m_prev = 0.f;
                if constexpr (!kRagged) {
                    tmem_ld_32x32b_x32(t_s, sr[0]);
                    tmem_ld_32x32b_x32(t_s + 32, sr[1]);
                    if constexpr (kProbe) p_probed = mbar_test_wait(pv_done(w), par ^ 1u);
                    tmem_wait_ld();
                    tmem_ld_32x32b_x32(t_s + 64, sr[2]);
                    tmem_ld_32x32b_x32(t_s + 96, sr[3]);
                    m_lo = row_max_frags<0, 2>(sr);
                    tmem_wait_ld();
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) tmem_ld_32x32b_x32(t_s + q * 32, sr[q]);
                    tmem_wait_ld();
                }
                // S is in registers: S_w may receive S(g+2)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) arrive_leader(s_free(w));
                if constexpr (kDebug) {
                    if (tr) tr[2] = clk32();
                }
                if constexpr (kDebug) {  // S of the first block of the first tile of CTA pair 0 (both CTAs)
                    if (dbg.dump != nullptr && blockIdx.x < 2 && it == 0 && j == 0) {
                        for (int q = 0; q < 4; ++q)
                            for (int i = 0; i < 32; ++i)
                                dbg.dump[((int)rank * 128 + row) * 128 + q * 32 + i] = __uint_as_float(sr[q][i]);
                    }
                }
                float mx;
                if constexpr (!kRagged) {
                    mx = fmaxf(m_lo, row_max_frags<2, 2>(sr));
                } else {
                    if (j + 1 == n_blocks && kv_tail != 0) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (q * 32 + i >= kv_tail) sr[q][i] = 0xff800000u;  // -inf: P = 0
                    }
                    mx = row_max_128(sr);
                }

                // ---- shared row state: the reference max of O ----
                // Decision of block g-1 (sibling warp of the other warpgroup, same rows) published.  Also awaited by the
                // first block of a tile, which does not use the value: the slot has one writer at a time, in block order.
                if (g > 0) named_bar_sync(bar_m(w ^ 1, wq), 64);
                const uint32_t slot_v = lds_volatile_u32(slot_m);
                m_prev = __uint_as_float(slot_v);
                float m_cur;
                if (j == 0) {
                    m_cur = mx;  // first block of the tile: O is overwritten by PV(g) (accumulate = 0)
                } else {
                    if (have_l) l_run *= ex2_approx((m_known - m_prev) * c);  // exactly 1 when m did not move
                    const float delta = (mx - m_prev) * c;
                    const bool need = delta > kRescaleThreshold;
                    m_cur = m_prev;
                    if (__any_sync(0xffffffffu, need)) {
                        float alpha = 1.f;
                        if (need) {
                            alpha = ex2_approx(-delta);
                            m_cur = mx;
                        }
                        // O must be quiescent: PV(g-1) retired (PV(g) needs our P)
                        wait(pv_done(w ^ 1), (uint32_t)(((g - 1) >> 1) & 1), 320 + w);
                        tc_fence_after();
#pragma unroll 1
                        for (int q = 0; q < 8; ++q) {  // 16 columns at a time: S(g) stays in registers
                            uint32_t o[16];
                            tmem_ld_32x32b_x16(t_o + q * 16, o);
                            tmem_wait_ld();
#pragma unroll
                            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                            tmem_st_32x32b_x16(t_o + q * 16, o);
                        }
                        l_run *= alpha;
                    }
                }
                // publish the reference max of O after this block (one word per row)
                sts_volatile_u32(slot_m, __float_as_uint(m_cur));
                named_bar_arrive(bar_m(w, wq), 64);  // the owner of block g+1 (sibling warp) may read it
                m_known = m_cur;
                have_l = true;
                if constexpr (kDebug) {
                    if (tr) tr[3] = clk32();
                }

                // ---- P = exp2(S c - m c), row sum, 16-bit P into tensor memory ----
                if constexpr (kExpToken) {
                    // the sibling warp (same sub-partition, other warpgroup) has issued the exponentials of block g-1
                    if (g > 0) wait(exp_done(w ^ 1, wq), (uint32_t)(((g - 1) >> 1) & 1), 360 + w);
                }
                if constexpr (kDebug) {
                    if (tr) tr[4] = clk32();
                }
                const float neg_mc = -m_cur * c;
                const float2 c2 = make_float2(c, c);
                const float2 nm2 = make_float2(neg_mc, neg_mc);
                float2 sum_a = make_float2(0.f, 0.f), sum_b = make_float2(0.f, 0.f);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t pk[16];
                    if (q == 3) exp_fragment<kBF16, kEmuPairsLast, FA_EXP_VARIANT>(sr[q], c2, nm2, sum_a, sum_b, pk);
                    else exp_fragment<kBF16, kEmuPairs, FA_EXP_VARIANT>(sr[q], c2, nm2, sum_a, sum_b, pk);
                    if constexpr (kExpToken) {
                        if (q == 3) {
                            __syncwarp();
                            if (lane == 0) mbar_arrive(exp_done(w, wq));
                        }
                    }
                    if (q == 0) {
                        // P_w still holds P(g-2): PV(g-2) must have read it (fresh barrier: parity 1 passes)
                        if (!p_probed) wait(pv_done(w), par ^ 1u, 330 + w);
                        tc_fence_after();
                        if constexpr (kDebug) {
                            if (tr) tr[5] = clk32();
                        }
                    }
                    tmem_st_32x32b_x16(t_p + q * 16, pk);
                    if (q == 2) {
                        tmem_wait_st();  // also covers the rescaled O
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) arrive_leader(p_full(w));
                        if constexpr (kDebug) {
                            if (tr) tr[6] = clk32();
                        }
                    }
                }
                if constexpr (kProbe) {  // S of this warpgroup's next block (normally there since ~1000 clk)
                    s_probed = mbar_test_wait(s_full(w), par ^ 1u);
                }
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) arrive_leader(p_last(w));
                if constexpr (kDebug) {
                    if (tr) tr[7] = clk32();
                }
                l_run += (sum_a.x + sum_a.y) + (sum_b.x + sum_b.y);
            }

            // hand (l_w, m_w) of this tile to the epilogue warpgroup and move on to the next tile
            if (it > 0) named_bar_sync(12 + w, 256);  // the previous tile's pair was read
            sm_lm[w * 128 + row] = have_l ? make_float2(l_run, m_known) : make_float2(0.f, -INFINITY);
            named_bar_arrive(10 + w, 256);
        }
    } else if (wg == 3 && run) {
        // =================================== epilogue =====================================
        // Own warpgroup (FlashAttention-4 calls it the correction warpgroup) so that the softmax warpgroups go
        // straight from the last block of a tile to the first block of the next and are never re-synchronised:
        // O / l -> 16 bit -> swizzled shared memory -> TMA store, two 64-column boxes per tile.
        setmaxnreg_dec<FA_PP_REGS_EPI>();
        const int row = threadIdx.x & 127;
        const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const uint32_t t_o = tmem_base + lane_sel + kColO;
        const float c = prm.scale_log2;
        const float2* sm_lm = reinterpret_cast<const float2*>(smem_gen + kSmemL);
        for (int it = 0; it < n_local; ++it) {
            const int g_last = (it + 1) * n_blocks - 1;
            const int w_last = g_last & 1;
            named_bar_sync(10, 256);
            named_bar_sync(11, 256);
            const float2 a0 = sm_lm[row], a1 = sm_lm[128 + row];
            named_bar_arrive(12, 256);
            named_bar_arrive(13, 256);
            // O is accumulated against the last owner's reference max; the other partial sum is re-based
            const float m_fin = w_last ? a1.y : a0.y;
            const float l = a0.x * ex2_approx((a0.y - m_fin) * c) + a1.x * ex2_approx((a1.y - m_fin) * c);
            const float inv_l = 1.0f / l;
            if constexpr (kDebug) {  // l and m of the first tile of CTA pair 0
                if (dbg.dump != nullptr && blockIdx.x < 2 && it == 0) {
                    dbg.dump[2 * 128 * 128 + (int)rank * 128 + row] = l;
                    dbg.dump[2 * 128 * 128 + 256 + (int)rank * 128 + row] = m_fin;
                }
            }
            // last PV of the tile retired (in-order pipe: all of them).  Both softmax warpgroups have finished
            // the tile, so pv_done(w_last) is already in the phase waited for (no parity aliasing).
            wait(pv_done(w_last), (uint32_t)((g_last >> 1) & 1), 310);
            tc_fence_after();
            const TileCoord tc = coord_of(it);
#pragma unroll
            for (int q = 0; q < 4; ++q) {  // 32 columns at a time (64 registers per thread in this warpgroup)
                const int h = q >> 1;      // TMA box / staging buffer
                uint32_t o[32];
                tmem_ld_32x32b_x32(t_o + 32 * q, o);
                tmem_wait_ld();
                if (q == 3) {  // O is in registers: the next tile's first PV may overwrite it
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) arrive_leader(o_free);
                }
                if ((q & 1) == 0) {
                    // staging buffer h: the previous tile's store of this box must have read it (bulk groups of
                    // the issuing thread complete in order: at most one newer group may still be pending)
                    if (row == 0) tma_store_wait_read<1>();
                    named_bar_sync(1, 128);
                }
                uint8_t* stage_row = smem_gen + kSmemStage + h * kHalfBytes + row * 128;
#pragma unroll
                for (int cidx = 0; cidx < 4; ++cidx) {
                    uint4 v;
                    v.x = pack_16x2<kBF16>(__uint_as_float(o[cidx * 8 + 0]) * inv_l, __uint_as_float(o[cidx * 8 + 1]) * inv_l);
                    v.y = pack_16x2<kBF16>(__uint_as_float(o[cidx * 8 + 2]) * inv_l, __uint_as_float(o[cidx * 8 + 3]) * inv_l);
                    v.z = pack_16x2<kBF16>(__uint_as_float(o[cidx * 8 + 4]) * inv_l, __uint_as_float(o[cidx * 8 + 5]) * inv_l);
                    v.w = pack_16x2<kBF16>(__uint_as_float(o[cidx * 8 + 6]) * inv_l, __uint_as_float(o[cidx * 8 + 7]) * inv_l);
                    const int chunk = ((q & 1) * 4 + cidx) ^ (row & 7);  // TMA 128B swizzle
                    *reinterpret_cast<uint4*>(stage_row + chunk * 16) = v;
                }
                if (q & 1) {
                    fence_proxy_async_smem();
                    named_bar_sync(1, 128);
                    if (row == 0) {
                        tma_store_4d(&tm_o, smem_base + kSmemStage + h * kHalfBytes, 64 * h, tc.head, tc.q_row0,
                                     tc.batch);
                        tma_store_commit();
                    }
                }
            }
        }
        if (row == 0) tma_store_wait_read<0>();  // shared memory must outlive the last store's read
    }

    // ------------------------------------ teardown ---------------------------------------
    tc_fence_before();
    cluster_sync();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc_2cta(tmem_base, kTmemCols);
    }
}

template <bool kBF16, bool kDebug, bool kRagged>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
fa_fwd_kernel_pp(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o,
                 const FwdParams prm, const FwdDebug dbg) {
    fa_fwd_body_pp<kBF16, kDebug, kRagged>(tm_q, tm_k, tm_v, tm_o, prm, dbg);
}

}  // namespace pp
}  // namespace fa

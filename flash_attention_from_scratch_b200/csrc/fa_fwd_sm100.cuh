// Fused attention forward for NVIDIA B200 (sm_100a):  O = softmax(Q K^T / sqrt(d)) V
//
// Replaces the reference's Ampere kernel `flash::flash_forward_kernel`
// (/root/reference/src/include/forward_kernel.cuh:85-204) and its device primitives
// (gemm.cuh / softmax.cuh / load_store.cuh).  Same arithmetic contract
// (/root/reference/src/include/softmax.cuh:15-128, forward_kernel.cuh:150-152):
//     c = log2(e)/sqrt(d);  m = running row max of raw S;  P = exp2(S*c - m*c) in fp32;
//     l += rowsum(P) (un-rounded fp32);  O += rn16(P) V (fp32 accumulate);  out = rn16(O / l)
// but a different machine mapping -- nothing of the mma.sync / ldmatrix / cp.async path is kept:
//
//   * PERSISTENT: every CTA walks a static list of work tiles.  A CTA owns 2 Q tiles of 128 rows
//     (256 query rows of one (batch, head)); consecutive tiles share K/V in L2.  Loads, MMAs and the
//     epilogue of neighbouring tiles overlap.
//   * TMA (cp.async.bulk.tensor, 128B swizzle) stages Q tiles and K/V blocks through shared
//     memory guarded by full/empty mbarriers.
//   * two warps issue tcgen05.mma (one elected thread each): S_s = Q_s K_j^T (operands from smem) and
//     O_s += P_s V_j (P read from tensor memory, V from smem, MN-major); accumulators in tensor memory.
//   * two softmax warpgroups (one per Q tile, one thread per row, no shuffles): tcgen05.ld S,
//     fp32 row max / exp2 / row sum in registers, P written back to TMEM as packed 16-bit in two
//     parts (96 + 32 columns) so PV starts early, lazy rescale of O (only when the row max grew
//     by more than 2^8).
//   * an epilogue warpgroup (generation 15): final 1/l scaling and TMA store of O through swizzled shared memory,
//     off the softmax warpgroups' path -- they go straight from a tile's last block to the next tile's first.
//
// History (profiles/r01_*_notes.md, profiles/r02_pp_notes.md have the measurements behind each step; the
// losing protocols are gone from this file, their patches are archived under profiles/):
//   generation 6: ONE S accumulator shared by both Q tiles, P_s in its own tensor-memory columns; S_s(j+1) is
//       issued as soon as the other warpgroup has read the previous S out, i.e. while softmax_s(j) is running.
//   generation 7 (kPair): two CTAs of a cluster form a tcgen05 `cta_group::2` pair.  The even CTA issues M = 256
//       MMAs for both; each CTA keeps its own Q tiles, softmax and epilogue but loads only HALF of every K block
//       (64 keys) and V block (64 of the 128 d columns): 6 KiB instead of 8 KiB of smem operands per MMA.
//   generation 9 (FA_UNIFORM_WARP, FA_LD_SPLIT): the warp index is broadcast with shfl so that ptxas keeps the
//       issuing warps' counters and descriptors in UNIFORM registers (without it ~26 R2UR sat between every
//       barrier wait and the first UTCHMMA: +13 % single, +2 % pair), and the softmax warps reduce the row max of
//       S[:, :64] while S[:, 64:] is still being fetched.
//   generation 14: TWO MMA-issuing warps (warp 8: the QK^T groups, warp 10: the PV groups) and separate K and V
//       rings.  A cycle trace of the ping-pong kernel (fa_fwd_pp_sm100.cuh) showed one in-order issuing warp
//       needing ~1400 clk of its own instruction time per KV block -- mbarrier waits at ~90 clk even when long
//       complete, ~25-30 clk per tcgen05.mma, commits -- and in this kernel that time sits inside the serial
//       chain through the shared S accumulator (pair kernel +2 % at seq_len >= 8192, single +10 % at 512).
//   generation 15: the epilogue in its own warpgroup (512 threads, setmaxnreg 192 / 64 / 56).  A fit of tile time
//       against block count had shown 1.5-3 us per work tile that was not block work (last PV -> read O -> convert
//       -> store -> only then the next tile's first softmax): 1513 vs 1375 TFLOP/s at the headline on the same box.
//       The TMA producer also asks for the next work tile's Q rows in L2.  What paces the kernel now, measured with
//       -DFA_TRACE=1 cycle stamps: the serial chain through the shared S accumulator (profiles/r02_g15_notes.md).
//   Removed after measurement: generation 4b (P aliased onto per-tile S, FlashAttention-4's layout: 1399 vs 1450),
//   generation 8 (S prefetched into the exp2 shadow), generation 10 (S in 64-column halves), generation 11 (P
//   through shared memory: 1422 vs 1464, shared-memory bandwidth), generation 16 (PV_s(j) issued only behind
//   S_s(j+1): noise), the softmax-side epilogue of generations 1-14 (r02_g14_softmax_side_epilogue_removed.patch).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#include "ptx_sm100.cuh"
#include "softmax_sm100.cuh"

namespace fa {

constexpr int kBlockM = 128;   // query rows per Q tile == TMEM lanes
constexpr int kBlockN = 128;   // key/value rows per block
constexpr int kHeadDim = 128;  // d_head (the only one the reference supports, README.md:9-15)
constexpr int kQStages = 2;    // Q tiles per CTA
#ifndef FA_KV_STAGES
#define FA_KV_STAGES 4
#endif
constexpr int kKVStages = FA_KV_STAGES;  // K/V ring slots of 32 KiB (a CTA pair uses 2x as many 16 KiB slots)
constexpr int kMaxStages = 2 * kKVStages;
constexpr int kTileBytes = kBlockN * kHeadDim * 2;  // 32 KiB: one 128x128 16-bit tile
constexpr int kHalfBytes = kTileBytes / 2;          // one TMA box: 128 rows x 64 cols (128 B rows)
constexpr int kNumThreads = 512;                    // 2 softmax warpgroups + control warpgroup + epilogue warpgroup
constexpr int kTmemCols = 512;

constexpr int kSmemQ = 0;                                       // Q_0, Q_1
constexpr int kSmemKV = kSmemQ + kQStages * kTileBytes;         // K/V ring
constexpr int kSmemStage = kSmemKV + kKVStages * kTileBytes;    // O staging: 16 KiB per Q tile
constexpr int kSmemL = kSmemStage + kQStages * kHalfBytes;     // float l[2][128]: row sums handed to the epilogue
constexpr int kSmemBar = kSmemL + kQStages * kBlockM * 4;
constexpr int kNumBarriers = 4 + 2 * kMaxStages + 13;
constexpr int kSmemTmemPtr = kSmemBar + kNumBarriers * 8;
constexpr int kSmemTotal = kSmemTmemPtr + 16;
constexpr int kSmemLaunchBytes = kSmemTotal + 1024;  // slack (the base is 1024-byte aligned by declaration)
static_assert(kSmemLaunchBytes <= 232448, "exceeds the 227 KiB opt-in shared memory of sm_100");

// Tunables (overridable with -D at build time; tools/build_variants.py sweeps them).
#ifndef FA_EMU_PAIRS
#define FA_EMU_PAIRS 4        // of every 16 (p0,p1) pairs, how many use the FMA-pipe exp2 (rest: MUFU)
#endif
#ifndef FA_EMU_PAIRS_LAST
#define FA_EMU_PAIRS_LAST 0   // same for the last 32-column fragment
#endif
#ifndef FA_SPLIT_P
#define FA_SPLIT_P 1          // 1: signal the MMA warp after 96 of 128 P columns, again after the rest
#endif
#ifndef FA_REGS_SOFTMAX
#define FA_REGS_SOFTMAX 192   // setmaxnreg for the softmax warpgroups ...
#endif
#ifndef FA_REGS_CTRL
#define FA_REGS_CTRL 64       // ... the control warpgroup ...
#endif
#ifndef FA_REGS_EPI
#define FA_REGS_EPI 56        // ... and the epilogue warpgroup (reads O 32 columns at a time)
#endif
#ifndef FA_TRACE
#define FA_TRACE 0            // 1: the production instantiations can record the cycle trace too (development builds)
#endif
#ifndef FA_Q_PREFETCH
#define FA_Q_PREFETCH 1       // the TMA producer asks for the NEXT work tile's Q rows in L2 while the current tile runs
#endif
constexpr int kEmuPairs = FA_EMU_PAIRS;
constexpr int kEmuPairsLast = FA_EMU_PAIRS_LAST;
constexpr bool kSplitP = FA_SPLIT_P != 0;
static_assert(256 * FA_REGS_SOFTMAX + 128 * FA_REGS_CTRL + 128 * FA_REGS_EPI <= 512 * 128, "register pool exceeded");
#ifndef FA_EXP_VARIANT
#define FA_EXP_VARIANT 0      // code shape of exp_fragment (softmax_sm100.cuh)
#endif
#ifndef FA_QK_ONE_ASM
#define FA_QK_ONE_ASM 1       // CTA-pair kernel: a QK^T group is one asm statement (see umma_ss_2cta_k8)
#endif
#ifndef FA_UNIFORM_WARP
#define FA_UNIFORM_WARP 1
#endif
#ifndef FA_LD_SPLIT
#define FA_LD_SPLIT 1         // 1: fetch S in two halves and reduce the row max of the first 64 columns
                              // while the tcgen05.ld of the last 64 is in flight (TMEM reads run at
                              // ~64 B/clk per sub-partition: 256 clk for the 16 KiB a warp owns)
#endif
static_assert(kKVStages >= 4, "generation 6/7 keep V_j, K_j+1, V_j+1, K_j+2 in flight");

// Tensor-memory column map (512 columns, base 0):
//   [0,64) P_0   [64,128) P_1   [128,256) S (shared by both Q tiles)   [256,384) O_0   [384,512) O_1
__host__ __device__ constexpr uint32_t tmem_col_s() { return 128u; }
__host__ __device__ constexpr uint32_t tmem_col_p(int s) { return static_cast<uint32_t>(s) * 64u; }
__host__ __device__ constexpr uint32_t tmem_col_o(int s) {
    return 256u + static_cast<uint32_t>(s) * 128u;
}

// Lazy rescale threshold in log2 units: O and l are only rescaled when the running max grows by
// more than this; until then P is computed against the stale max, i.e. P <= 2^8 (exact in fp32,
// representable in bf16/fp16).  The final O/l is unaffected.
constexpr float kRescaleThreshold = 8.0f;

// Bring-up hooks (only read by the kDebug instantiation; see tools/gpu_bringup.py).
struct FwdDebug {
    float* dump;      // level 2: raw smem Q_0 | first 32 KiB of the K/V ring (2 x 8192 words);
                      // level >= 3: S(block 0) [2][128][128] of work tile 0, then l [2][128], m [2][128]
    uint32_t level;   // 1: setup/teardown only, 2: + TMA Q_0,K_0, 3: + S = QK^T, >= 4: everything,
                      // 5: everything + cycle trace of CTA 0 / tile 0 (words from kTraceBase on:
                      // softmax [stage][block<32][8 events], PV issue [block<32][stage][4 events],
                      // S issue [block<32][stage][4 events], tensor-pipe observer [block<32][4 events])
                      // 6..9: level 5 with parts of the kernel switched off (timing ablations, results
                      // are garbage): 6 = no softmax arithmetic (S is read, a constant P is stored),
                      // 7 = K/V ring slots are not refilled by TMA after the first pass, 8 = 6 + 7,
                      // 9 = 8 + S is not read out of tensor memory either
    uint32_t* diag;   // host-mapped diagnostics ring (hang-guard builds)
};

constexpr int kTraceBase = 2 * 128 * 128 + 512;  // word offset of the trace inside FwdDebug::dump
constexpr int kTraceTile = 1;                    // the CTA's work tile the cycle trace records (steady state)
__device__ __forceinline__ uint32_t clk32() {
    uint32_t c;
    asm volatile("mov.u32 %0, %%clock;" : "=r"(c));
    return c;
}

struct FwdParams {
    int batch;
    int seq_len;
    int n_heads;
    int n_kv_blocks;   // ceil(seq_len / 128); keys beyond seq_len in the last block are masked
    int n_q_groups;    // work tiles per (batch, head): ceil(seq_len / 256), CTA pairs: ceil(seq_len / 512)
    int n_tiles;       // batch * n_heads * n_q_groups
    float scale_log2;  // log2(e) / sqrt(d_head)
};

// kRagged: seq_len % 128 != 0 (tail keys masked); a separate instantiation because the masking
// code costs 1.6 % at the headline even when never taken (profiles/r01_sweep9_ragged_ab.json).
// kPair: generation 7 (the kernel is then launched with cluster dimension 2).
template <bool kBF16, bool kDebug, bool kRagged, bool kPair>
__device__ __forceinline__ void fa_fwd_body(const CUtensorMap& tm_q, const CUtensorMap& tm_k,
                                            const CUtensorMap& tm_v, const CUtensorMap& tm_o,
                                            const FwdParams& prm, const FwdDebug& dbg) {
    constexpr bool kLdSplit = !kRagged && (FA_LD_SPLIT != 0);
    constexpr int kStages = kPair ? 2 * kKVStages : kKVStages;  // K/V ring slots ...
    constexpr int kSlotBytes = kPair ? kTileBytes / 2 : kTileBytes;  // ... of this size
    constexpr int kKHalfBytes = kPair ? kHalfBytes / 2 : kHalfBytes; // K: bytes per 64-d-column box
    constexpr uint32_t kArrivals = kPair ? 8u : 4u;  // softmax warps arriving on one barrier
    constexpr int kRowsPerTile = (kPair ? 2 : 1) * kQStages * kBlockM;
    auto col_s = [](int) -> uint32_t { return tmem_col_s(); };  // one S accumulator for both Q tiles

    // 1024-byte alignment (128B-swizzle atoms) is requested from the toolchain, which makes every
    // shared-memory address below a link-time constant instead of a live register.
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = smem_u32(smem_raw);
    uint8_t* smem_gen = smem_raw;
    if ((smem_base & 1023u) != 0u) {
        if (threadIdx.x == 0) printf("[fa] dynamic shared memory is not 1024-byte aligned\n");
        __trap();
    }

#if FA_UNIFORM_WARP
    // broadcast from lane 0: ptxas then knows the warp index (and every role branch on it) is
    // warp-uniform and keeps the MMA warp's counters and descriptors in uniform registers
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
#else
    const int warp = threadIdx.x >> 5;
#endif
    const int lane = threadIdx.x & 31;
    const int wg = warp >> 2;
    const uint32_t rank = kPair ? cluster_ctarank() : 0u;  // CTA inside its pair
    const bool is_leader = rank == 0;                       // the CTA whose MMA warp issues
    const int cta_lin = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int n_cta = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;

    // Barriers.  Every CTA holds a full set at the same offsets.  In a pair:
    //   * q_full / kv_full / p_full / p_last / o_free / s_free are used in the LEADER's copy only
    //     (TMA of both CTAs credits its bytes there; softmax warps of the peer arrive remotely);
    //   * q_empty / kv_empty / s_full / pv_done are signalled in BOTH copies by one multicast
    //     tcgen05.commit and waited on locally.
    const uint32_t bar0 = smem_base + kSmemBar;
    auto q_full = [&](int s) { return bar0 + 8u * s; };
    auto q_empty = [&](int s) { return bar0 + 8u * (2 + s); };
    auto kv_full = [&](int i) { return bar0 + 8u * (4 + i); };
    auto kv_empty = [&](int i) { return bar0 + 8u * (4 + kMaxStages + i); };
    auto s_full = [&](int s) { return bar0 + 8u * (4 + 2 * kMaxStages + s); };
    auto p_full = [&](int s) { return bar0 + 8u * (6 + 2 * kMaxStages + s); };
    auto p_last = [&](int s) { return bar0 + 8u * (8 + 2 * kMaxStages + s); };
    auto o_free = [&](int s) { return bar0 + 8u * (12 + 2 * kMaxStages + s); };
    const uint32_t s_free = bar0 + 8u * (14 + 2 * kMaxStages);                   // generation 6/7
    auto pv_done = [&](int s) { return bar0 + 8u * (15 + 2 * kMaxStages + s); };  // generation 6/7
    const uint32_t tmem_ptr_smem = smem_base + kSmemTmemPtr;

    auto wait = [&](uint32_t bar, uint32_t parity, int tag) { mbar_wait(bar, parity, tag); };
    auto arrive_leader = [&](uint32_t bar) {  // one arrival on the leader CTA's copy of `bar`
        if constexpr (kPair) mbar_arrive_cluster(mapa_shared(bar, 0));
        else mbar_arrive(bar);
    };
    auto commit = [&](uint32_t bar) {  // all earlier MMAs retired -> arrive (pair: in both CTAs)
        if constexpr (kPair) umma_commit_2cta(bar, 3);
        else umma_commit(bar);
    };

    const int n_blocks = prm.n_kv_blocks;
    const uint32_t level = kDebug ? dbg.level : 4u;
    // cycle trace (tools/gpu_trace2.py): the debug instantiation at level >= 5, or -- in -DFA_TRACE=1 builds -- the
    // production instantiation whenever a dump buffer is passed (fa_fwd_debug with level 40)
    constexpr bool kTr = kDebug || (FA_TRACE != 0);
    const bool tracing = kTr && (kDebug ? level >= 5 : true) && dbg.dump != nullptr && blockIdx.x == 0;
    // work tiles of this CTA (pair): cta_lin, cta_lin + n_cta, ...  (q-group index fastest so the
    // CTAs running at the same time share the K/V of a few (batch, head) pairs in L2)
    const int tile_end = (kDebug && level < 4) ? min(prm.n_tiles, cta_lin + 1) : prm.n_tiles;
    struct TileCoord {
        int q_row0, head, batch;
    };
    auto coord_of = [&](int tile) {
        TileCoord tc;
        const int qgroup = tile % prm.n_q_groups;
        const int bh = tile / prm.n_q_groups;
        tc.q_row0 = qgroup * kRowsPerTile + (int)rank * (kQStages * kBlockM);
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
            for (int s = 0; s < kQStages; ++s) {
                mbar_init(q_full(s), 1);
                mbar_init(q_empty(s), 1);
                mbar_init(s_full(s), 1);
                mbar_init(p_full(s), kArrivals);  // one elected arrive per softmax warp
                mbar_init(p_last(s), kArrivals);
                mbar_init(o_free(s), kArrivals);
                mbar_init(pv_done(s), 1);
            }
            mbar_init(s_free, kArrivals);
            for (int i = 0; i < kStages; ++i) {
                mbar_init(kv_full(i), 1);
                mbar_init(kv_empty(i), 1);
            }
            fence_mbar_init();
        }
        __syncwarp();
        if constexpr (kPair) {
            tmem_alloc_2cta(tmem_ptr_smem, kTmemCols);
            tmem_relinquish_2cta();
        } else {
            tmem_alloc(tmem_ptr_smem, kTmemCols);
            tmem_relinquish();
        }
    } else if (warp == 9 && lane == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_k);
        tma_prefetch_desc(&tm_v);
        tma_prefetch_desc(&tm_o);
    }
    tc_fence_before();
    if constexpr (kPair) cluster_sync();  // the peer's barriers exist before anything targets them
    __syncthreads();  // (pairs: orders the TMEM-address word for compute-sanitizer's racecheck, which does not
                      // model barrier.cluster; one CTA barrier per launch)
    tc_fence_after();
    // All 512 TMEM columns are allocated by the only CTA on this SM, so the base address is 0.
    // Using the literal keeps every tcgen05 operand warp-uniform (no R2UR per MMA).
    // (checked in the debug instantiation, which tests/test_kernel_gpu.py runs in every mapping: in a CTA pair
    // compute-sanitizer's racecheck cannot order tcgen05.alloc's write of this word against a read)
    if constexpr (kDebug || !kPair) {
        if (*reinterpret_cast<volatile uint32_t*>(smem_gen + kSmemTmemPtr) != 0u) {
            if (threadIdx.x == 0) printf("[fa] unexpected TMEM base address\n");
            __trap();
        }
    }
    constexpr uint32_t tmem_base = 0u;

    if (wg == 2) {
        setmaxnreg_dec<FA_REGS_CTRL>();
        if (warp == 9) {
            // ================================ TMA producer ================================
            // Convergent warp, one elected lane issues the copies (same reason as the MMA warp).
            // In a pair every CTA loads its own Q tiles, keys [64 rank, 64 rank + 64) of each K
            // block and d columns [64 rank, 64 rank + 64) of each V block; all bytes are credited to
            // the leader's full barrier, on which only the leader posts the expected total.
            int item = 0;  // K/V ring item counter (runs across tiles): K0, V0, K1, V1, ...
            int it = 0;    // local tile counter
            for (int tile = cta_lin; tile < tile_end; tile += n_cta, ++it) {
                const TileCoord tc = coord_of(tile);
                auto load_box = [&](uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int row0) {
                    if constexpr (kPair)
                        tma_load_4d_2cta(dst, map, mapa_shared(bar, 0), c0, tc.head, row0, tc.batch);
                    else
                        tma_load_4d(dst, map, bar, c0, tc.head, row0, tc.batch);
                };
                auto load_q = [&](int s) {
                    // Q_s smem is free once the previous tile's last S_s MMA retired
                    wait(q_empty(s), (uint32_t)((it & 1) ^ 1), 110 + s);
                    if (elect_one()) {
                        const uint32_t dst = smem_base + kSmemQ + s * kTileBytes;
                        const int row0 = tc.q_row0 + s * kBlockM;
                        if (is_leader) mbar_arrive_expect_tx(q_full(s), (kPair ? 2 : 1) * kTileBytes);
                        load_box(dst, &tm_q, q_full(s), 0, row0);
                        load_box(dst + kHalfBytes, &tm_q, q_full(s), 64, row0);
                    }
                    __syncwarp();
                };
                auto load_kv = [&](bool is_k, int blk) {
                    // K blocks cycle through the first half of the slots, V blocks through the second
                    // (item = 2 * block + is_v, running across tiles): the two MMA-issuing warps each own a ring
                    constexpr int kH = kStages / 2;
                    const int slot = (is_k ? 0 : kH) + (item >> 1) % kH;
                    const uint32_t use = (uint32_t)((item >> 1) / kH);
                    wait(kv_empty(slot), (use & 1u) ^ 1u, 100 + slot);
                    if (kDebug && level >= 7 && use > 0) {
                        // ablation: keep the barrier protocol, skip the copy (the slot keeps old data)
                        if (is_leader && elect_one()) mbar_arrive(kv_full(slot));
                    } else if (elect_one()) {
                        const uint32_t dst = smem_base + kSmemKV + slot * kSlotBytes;
                        if (is_leader) mbar_arrive_expect_tx(kv_full(slot), kTileBytes);
                        if (kPair && !is_k) {  // V: one box, this CTA's 64 d columns of all 128 keys
                            load_box(dst, &tm_v, kv_full(slot), 64 * (int)rank, blk * kBlockN);
                        } else {  // K (pair: this CTA's 64 keys; tm_k then has a 64-row box), or K / V whole
                            const CUtensorMap* map = is_k ? &tm_k : &tm_v;
                            const int row0 = blk * kBlockN + (kPair ? 64 * (int)rank : 0);
                            load_box(dst, map, kv_full(slot), 0, row0);
                            load_box(dst + kKHalfBytes, map, kv_full(slot), 64, row0);
                        }
                    }
                    __syncwarp();
                    ++item;
                };
                if (level >= 2) {
                    load_q(0);
                    load_kv(true, 0);
                }
                if (level >= 3) load_q(1);
                if (level >= 4) {
                    load_kv(false, 0);
                    if constexpr (FA_Q_PREFETCH != 0) {
                        // next work tile of this CTA: its Q rows (2 x 32 KiB) come from HBM exactly once; asking for
                        // them in L2 now takes the miss latency out of the tile boundary (Q_s is reloaded between the
                        // last S_s of this tile and the first of the next)
                        if (tile + n_cta < tile_end && elect_one()) {
                            const TileCoord nx = coord_of(tile + n_cta);
#pragma unroll
                            for (int s = 0; s < kQStages; ++s) {
                                tma_prefetch_l2_4d(&tm_q, 0, nx.head, nx.q_row0 + s * kBlockM, nx.batch);
                                tma_prefetch_l2_4d(&tm_q, 64, nx.head, nx.q_row0 + s * kBlockM, nx.batch);
                            }
                            if constexpr (FA_Q_PREFETCH >= 2) {
                                // ... and its first two K / V blocks when it belongs to another (batch, head): the
                                // first CTA to get there would otherwise take the HBM latency inside the boundary
                                if (nx.head != tc.head || nx.batch != tc.batch) {
                                    for (int j = 0; j < 2 && j < n_blocks; ++j) {
                                        const int krow = j * kBlockN + (kPair ? 64 * (int)rank : 0);
                                        tma_prefetch_l2_4d(&tm_k, 0, nx.head, krow, nx.batch);
                                        tma_prefetch_l2_4d(&tm_k, 64, nx.head, krow, nx.batch);
                                        if constexpr (kPair) {
                                            tma_prefetch_l2_4d(&tm_v, 64 * (int)rank, nx.head, j * kBlockN, nx.batch);
                                        } else {
                                            tma_prefetch_l2_4d(&tm_v, 0, nx.head, j * kBlockN, nx.batch);
                                            tma_prefetch_l2_4d(&tm_v, 64, nx.head, j * kBlockN, nx.batch);
                                        }
                                    }
                                }
                            }
                        }
                        __syncwarp();
                    }
                    for (int j = 1; j < n_blocks; ++j) {
                        load_kv(true, j);
                        load_kv(false, j);
                    }
                }
            }
        } else if (warp == 8 && is_leader) {
            // ================================ MMA issuer ==================================
            // The whole warp runs this loop CONVERGENTLY (all lanes poll the barriers) and one
            // elected lane issues the tcgen05 instructions.  Issuing from a `lane == 0` divergent
            // branch made ptxas wrap every UTCHMMA in an ELECT/BRA.U.ANY serialisation loop with
            // three R2URs: ~90 cycles per MMA, more than the 64 cycles the MMA takes to execute
            // (profiles/r01_v4_trace_notes.md).
            constexpr int kM = kPair ? 2 * kBlockM : kBlockM;
            constexpr uint32_t idesc_qk = umma_idesc_f16(kBF16, kM, kBlockN, false);
            // Q/K tiles: K-major, 8-row x 128 B swizzle atoms 1024 B apart (SBO); LBO unused.
            // V tiles: MN-major; next 64-wide d chunk 16 KiB away (LBO), next 8 kv rows 1 KiB
            // (SBO); one k-step = 16 kv rows = 2 KiB.  P (A operand in TMEM): 8 columns/k-step.
            // Pair: the B descriptors cover this CTA's half (64 keys of K / 64 d columns of V);
            // the peer CTA's tensor core reads the same offsets of its own shared memory.
            auto k_desc = [&](int slot) {
                return umma_smem_desc_sw128(smem_base + kSmemKV + slot * kSlotBytes, 16, 1024);
            };
            auto issue_qk_b = [&](int s, uint64_t b0) {
                const uint64_t a0 = umma_smem_desc_sw128(smem_base + kSmemQ + s * kTileBytes, 16, 1024);
                if constexpr (kPair && FA_QK_ONE_ASM) {
                    umma_ss_2cta_k8<(kHalfBytes >> 4), (kKHalfBytes >> 4)>(tmem_base + col_s(s), a0, b0, idesc_qk);
                    return;
                }
#pragma unroll
                for (int k = 0; k < kHeadDim / 16; ++k) {
                    const uint32_t a_off = ((k >> 2) * kHalfBytes + (k & 3) * 32) >> 4;
                    const uint32_t b_off = ((k >> 2) * kKHalfBytes + (k & 3) * 32) >> 4;
                    if constexpr (kPair)
                        umma_ss_2cta(tmem_base + col_s(s), a0 + a_off, b0 + b_off, idesc_qk, k > 0);
                    else
                        umma_ss(tmem_base + col_s(s), a0 + a_off, b0 + b_off, idesc_qk, k > 0);
                }
            };

            if constexpr (kDebug) {
                if (level == 2) {  // raw smem images of Q_0 and K_0 as TMA wrote them
                    wait(kv_full(0), 0, 200);
                    wait(q_full(0), 0, 210);
                    if (lane == 0 && dbg.dump != nullptr && blockIdx.x == 0) {
                        const uint32_t* qs = reinterpret_cast<const uint32_t*>(smem_gen + kSmemQ);
                        const uint32_t* ks = reinterpret_cast<const uint32_t*>(smem_gen + kSmemKV);
                        uint32_t* out = reinterpret_cast<uint32_t*>(dbg.dump);
                        for (int i = 0; i < kTileBytes / 4; ++i) out[i] = qs[i];
                        for (int i = 0; i < kTileBytes / 4; ++i) out[kTileBytes / 4 + i] = ks[i];
                    }
                }
            }
            {
                // ------------- generation 14: this warp streams the QK^T groups only -------------
                //   S_0(0) S_1(0) S_0(1) S_1(1) ...   each into the shared accumulator as soon as the previous S was
                //   read out (`s_free`) and K_j landed; the PV groups are issued by warp 10.
                constexpr int kH = kStages / 2;
                int kb = 0;      // K blocks consumed so far (all tiles): ring slot / parity
                uint32_t u = 0;  // S accumulators issued so far (all tiles): s_free parity
                int it = 0;
                for (int tile = cta_lin; level >= 3 && tile < tile_end; tile += n_cta, ++it) {
                    const int nb = (level >= 4) ? n_blocks : 1;
                    for (int jj = 0; jj < nb; ++jj, ++kb) {
                        const int slot = kb % kH;
                        const uint64_t kd = k_desc(slot);
                        wait(kv_full(slot), (uint32_t)((kb / kH) & 1), 200);
#pragma unroll
                        for (int s = 0; s < kQStages; ++s) {
                            uint32_t* tr = nullptr;
                            if constexpr (kTr) {
                                if (tracing && it == kTraceTile && jj < 32 && lane == 0)
                                    tr = reinterpret_cast<uint32_t*>(dbg.dump) + kTraceBase + 768 + (jj * 2 + s) * 4;
                                if (tr) tr[0] = clk32();
                            }
                            if (jj == 0) wait(q_full(s), (uint32_t)(it & 1), 210 + s);
                            if (u > 0) wait(s_free, (u - 1u) & 1u, 270 + s);
                            if constexpr (kTr) {
                                if (tr) tr[1] = clk32();
                            }
                            tc_fence_after();
                            if (elect_one()) {
                                issue_qk_b(s, kd);
                                commit(s_full(s));
                                if (jj + 1 == n_blocks) commit(q_empty(s));  // last use of Q_s
                                if (s == 1) commit(kv_empty(slot));          // both tiles used K_jj
                            }
                            __syncwarp();
                            if constexpr (kTr) {
                                if (tr) tr[2] = clk32();
                            }
                            ++u;
                        }
                    }
                }
            }
        } else if (warp == 10 && is_leader) {
            // ======================= MMA issuer 2 (generation 14): the PV groups =======================
            //   PV_0(0) PV_1(0) PV_0(1) PV_1(1) ...   each as soon as P_s(j) is stored (96 + 32 keys) and V_j landed
            constexpr int kM = kPair ? 2 * kBlockM : kBlockM;
            constexpr uint32_t idesc_pv = umma_idesc_f16(kBF16, kM, kHeadDim, true);
            constexpr int kH = kStages / 2;
            int vb = 0;       // V blocks consumed so far (all tiles)
            uint32_t g0 = 0;  // KV blocks of earlier tiles: parity base of the p / pv barriers
            int it = 0;
            for (int tile = cta_lin; level >= 4 && tile < tile_end; tile += n_cta, ++it) {
                for (int j = 0; j < n_blocks; ++j, ++vb) {
                    const int slot = kH + vb % kH;
                    const uint64_t vd =
                        umma_smem_desc_sw128(smem_base + kSmemKV + slot * kSlotBytes, kHalfBytes, 1024);
                    const uint32_t par = (g0 + (uint32_t)j) & 1u;
                    wait(kv_full(slot), (uint32_t)((vb / kH) & 1), 220);
#pragma unroll
                    for (int s = 0; s < kQStages; ++s) {
                        uint32_t* tr = nullptr;
                        if constexpr (kTr) {
                            if (tracing && it == kTraceTile && j < 32 && lane == 0)
                                tr = reinterpret_cast<uint32_t*>(dbg.dump) + kTraceBase + 512 + (j * 2 + s) * 4;
                            if (tr) tr[0] = clk32();
                        }
                        wait(p_full(s), par, 230 + s);  // P_s(j) stored (first 96 keys), O_s rescaled
                        if (j == 0)  // previous tile's epilogue has read O_s out of TMEM
                            wait(o_free(s), (uint32_t)((it & 1) ^ 1), 260 + s);
                        tc_fence_after();
                        auto pv = [&](int k_begin, int k_end) {
                            if constexpr (kPair && kSplitP && FA_QK_ONE_ASM) {  // one asm statement per part
                                if (k_begin == 0)
                                    umma_ts_2cta_k0to5(tmem_base + tmem_col_o(s), tmem_base + tmem_col_p(s), vd, idesc_pv,
                                                       j > 0 ? 1u : 0u);
                                else
                                    umma_ts_2cta_k6to7(tmem_base + tmem_col_o(s), tmem_base + tmem_col_p(s), vd, idesc_pv);
                                return;
                            }
#pragma unroll
                            for (int k = k_begin; k < k_end; ++k) {
                                const uint32_t acc = (j > 0 || k > 0) ? 1u : 0u;
                                if constexpr (kPair)
                                    umma_ts_2cta(tmem_base + tmem_col_o(s), tmem_base + tmem_col_p(s) + k * 8,
                                                 vd + ((k * 2048) >> 4), idesc_pv, acc);
                                else
                                    umma_ts(tmem_base + tmem_col_o(s), tmem_base + tmem_col_p(s) + k * 8,
                                            vd + ((k * 2048) >> 4), idesc_pv, acc);
                            }
                        };
                        if constexpr (kTr) {
                            if (tr) tr[1] = clk32();
                        }
                        if (elect_one()) pv(0, kSplitP ? 6 : 8);
                        __syncwarp();
                        if constexpr (kSplitP) {
                            wait(p_last(s), par, 250 + s);  // last 32 keys of P_s(j)
                            tc_fence_after();
                        }
                        if constexpr (kTr) {
                            if (tr) tr[2] = clk32();
                        }
                        if (elect_one()) {
                            if constexpr (kSplitP) pv(6, 8);
                            commit(pv_done(s));
                            if (s == 1) commit(kv_empty(slot));
                        }
                        __syncwarp();
                        if constexpr (kTr) {
                            if (tr) tr[3] = clk32();
                        }
                    }
                }
                g0 += (uint32_t)n_blocks;
            }
        }
        __syncwarp();
    } else if (wg < 2) {
        // ==================================== softmax =====================================
        setmaxnreg_inc<FA_REGS_SOFTMAX>();
        const int s = wg;                    // Q tile handled by this warpgroup
        const int row = threadIdx.x & 127;   // row inside the tile == TMEM lane
        const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const uint32_t t_s = tmem_base + lane_sel + col_s(s);
        const uint32_t t_p = tmem_base + lane_sel + tmem_col_p(s);
        const uint32_t t_o = tmem_base + lane_sel + tmem_col_o(s);
        const float c = prm.scale_log2;
        const int kv_tail = prm.seq_len & (kBlockN - 1);  // valid keys in the last block (0 = all)
        uint32_t g = 0;  // KV blocks processed so far (all tiles)
        int it = 0;

        for (int tile = cta_lin; level >= 3 && tile < tile_end; tile += n_cta, ++it) {
            float m_run = -INFINITY;  // running (possibly stale) row max, raw S units
            float l_run = 0.f;        // running row sum of exp2
            const int n_iter = (level >= 4) ? n_blocks : 1;
            uint32_t sr[4][32];       // this thread's row of S(j): 4 fragments of 32 columns
            auto mask_tail = [&](int q_begin, int q_end) {
                // ragged tail (seq_len % 128 != 0, beyond the reference's contract): TMA zero-filled
                // the missing K/V rows; their scores are forced to -inf so that P = 0 exactly.
#pragma unroll
                for (int q = q_begin; q < q_end; ++q) {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (q * 32 + i >= kv_tail) sr[q][i] = 0xff800000u;
                }
            };
            for (int j = 0; j < n_iter; ++j, ++g) {
                wait(s_full(s), g & 1u, 300 + s);
                tc_fence_after();
                uint32_t* tr = nullptr;
                if constexpr (kTr) {
                    if (tracing && it == kTraceTile && j < 32 && (warp & 3) == 0 && lane == 0)
                        tr = reinterpret_cast<uint32_t*>(dbg.dump) + kTraceBase + (s * 32 + j) * 8;
                    if (tr) tr[0] = clk32();
                }
                float m_lo = -INFINITY;  // kLdSplit: row max of columns [0, 64)
                if (kDebug && level >= 9) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
#pragma unroll
                        for (int i = 0; i < 32; ++i) sr[q][i] = 0u;
                } else if constexpr (kLdSplit) {
                    tmem_ld_32x32b_x32(t_s, sr[0]);
                    tmem_ld_32x32b_x32(t_s + 32, sr[1]);
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
                // S is in registers: hand the shared accumulator to the other Q tile's next S
                tc_fence_before();
                __syncwarp();
                if constexpr (kLdSplit) {
                    // ptxas hoists the arrive (and with it the stall on the last tcgen05.ld) above the reduction of
                    // the first half; a never-true dependency on m_lo keeps it behind.  Never true: m_lo comes out
                    // of an FMNMX chain, whose only NaN result is the canonical 0x7fffffff (one input NaN -> the
                    // other input; both NaN -> canonical NaN), never this payload.  (A form that is correct for ANY
                    // m_lo -- two equivalent predicated arrives chosen by the same comparison -- was measured: -2.5 %
                    // pair, -8 % single, profiles/r02_g15_arrive_form_sweep.json; comments only, SASS unchanged.)
                    const uint32_t skew = (__float_as_uint(m_lo) == 0x7fc12345u) ? 8u : 0u;
                    if (lane == 0) arrive_leader(s_free + skew);
                } else {
                    if (lane == 0) arrive_leader(s_free);
                }
                
                if constexpr (kTr) {
                    if (tr) tr[1] = clk32();
                }
                if (kRagged && j + 1 == n_blocks && kv_tail != 0) {
                    mask_tail(0, 4);
                }

                if constexpr (kDebug) {
                    if (level < 5 && dbg.dump != nullptr && blockIdx.x == 0 && it == 0 && j == 0) {
                        for (int q = 0; q < 4; ++q)
                            for (int i = 0; i < 32; ++i)
                                dbg.dump[(s * 128 + row) * 128 + q * 32 + i] =
                                    __uint_as_float(sr[q][i]);
                    }
                    if (level == 3) break;
                }
                const bool no_math = kDebug && level >= 6 && level != 7;  // timing ablation
                float mx;
                if (no_math) {
                    mx = 0.f;
                } else if constexpr (kLdSplit) {
                    mx = fmaxf(m_lo, row_max_frags<2, 2>(sr));
                } else {
                    mx = row_max_128(sr);
                }
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
                        // O_s must be quiescent.  Generation 4b: S_s(j) was committed after
                        // PV_s(j-1); generation 6/7: wait for PV_s(j-1) explicitly.
                                wait(pv_done(s), (g - 1u) & 1u, 320 + s);
                        tc_fence_after();
                    
#pragma unroll 1
                        for (int q = 0; q < 8; ++q) {  // 16 columns at a time: S(j) stays in registers
                            uint32_t o[16];
                            tmem_ld_32x32b_x16(t_o + q * 16, o);
                            tmem_wait_ld();
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                            tmem_st_32x32b_x16(t_o + q * 16, o);
                        }
                    }
                }
                if constexpr (kTr) {
                    if (tr) tr[2] = clk32();
                }
                const float neg_mc = -m_run * c;
                const float2 c2 = make_float2(c, c);
                const float2 nm2 = make_float2(neg_mc, neg_mc);
                float2 sum_a = make_float2(0.f, 0.f), sum_b = make_float2(0.f, 0.f);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t pk[16];
                    if (no_math) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) pk[i] = 0x3c003c00u;
                    } else if (q == 3) {
                        exp_fragment<kBF16, kEmuPairsLast, FA_EXP_VARIANT>(sr[q], c2, nm2, sum_a, sum_b, pk);
                    } else {
                        exp_fragment<kBF16, kEmuPairs, FA_EXP_VARIANT>(sr[q], c2, nm2, sum_a, sum_b, pk);
                    }
                        // P_s(j) overwrites P_s(j-1): PV_s(j-1) must have read it (issued about one
                    // softmax fragment ago, so this rarely spins)
                    // (the first block of a tile overwrites the previous tile's last P_s too: generation 14's
                    // epilogue inside this warpgroup used to wait for that PV)
                    if (q == 0 && g > 0) {
                        if constexpr (kTr) {
                            if (tr) tr[5] = clk32();
                        }
                        wait(pv_done(s), (g - 1u) & 1u, 330 + s);
                        tc_fence_after();
                        if constexpr (kTr) {
                            if (tr) tr[6] = clk32();
                        }
                    }
                
                    tmem_st_32x32b_x16(t_p + q * 16, pk);
                    if (kSplitP && q == 2) {
                        tmem_wait_st();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) arrive_leader(p_full(s));
                        if constexpr (kTr) {
                            if (tr) tr[3] = clk32();
                        }
                    }
                }
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) arrive_leader(kSplitP ? p_last(s) : p_full(s));
                if constexpr (kTr) {
                    if (tr) tr[4] = clk32();
                }
                l_run = l_run * alpha + ((sum_a.x + sum_a.y) + (sum_b.x + sum_b.y));
            }

            // generation 15: hand the row sum to the epilogue warpgroup and go straight on to the next tile
            if (level >= 4) {
                if constexpr (kDebug) {
                    if (dbg.dump != nullptr && blockIdx.x == 0 && it == 0) {
                        dbg.dump[2 * 128 * 128 + s * 128 + row] = l_run;
                        dbg.dump[2 * 128 * 128 + 256 + s * 128 + row] = m_run;
                    }
                }
                if (it > 0) named_bar_sync(12 + s, 256);  // the previous tile's row sum was read
                reinterpret_cast<float*>(smem_gen + kSmemL)[s * kBlockM + row] = l_run;
                named_bar_arrive(10 + s, 256);
            }
        }
    } else {
        // =================================== epilogue (generation 15) =====================================
        // Own warpgroup (as in the ping-pong kernel): O_s / l -> 16 bit -> swizzled shared memory -> TMA store, for
        // Q tile 0 then Q tile 1 of every work tile.  With the epilogue inside the softmax warpgroups a tile boundary
        // cost 1.5-3 us (last PV -> read O -> convert -> store -> only then the next tile's first softmax, with the
        // tensor pipe idle): ~4 % at seq_len 4096, where a tile is 32 KV blocks (profiles/r02_g15_notes.md).
        setmaxnreg_dec<FA_REGS_EPI>();
        const int row = threadIdx.x & 127;
        const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const float* sm_l = reinterpret_cast<const float*>(smem_gen + kSmemL);
        uint32_t g_last = 0;  // stream index of the current tile's last KV block (parity of pv_done)
        int it = 0;
        for (int tile = cta_lin; level >= 4 && tile < tile_end; tile += n_cta, ++it) {
            g_last += (uint32_t)n_blocks;
            const TileCoord tc = coord_of(tile);
#pragma unroll 1
            for (int s = 0; s < kQStages; ++s) {
                named_bar_sync(10 + s, 256);  // softmax warpgroup s has finished the tile: its row sums are written
                const float inv_l = 1.0f / sm_l[s * kBlockM + row];
                named_bar_arrive(12 + s, 256);
                // last PV_s of the tile retired.  No parity aliasing: warpgroup s waited for PV_s(last - 1) before it
                // stored its last P, and PV_s of the next tile waits for o_free(s) below.
                wait(pv_done(s), (g_last - 1u) & 1u, 310 + s);
                tc_fence_after();
                const uint32_t t_o = tmem_base + lane_sel + tmem_col_o(s);
#pragma unroll
                for (int q = 0; q < 4; ++q) {  // 32 columns at a time
                    const int h = q >> 1;      // TMA box / staging buffer
                    uint32_t o[32];
                    tmem_ld_32x32b_x32(t_o + 32 * q, o);
                    tmem_wait_ld();
                    if (q == 3) {  // O_s is in registers: the next tile's first PV_s may overwrite it
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) arrive_leader(o_free(s));
                    }
                    if ((q & 1) == 0) {
                        // staging buffer h: the store issued two groups ago must have read it (bulk groups of the
                        // issuing thread complete in order: at most one newer group may still be pending)
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
                            tma_store_4d(&tm_o, smem_base + kSmemStage + h * kHalfBytes, 64 * h, tc.head,
                                         tc.q_row0 + s * kBlockM, tc.batch);
                            tma_store_commit();
                        }
                    }
                }
            }
        }
        if (row == 0) tma_store_wait_read<0>();  // shared memory must outlive the last store's read
    }

    // ------------------------------------ teardown ---------------------------------------
    // Pair: nobody may leave while the other CTA can still arrive on its barriers or while the
    // leader's MMAs still write its tensor memory (all of them have retired once both epilogues ran).
    tc_fence_before();
    if constexpr (kPair) cluster_sync();
    else __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        if constexpr (kPair) tmem_dealloc_2cta(tmem_base, kTmemCols);
        else tmem_dealloc(tmem_base, kTmemCols);
    }
}

template <bool kBF16, bool kDebug, bool kRagged>
__global__ void __launch_bounds__(kNumThreads, 1)
fa_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
              const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o,
              const FwdParams prm, const FwdDebug dbg) {
    fa_fwd_body<kBF16, kDebug, kRagged, false>(tm_q, tm_k, tm_v, tm_o, prm, dbg);
}

// Generation 7: clusters of two CTAs (tcgen05 cta_group::2).  `tm_k` must have a 64-row box.
template <bool kBF16, bool kDebug, bool kRagged>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNumThreads, 1)
fa_fwd_kernel_pair(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o,
                   const FwdParams prm, const FwdDebug dbg) {
    fa_fwd_body<kBF16, kDebug, kRagged, true>(tm_q, tm_k, tm_v, tm_o, prm, dbg);
}

}  // namespace fa

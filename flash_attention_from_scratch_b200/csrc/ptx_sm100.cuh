// Inline-PTX wrappers for the sm_100a features the attention forward kernel uses:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld / st / fence)
// and a few math helpers.
//
// This replaces the role of the reference's Ampere wrappers
// (/root/reference/src/include/ptx_functions.cuh:9-77: cp.async, ldmatrix, mma.sync) --
// none of those instructions are used here.
#pragma once
#include <cstdint>
#include <cuda.h>  // CUtensorMap (type only; no libcuda symbol is referenced)

namespace fa {

#ifndef FA_HANG_GUARD
#define FA_HANG_GUARD 0
#endif
#ifndef FA_WAIT_HINT
#define FA_WAIT_HINT 0    // > 0: suspend-time hint (ns) passed to every mbarrier.try_wait (power experiment)
#endif
#ifndef FA_WAIT_SLEEP
#define FA_WAIT_SLEEP 0   // > 0: nanosleep(ns) after every failed mbarrier.try_wait (power experiment)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
// The default (.acquire.cta) form is also what the CTA-pair kernel uses: an acquire at cluster scope
// makes ptxas emit CCTL.IVALL (L1 invalidate) after EVERY wait and MEMBAR.ALL.GPU before every remote
// arrive -- measured 2x slower end to end -- and it is not needed: everything that crosses the two CTAs
// travels through the async / tensor proxies (TMA complete_tx, tcgen05.commit, tensor memory) and is
// ordered by the tcgen05 fences, never through generic-proxy memory.
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Non-blocking test (try_wait may suspend the thread for a while when the phase is still running; this never does).
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ uint32_t lds_volatile_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.volatile.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_volatile_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.volatile.shared::cta.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// Blocks until the phase with the given parity has completed.  With FA_HANG_GUARD the spin is
// bounded and traps with a diagnostic instead of hanging the GPU (bring-up builds only).
#if FA_HANG_GUARD
// Host-mapped diagnostics ring (bring-up builds): [0] = record count, then 4 words per record.
__device__ uint32_t* g_fa_diag = nullptr;
__device__ __forceinline__ void diag_record(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    uint32_t* p = g_fa_diag;
    if (p == nullptr) return;
    const uint32_t i = atomicAdd(p, 1u);
    if (i < 60) {
        p[4 + 4 * i + 0] = a;
        p[4 + 4 * i + 1] = b;
        p[4 + 4 * i + 2] = c;
        p[4 + 4 * i + 3] = d;
    }
    __threadfence_system();
}
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag = 0) {
#if FA_HANG_GUARD
    for (uint32_t it = 0; it < (1u << 22); ++it) {
        if (mbar_try_wait(bar, parity)) return;
    }
    diag_record(0xDEAD0000u | (uint32_t)tag, parity, threadIdx.x, blockIdx.x);
    printf("[fa] mbarrier timeout: block %d thread %d tag %d parity %u\n", (int)blockIdx.x,
           (int)threadIdx.x, tag, parity);
    __trap();
#else
    (void)tag;
#if FA_WAIT_HINT
    // try_wait with an explicit suspend-time hint: the warp may sleep in hardware for up to that many ns
    // per attempt instead of the (short) default limit
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"((uint32_t)FA_WAIT_HINT)
            : "memory");
    } while (!ok);
#else
    while (!mbar_try_wait(bar, parity)) {
#if FA_WAIT_SLEEP
        __nanosleep(FA_WAIT_SLEEP);
#endif
    }
#endif
#endif
}

// Waits for two barriers at once: both try_waits are in flight together, so two phases that completed long
// ago cost one ~90-cycle round trip instead of two (it matters for the MMA-issuing warps, whose own
// instruction time per KV block is what paces the tensor pipe).
__device__ __forceinline__ void mbar_wait2(uint32_t bar_a, uint32_t par_a, uint32_t bar_b, uint32_t par_b,
                                           int tag = 0) {
#if FA_HANG_GUARD
    mbar_wait(bar_a, par_a, tag);
    mbar_wait(bar_b, par_b, tag + 1);
#else
    (void)tag;
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p, q;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 q, [%3], %4;\n\t"
            "and.pred p, p, q;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar_a), "r"(par_a), "r"(bar_b), "r"(par_b)
            : "memory");
    } while (!ok);
#endif
}

// ----------------------------------------------------------------------------------------------
// TMA
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
// 4-D tiled load global -> shared, completion signalled on an mbarrier (complete_tx::bytes).
__device__ __forceinline__ void tma_load_4d(uint32_t dst_smem, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst_smem), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// Same with an L2 cache-policy hint.
__device__ __forceinline__ void tma_load_4d_hint(uint32_t dst_smem, const CUtensorMap* map,
                                                 uint32_t bar, int c0, int c1, int c2, int c3,
                                                 uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        ".L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
        ::"r"(dst_smem), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(policy)
        : "memory");
}
// 4-D tiled prefetch global -> L2 only (no shared-memory destination, no completion to wait for).
__device__ __forceinline__ void tma_prefetch_l2_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
// 4-D tiled store shared -> global (bulk-group completion).
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src_smem, int c0,
                                             int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
        ::"l"(map), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// Make generic-proxy shared-memory writes visible to the async proxy (TMA store source).
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// ----------------------------------------------------------------------------------------------
// named barriers (sub-CTA sync)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ----------------------------------------------------------------------------------------------
// register re-balancing between warpgroups
// ----------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

// ----------------------------------------------------------------------------------------------
// tcgen05: tensor memory + 5th-gen tensor core MMA
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_wait_st() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]      (both operands from shared memory)
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                        uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]           (A operand read from tensor memory)
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                        uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive (count 1) on an mbarrier once every tcgen05 op issued so far by this thread retired.
// Implies tcgen05.fence::before_thread_sync.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     bar)
                 : "memory");
}

// ----------------------------------------------------------------------------------------------
// CTA pairs (thread-block cluster of 2, tcgen05 .cta_group::2): one thread of the even CTA issues
// an MMA with M = 256 whose accumulator rows [0,128) live in its own tensor memory and rows
// [128,256) in the odd CTA's; A comes from each CTA's own shared / tensor memory, B is split in
// halves along N between the two CTAs' shared memories (same offsets in both).
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_sync() {
    cluster_arrive();
    cluster_wait();
}
// shared::cta address of this CTA -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// same with release semantics at cluster scope (orders this thread's earlier generic-proxy writes for
// an observer in the other CTA; ptxas emits a MEMBAR for it)
__device__ __forceinline__ void mbar_arrive_cluster_release(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void umma_ss_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                             uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_ts_2cta(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                             uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// One QK^T group (8 MMAs, K = 128) as ONE asm statement: the per-k descriptors are derived from the two base
// descriptors inside the PTX, so the compiler has 4 operands to place in uniform registers instead of 8 x 5 (with
// eight separate statements ptxas ran out of uniform registers in the CTA-pair kernel and moved 26 values through
// vector registers -- R2UR -- between the barrier wait and the first UTCHMMA of every group).
// kAHalf / kBHalf: descriptor distance (bytes >> 4) between the two 64-column halves of the A / B tile.
template <int kAHalf, int kBHalf>
__device__ __forceinline__ void umma_ss_2cta_k8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 a, b;\n\t"
        "setp.ne.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "setp.eq.b32 p, 0, 0;\n\t"
        "add.s64 a, %1, 2;\n\tadd.s64 b, %2, 2;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a, b, %3, p;\n\t"
        "add.s64 a, %1, 4;\n\tadd.s64 b, %2, 4;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a, b, %3, p;\n\t"
        "add.s64 a, %1, 6;\n\tadd.s64 b, %2, 6;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a, b, %3, p;\n\t"
        "add.s64 a, %1, %4;\n\tadd.s64 b, %2, %5;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a, b, %3, p;\n\t"
        "add.s64 a, %1, %6;\n\tadd.s64 b, %2, %7;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a, b, %3, p;\n\t"
        "add.s64 a, %1, %8;\n\tadd.s64 b, %2, %9;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a, b, %3, p;\n\t"
        "add.s64 a, %1, %10;\n\tadd.s64 b, %2, %11;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a, b, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "n"(kAHalf), "n"(kBHalf), "n"(kAHalf + 2),
        "n"(kBHalf + 2), "n"(kAHalf + 4), "n"(kBHalf + 4), "n"(kAHalf + 6), "n"(kBHalf + 6)
        : "memory");
}
// The PV group of one KV block in two asm statements (keys [0,96) and [96,128): the second waits for the last
// part of P): A = P in tensor memory (8 columns per k-step), B = V in shared memory (MN-major: 2 KiB = 128
// descriptor units per k-step).  `accumulate` applies to the very first MMA only (0: O is overwritten).
__device__ __forceinline__ void umma_ts_2cta_k0to5(uint32_t d_tmem, uint32_t p_tmem, uint64_t b_desc, uint32_t idesc,
                                                   uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 b;\n\t.reg .b32 a;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "setp.eq.b32 p, 0, 0;\n\t"
        "add.u32 a, %1, 8;\n\tadd.s64 b, %2, 128;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [a], b, %3, p;\n\t"
        "add.u32 a, %1, 16;\n\tadd.s64 b, %2, 256;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [a], b, %3, p;\n\t"
        "add.u32 a, %1, 24;\n\tadd.s64 b, %2, 384;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [a], b, %3, p;\n\t"
        "add.u32 a, %1, 32;\n\tadd.s64 b, %2, 512;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [a], b, %3, p;\n\t"
        "add.u32 a, %1, 40;\n\tadd.s64 b, %2, 640;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [a], b, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(p_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_ts_2cta_k6to7(uint32_t d_tmem, uint32_t p_tmem, uint64_t b_desc, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 b;\n\t.reg .b32 a;\n\t"
        "setp.eq.b32 p, 0, 0;\n\t"
        "add.u32 a, %1, 48;\n\tadd.s64 b, %2, 768;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [a], b, %3, p;\n\t"
        "add.u32 a, %1, 56;\n\tadd.s64 b, %2, 896;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [a], b, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(p_tmem), "l"(b_desc), "r"(idesc)
        : "memory");
}
// Arrive on the mbarrier at this shared-memory offset in every CTA of `cta_mask` once all
// tcgen05 ops issued so far by this thread (for the CTA pair) have retired.
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar, uint16_t cta_mask) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
        "[%0], %1;" ::"r"(bar), "h"(cta_mask)
        : "memory");
}
// TMA load whose completion bytes are credited to the mbarrier at cluster address `bar_cluster`
// (normally the even CTA's barrier, so that the MMA issuer sees both CTAs' halves arrive).
__device__ __forceinline__ void tma_load_4d_2cta(uint32_t dst_smem, const CUtensorMap* map,
                                                 uint32_t bar_cluster, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst_smem), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// TMEM -> registers: each lane of the warp reads 32 consecutive 32-bit columns of "its" TMEM lane
// (warp w of a warpgroup owns TMEM lanes 32*(w%4) .. 32*(w%4)+31).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
        "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]),
        "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
        "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]),
        "r"(r[14]), "r"(r[15])
        : "memory");
}

// ----------------------------------------------------------------------------------------------
// descriptors
// ----------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor (64 bit) for tcgen05.mma operands, 128-byte swizzle.
//   [0,14)  start address >> 4        [16,30) leading-dim byte offset >> 4
//   [32,46) stride-dim byte offset >> 4    [46,48) descriptor version (1 on sm_100)
//   [61,64) layout type (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes,
                                                         uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// Instruction descriptor (32 bit) for kind::f16 with fp32 accumulation.
//   [4,6) D format (1 = f32)  [7,10) A format (0 = f16, 1 = bf16)  [10,13) B format
//   [15] A major (0 = K)      [16] B major (0 = K, 1 = MN)
//   [17,23) N >> 3            [24,29) M >> 4
__host__ __device__ constexpr uint32_t umma_idesc_f16(bool bf16, int M, int N, bool b_mn_major) {
    return (1u << 4) | ((bf16 ? 1u : 0u) << 7) | ((bf16 ? 1u : 0u) << 10) |
           ((b_mn_major ? 1u : 0u) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
           (static_cast<uint32_t>(M >> 4) << 24);
}

// ----------------------------------------------------------------------------------------------
// math
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// exp2 of two values on the FMA pipe (no MUFU): Cody-Waite split with the 2^23+2^22 magic-number
// trick (round-down add leaves floor(x) in the low mantissa bits), degree-3 minimax polynomial for
// 2^f on [0,1) (max rel. error 8.6e-5, coefficients fitted for this kernel: p(0) = 1 exactly),
// then the integer part is added straight into the exponent field.  Requires x <= 127; inputs
// below -127 are clamped (result ~ 2^-127 ~ 0).  Accuracy is far below the 2^-9 (bf16) / 2^-11
// (fp16) rounding P receives anyway.
__device__ __forceinline__ float2 ex2_emulated_x2(float2 x) {
    const float kMagic = 12582912.0f;  // 2^23 + 2^22
    x.x = fmaxf(x.x, -127.0f);
    x.y = fmaxf(x.y, -127.0f);
    const float2 t = __fadd2_rd(x, make_float2(kMagic, kMagic));        // low bits = floor(x)
    const float2 r = __fadd2_rn(t, make_float2(-kMagic, -kMagic));      // floor(x) as float
    const float2 f = __fadd2_rn(x, make_float2(-r.x, -r.y));            // fractional part in [0,1)
    float2 p = __ffma2_rn(make_float2(0.07706724f, 0.07706724f), f,
                          make_float2(0.22764498f, 0.22764498f));
    p = __ffma2_rn(p, f, make_float2(0.69511664f, 0.69511664f));
    p = __ffma2_rn(p, f, make_float2(1.0f, 1.0f));
    float2 out;
    out.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23));
    out.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23));
    return out;
}

// Packs two fp32 into one 32-bit register of two 16-bit floats: lo -> bits [0,16), hi -> [16,32).
template <bool kBF16>
__device__ __forceinline__ uint32_t pack_16x2(float lo, float hi) {
    uint32_t r;
    if constexpr (kBF16) {
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    } else {
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    }
    return r;
}

}  // namespace fa

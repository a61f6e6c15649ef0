// Tensor-pipe probe (development aid): how many cycles does one tcgen05.mma (kind::f16, K = 16)
// take on B200 depending on where the operands come from?
//
// The attention kernel's S = Q K^T tiles are "SS" MMAs (A and B from shared memory, 8 KiB per
// 128x128x16 MMA = 128 B/clk at the nominal 64 clk) while its PV tiles are "TS" (A = P from tensor
// memory).  The cycle trace of generation 6 showed the SS tiles taking ~100 clk per MMA, i.e. the
// shared-memory operand fetch, not the tensor core, paces QK^T.  This probe measures, on all SMs at
// once and with nothing else running:
//     mode 0  SS, one CTA            (M = 128)        mode 1  TS, one CTA
//     mode 2  SS, CTA pair           (M = 256; B split in halves between the two CTAs' smem)
//     mode 3  TS, CTA pair
// for N in {64, 128, 256}, and checks every mode's result against a CPU product so that the
// descriptor conventions (in particular the .cta_group::2 operand split) are pinned before the
// attention kernel relies on them.
//
// Second part (added after the tensor-pipe observer of tools/gpu_trace.py showed ~117 clk per QK^T MMA
// inside the attention kernel while P V MMAs ran at ~64): the same loops with an interference agent in
// warps 4..7, to find out WHAT slows the SS MMAs down in situ:
//     noise 0 none   1 softmax-like FFMA/MUFU register math   2 bulk copies global -> shared at full rate
//     3 bulk copies paced like the kernel's K/V stream (16 KiB per 700 clk)   4 tcgen05.ld of 128 columns
//     5 = 1 + 3 + 4 together
// `rnd` fills the operands with random normal-ish bf16 instead of small integers (data-dependent power).
// Mode bit 2 (TS only): B is MN-major like the kernel's V operand.
//
// build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -DFA_HANG_GUARD=1
//        -I flash_attention_from_scratch_b200/csrc -o tools/mma_probe tools/mma_probe.cu
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#include "ptx_sm100.cuh"

using namespace fa;

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                   \
        }                                                                              \
    } while (0)

__host__ __device__ inline int a_val(int r, int k) { return ((r * 3 + k * 5) % 7) - 3; }
__host__ __device__ inline int b_val(int n, int k) { return ((n * 5 + k * 7) % 5) - 2; }

__device__ __forceinline__ uint16_t bf16_bits(int v) {  // small integers are exact in bf16
    return static_cast<uint16_t>(__float_as_uint(static_cast<float>(v)) >> 16);
}

constexpr int kSmemA = 0;           // 128 rows x 128 k, two 16 KiB halves (k < 64 | k >= 64)
constexpr int kSmemB = 32 * 1024;   // up to 256 rows x 128 k, two halves of rows*128 B
constexpr int kSmemBar = 96 * 1024;
constexpr int kSmemScratch = 100 * 1024;  // 2 x 16 KiB landing zone of the interference bulk copies
constexpr int kSmemBytes = kSmemScratch + 32 * 1024;
constexpr uint32_t kColD = 0, kColA = 256;

struct ProbeOut {
    unsigned long long cycles;
    unsigned long long ns;
    unsigned long long first_group;  // one 8-MMA group + commit on an idle tensor pipe: issue -> barrier seen
    unsigned long long warm_group;   // the same group issued right behind another one (minus that one's 8 MMAs)
};

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
// random bf16 with |value| in [0.5, 4): random sign, 3 exponents, random mantissa
__device__ __forceinline__ uint16_t rnd_bf16(uint32_t idx) {
    const uint32_t h = hash32(idx * 2654435761u + 12345u);
    const uint32_t sign = (h >> 31) << 15, exp = (126u + (h >> 8) % 3u) << 7, man = h & 0x7fu;
    return static_cast<uint16_t>(sign | exp | man);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// kMode: bit 0 = A from tensor memory, bit 1 = CTA pair, bit 2 = B MN-major (the kernel's V operand)
template <int kMode>
__device__ __forceinline__ void probe_body(int N, int groups, ProbeOut* out, float* dout, int noise, int rnd,
                                           const uint8_t* gsrc, float* sink, int pattern) {
    constexpr bool kTS = (kMode & 1) != 0;
    constexpr bool kPair = (kMode & 2) != 0;
    constexpr bool kBT = (kMode & 4) != 0;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar = sbase + kSmemBar;
    const uint32_t tmem_ptr = sbase + kSmemBar + 16;
    const uint32_t bar_noise = bar + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = kPair ? cluster_ctarank() : 0u;
    const int rows_b = kPair ? N / 2 : N;  // B rows held by this CTA

    if (warp == 0) {
        if (lane == 0) {
            mbar_init(bar, 1);
            mbar_init(bar + 8, 1);   // interference bulk copies
            mbar_init(bar + 24, 1);  // pattern 6: per-group commits nobody waits for
            fence_mbar_init();
        }
        __syncwarp();
        if constexpr (kPair) {
            tmem_alloc_2cta(tmem_ptr, 512);
            tmem_relinquish_2cta();
        } else {
            tmem_alloc(tmem_ptr, 512);
            tmem_relinquish();
        }
    }
    // operands, K-major with the 128-byte swizzle TMA would produce
    for (int idx = threadIdx.x; idx < 128 * 128; idx += blockDim.x) {
        const int r = idx >> 7, k = idx & 127;
        const int h = k >> 6, c = (k & 63) >> 3, e = k & 7;
        const uint32_t off = h * (128 * 128) + r * 128 + ((c ^ (r & 7)) * 16) + e * 2;
        *reinterpret_cast<uint16_t*>(smem + kSmemA + off) =
            rnd ? rnd_bf16(idx + 7919u * rank) : bf16_bits(a_val((int)rank * 128 + r, k));
    }
    for (int idx = threadIdx.x; idx < rows_b * 128; idx += blockDim.x) {
        const int r = idx >> 7, k = idx & 127;
        const int h = k >> 6, c = (k & 63) >> 3, e = k & 7;
        const uint32_t off = h * (rows_b * 128) + r * 128 + ((c ^ (r & 7)) * 16) + e * 2;
        *reinterpret_cast<uint16_t*>(smem + kSmemB + off) =
            rnd ? rnd_bf16(idx + 104729u * (rank + 2)) : bf16_bits(b_val((int)rank * rows_b + r, k));
    }
    if constexpr (kBT) {  // timing only: any 128 keys x (pair: 64, else 128) d columns will do
        for (int idx = threadIdx.x; idx < 16 * 1024; idx += blockDim.x)
            reinterpret_cast<uint16_t*>(smem + kSmemB)[idx] = rnd ? rnd_bf16(idx + 31u) : bf16_bits((idx % 5) - 2);
    }
    if (pattern != 0) {  // a V-like MN-major operand behind the K-like one (timing only)
        for (int idx = threadIdx.x; idx < 16 * 1024; idx += blockDim.x)
            reinterpret_cast<uint16_t*>(smem + kSmemB + 32 * 1024)[idx] = rnd ? rnd_bf16(idx + 77u) : bf16_bits((idx % 5) - 2);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *reinterpret_cast<volatile uint32_t*>(smem + kSmemBar + 16);
    if ((kTS || pattern != 0) && warp < 4) {  // A as the MMA reads it from tensor memory: lane = row, column = k / 2
        const uint32_t t_a = tbase + (static_cast<uint32_t>(warp * 32) << 16) + kColA;
        const int r = (int)rank * 128 + warp * 32 + lane;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            uint32_t v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int k = (q * 32 + i) * 2;
                v[i] = rnd ? ((uint32_t)rnd_bf16(r * 128 + k) | ((uint32_t)rnd_bf16(r * 128 + k + 1) << 16))
                           : ((uint32_t)bf16_bits(a_val(r, k)) | ((uint32_t)bf16_bits(a_val(r, k + 1)) << 16));
            }
            tmem_st_32x32b_x32(t_a + q * 32, v);
        }
        tmem_wait_st();
    }
    tc_fence_before();
    if constexpr (kPair) cluster_sync();
    else __syncthreads();
    tc_fence_after();

    const uint32_t idesc = umma_idesc_f16(true, kPair ? 256 : 128, N, kBT);
    const uint64_t a0 = umma_smem_desc_sw128(sbase + kSmemA, 16, 1024);
    const uint64_t b0 = kBT ? umma_smem_desc_sw128(sbase + kSmemB, 16 * 1024, 1024)
                            : umma_smem_desc_sw128(sbase + kSmemB, 16, 1024);
    auto issue_group = [&](bool first) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t a_off = ((k >> 2) * (128 * 128) + (k & 3) * 32) >> 4;
            const uint32_t b_off = kBT ? (uint32_t)((k * 2048) >> 4)
                                       : (uint32_t)(((k >> 2) * (rows_b * 128) + (k & 3) * 32) >> 4);
            const uint32_t acc = (first && k == 0) ? 0u : 1u;
            if constexpr (kPair) {
                if constexpr (kTS) umma_ts_2cta(tbase + kColD, tbase + kColA + k * 8, b0 + b_off, idesc, acc);
                else umma_ss_2cta(tbase + kColD, a0 + a_off, b0 + b_off, idesc, acc);
            } else {
                if constexpr (kTS) umma_ts(tbase + kColD, tbase + kColA + k * 8, b0 + b_off, idesc, acc);
                else umma_ss(tbase + kColD, a0 + a_off, b0 + b_off, idesc, acc);
            }
        }
    };
    // Alternation patterns (N = 128): the attention kernel never runs more than 8 MMAs of one kind in a
    // row -- S = Q K^T (SS, fresh accumulator) and O += P V (TS, V MN-major) groups alternate.
    const uint32_t idesc_s = umma_idesc_f16(true, kPair ? 256 : 128, 128, false);
    const uint32_t idesc_o = umma_idesc_f16(true, kPair ? 256 : 128, 128, true);
    const uint64_t v0 = umma_smem_desc_sw128(sbase + kSmemB + 32 * 1024, 16 * 1024, 1024);
    auto mma_s = [&](int k, uint32_t d_col, uint32_t acc) {
        const uint32_t a_off = ((k >> 2) * (128 * 128) + (k & 3) * 32) >> 4;
        const uint32_t b_off = ((k >> 2) * (rows_b * 128) + (k & 3) * 32) >> 4;
        if constexpr (kPair) umma_ss_2cta(tbase + d_col, a0 + a_off, b0 + b_off, idesc_s, acc);
        else umma_ss(tbase + d_col, a0 + a_off, b0 + b_off, idesc_s, acc);
    };
    auto mma_o = [&](int k, uint32_t d_col) {
        if constexpr (kPair) umma_ts_2cta(tbase + d_col, tbase + kColA + k * 8, v0 + ((k * 2048) >> 4), idesc_o, 1u);
        else umma_ts(tbase + d_col, tbase + kColA + k * 8, v0 + ((k * 2048) >> 4), idesc_o, 1u);
    };
    auto issue_pattern = [&]() {
        if (pattern == 1) {         // S group, then PV group (the kernel's order)
#pragma unroll
            for (int k = 0; k < 8; ++k) mma_s(k, 0, k > 0);
#pragma unroll
            for (int k = 0; k < 8; ++k) mma_o(k, 128);
        } else if (pattern == 2) {  // the same 16 MMAs interleaved one by one
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                mma_s(k, 0, k > 0);
                mma_o(k, 128);
            }
        } else if (pattern == 3) {  // S groups alternating between two accumulators
#pragma unroll
            for (int k = 0; k < 8; ++k) mma_s(k, 0, k > 0);
#pragma unroll
            for (int k = 0; k < 8; ++k) mma_s(k, 128, k > 0);
        } else if (pattern == 4) {  // S groups into one accumulator, first MMA of each overwrites
#pragma unroll
            for (int k = 0; k < 8; ++k) mma_s(k, 0, k > 0);
#pragma unroll
            for (int k = 0; k < 8; ++k) mma_s(k, 0, k > 0);
        } else if (pattern == 5) {  // PV groups alternating between two accumulators
#pragma unroll
            for (int k = 0; k < 8; ++k) mma_o(k, 0);
#pragma unroll
            for (int k = 0; k < 8; ++k) mma_o(k, 128);
        } else {                    // 6: S group and PV group, one commit after each (as the kernel does)
#pragma unroll
            for (int k = 0; k < 8; ++k) mma_s(k, 0, k > 0);
            if constexpr (kPair) umma_commit_2cta(bar + 24, 3);
            else umma_commit(bar + 24);
#pragma unroll
            for (int k = 0; k < 8; ++k) mma_o(k, 128);
            if constexpr (kPair) umma_commit_2cta(bar + 24, 3);
            else umma_commit(bar + 24);
        }
    };
    auto commit = [&]() {
        if constexpr (kPair) umma_commit_2cta(bar, 3);
        else umma_commit(bar);
    };
    uint32_t phase = 0;
    // ---- 1. one K = 128 product, checked on the host; its latency on an idle pipe is recorded ----
    unsigned long long g0 = 0, g1 = 0;
    if (warp == 0 && rank == 0) {
        g0 = clock64();
        if (elect_one()) {
            issue_group(true);
            commit();
        }
        __syncwarp();
    }
    mbar_wait(bar, phase, 1);
    phase ^= 1;
    if (warp == 0 && rank == 0) {
        g1 = clock64();
        if (lane == 0) out[blockIdx.x].first_group = g1 - g0;
    }
    tc_fence_after();
    if (dout != nullptr && blockIdx.x < 2 && warp < 4) {
        const uint32_t t_d = tbase + (static_cast<uint32_t>(warp * 32) << 16) + kColD;
        const int r = warp * 32 + lane;
        for (int q = 0; q < N / 32; ++q) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(t_d + q * 32, v);
            tmem_wait_ld();
            for (int i = 0; i < 32; ++i)
                dout[((size_t)blockIdx.x * 128 + r) * 256 + q * 32 + i] = __uint_as_float(v[i]);
        }
    }
    tc_fence_before();
    if constexpr (kPair) cluster_sync();
    else __syncthreads();
    tc_fence_after();
    // ---- 1b. the same group issued right behind another one (pipe and operand path warm) ----
    if (pattern == 0) {
        if (warp == 0 && rank == 0) {
            g0 = clock64();
            if (elect_one()) {
                issue_group(false);
                issue_group(false);
                commit();
            }
            __syncwarp();
        }
        mbar_wait(bar, phase, 3);
        phase ^= 1;
        if (warp == 0 && rank == 0) {
            g1 = clock64();
            if (lane == 0) out[blockIdx.x].warm_group = g1 - g0;
        }
        tc_fence_before();
        if constexpr (kPair) cluster_sync();
        else __syncthreads();
        tc_fence_after();
    }
    // ---- 2. `groups` x 8 MMAs back to back, one commit at the end ----
    unsigned long long t0 = 0, t1 = 0, n0 = 0, n1 = 0;
    if (warp == 0 && rank == 0) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(n0));
        t0 = clock64();
        for (int g = 0; g < groups; ++g) {
            if (elect_one()) {
                if (pattern == 0) issue_group(false);
                else issue_pattern();
            }
            __syncwarp();
        }
        if (elect_one()) commit();
        __syncwarp();
    }
    if (warp >= 4 && noise != 0) {
        // interference agents: run until the commit of the timed MMAs has landed
        const bool do_math = noise == 1 || noise == 5;
        const bool do_bulk = (noise == 2 || noise == 3 || noise == 5) && warp == 4;
        const bool do_ldtm = noise == 4 || noise == 5;
        const long long pace = noise == 2 ? 0 : 700;
        float x[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = -0.01f * (float)(lane + i);
        float acc = 0.f;
        uint32_t nph = 0;
        long long next = clock64();
        const uint32_t t_n = tbase + (static_cast<uint32_t>((warp & 3) * 32) << 16) + 384;
        for (int iter = 0; iter < (1 << 22); ++iter) {
            if (mbar_try_wait(bar, phase)) break;
            if (do_bulk && __shfl_sync(0xffffffffu, (int)(clock64() >= next), 0)) {
                next += pace;
                if (elect_one()) {
                    mbar_arrive_expect_tx(bar_noise, 16384);
                    bulk_g2s(sbase + kSmemScratch + (iter & 1) * 16384, gsrc + ((size_t)(blockIdx.x * 8 + (iter & 7)) << 14),
                             16384, bar_noise);
                }
                __syncwarp();
                mbar_wait(bar_noise, nph, 7);
                nph ^= 1;
            }
            if (do_ldtm) {
                uint32_t v[4][32];
#pragma unroll
                for (int q = 0; q < 4; ++q) tmem_ld_32x32b_x32(t_n + q * 32, v[q]);
                tmem_wait_ld();
                acc += __uint_as_float(v[0][0] ^ v[1][1] ^ v[2][2] ^ v[3][3]);
            }
            if (do_math) {
#pragma unroll
                for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float p = ex2_approx(fmaf(x[i], 0.99f, -0.001f));
                        acc += p;
                        x[i] = fmaf(p, -0.01f, x[i] * 0.5f);
                    }
                }
            }
        }
        if (sink != nullptr && acc == 123.456f) sink[threadIdx.x] = acc;
    }
    mbar_wait(bar, phase, 2);
    phase ^= 1;
    if (warp == 0 && rank == 0) {
        t1 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(n1));
        if (lane == 0) {
            out[blockIdx.x].cycles = t1 - t0;
            out[blockIdx.x].ns = n1 - n0;
        }
    }
    tc_fence_before();
    if constexpr (kPair) cluster_sync();
    else __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        if constexpr (kPair) tmem_dealloc_2cta(tbase, 512);
        else tmem_dealloc(tbase, 512);
    }
}

template <int kMode>
__global__ void __launch_bounds__(256, 1)
probe_1cta(int N, int groups, ProbeOut* out, float* dout, int noise, int rnd, const uint8_t* gsrc, float* sink,
           int pattern) {
    probe_body<kMode>(N, groups, out, dout, noise, rnd, gsrc, sink, pattern);
}
template <int kMode>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
probe_2cta(int N, int groups, ProbeOut* out, float* dout, int noise, int rnd, const uint8_t* gsrc, float* sink,
           int pattern) {
    probe_body<kMode>(N, groups, out, dout, noise, rnd, gsrc, sink, pattern);
}

template <class Kern>
static void run(const char* name, Kern kern, int mode, int N, int n_sm, int noise = 0, int rnd = 0,
                int pattern = 0) {
    const bool pair = (mode & 2) != 0;
    const int groups = 2000;
    const int grid = pair ? (n_sm / 2) * 2 : n_sm;
    ProbeOut* d_out;
    float* d_d;
    CK(cudaMalloc(&d_out, sizeof(ProbeOut) * grid));
    CK(cudaMemset(d_out, 0, sizeof(ProbeOut) * grid));
    CK(cudaMalloc(&d_d, sizeof(float) * 2 * 128 * 256));
    CK(cudaMemset(d_d, 0, sizeof(float) * 2 * 128 * 256));
    static uint8_t* d_src = nullptr;  // source of the interference bulk copies: 128 KiB per CTA
    if (d_src == nullptr) {
        CK(cudaMalloc(&d_src, (size_t)160 * 8 * 16384));
        CK(cudaMemset(d_src, 1, (size_t)160 * 8 * 16384));
    }
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    for (int rep = 0; rep < 2; ++rep) {
        kern<<<grid, 256, kSmemBytes>>>(N, groups, d_out, d_d, noise, rnd, d_src, nullptr, pattern);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("{\"name\": \"%s\", \"N\": %d, \"error\": \"%s\"}\n", name, N, cudaGetErrorString(e));
            exit(3);  // the context is gone
        }
    }
    std::vector<ProbeOut> out(grid);
    std::vector<float> d(2 * 128 * 256);
    CK(cudaMemcpy(out.data(), d_out, sizeof(ProbeOut) * grid, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(d.data(), d_d, sizeof(float) * d.size(), cudaMemcpyDeviceToHost));
    // check: CTA b of the first two CTAs holds accumulator rows (pair: b*128 + r; single: r)
    int bad = 0;
    double maxerr = 0;
    for (int b = 0; b < 2; ++b)
        for (int r = 0; r < 128; ++r)
            for (int n = 0; n < N; ++n) {
                const int rg = pair ? b * 128 + r : r;
                double ref = 0;
                for (int k = 0; k < 128; ++k) ref += (double)a_val(rg, k) * b_val(n, k);
                const double err = fabs(ref - d[((size_t)b * 128 + r) * 256 + n]);
                maxerr = std::max(maxerr, err);
                bad += err > 1e-3;
            }
    if (rnd || (mode & 4)) bad = -1;  // timing-only configurations: result not checked
    std::vector<double> per, first, warm;
    double mhz = 0;
    for (int i = 0; i < grid; i += pair ? 2 : 1) {
        per.push_back((double)out[i].cycles / ((pattern ? 16.0 : 8.0) * groups));
        first.push_back((double)out[i].first_group);
        warm.push_back((double)out[i].warm_group);
        mhz += out[i].ns ? (double)out[i].cycles / out[i].ns * 1e3 : 0;
    }
    std::sort(per.begin(), per.end());
    std::sort(first.begin(), first.end());
    std::sort(warm.begin(), warm.end());
    const double med = per[per.size() / 2];
    const double flop_per_mma = 2.0 * (pair ? 256 : 128) * N * 16;
    printf("{\"name\": \"%s\", \"mode\": %d, \"pattern\": %d, \"noise\": %d, \"rnd\": %d, \"N\": %d, \"clk_per_mma_median\": %.1f, \"min\": %.1f, "
           "\"max\": %.1f, \"ideal_clk\": %.1f, \"flop_per_clk_per_sm\": %.0f, \"sm_mhz\": %.0f, "
           "\"mismatches\": %d, \"maxerr\": %.3g, \"group8_latency_cold\": %.0f, \"two_groups_latency\": %.0f}\n",
           name, mode, pattern, noise, rnd, N, med, per.front(), per.back(), N / 2.0, flop_per_mma / med / (pair ? 2 : 1),
           mhz / per.size(), bad, maxerr, first[first.size() / 2], warm[warm.size() / 2]);
    fflush(stdout);
    cudaFree(d_out);
    cudaFree(d_d);
}

int main(int argc, char** argv) {
    int n_sm = 148;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    n_sm = prop.multiProcessorCount;
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    const int ns[3] = {64, 128, 256};
    for (int mode = 0; mode < 4; ++mode) {
        if (only == -2 || (only >= 0 && mode != only)) continue;
        for (int N : ns) {
            switch (mode) {
                case 0: run("ss_1cta", probe_1cta<0>, 0, N, n_sm); break;
                case 1: run("ts_1cta", probe_1cta<1>, 1, N, n_sm); break;
                case 2: run("ss_2cta", probe_2cta<2>, 2, N, n_sm); break;
                case 3: run("ts_2cta", probe_2cta<3>, 3, N, n_sm); break;
            }
        }
    }
    if (only >= 0) return 0;
    // group alternation (see issue_pattern): clk per MMA averaged over the 16 MMAs of one round
    for (int rnd = 0; rnd < 2; ++rnd)
        for (int pattern = 1; pattern <= 6; ++pattern) {
            run("alt_1cta", probe_1cta<0>, 0, 128, n_sm, 0, rnd, pattern);
            run("alt_2cta", probe_2cta<2>, 2, 128, n_sm, 0, rnd, pattern);
        }
    if (only == -2) return 0;
    // the attention kernel's four MMA flavours (N = 128) under interference, small-integer and random operands
    for (int rnd = 0; rnd < 2; ++rnd)
        for (int noise = 0; noise <= 5; ++noise) {
            run("ss_1cta", probe_1cta<0>, 0, 128, n_sm, noise, rnd);     // QK^T, one CTA
            run("ss_2cta", probe_2cta<2>, 2, 128, n_sm, noise, rnd);     // QK^T, CTA pair
            run("tsT_1cta", probe_1cta<5>, 5, 128, n_sm, noise, rnd);    // P V (V MN-major), one CTA
            run("tsT_2cta", probe_2cta<7>, 7, 128, n_sm, noise, rnd);    // P V, CTA pair
        }
    return 0;
}

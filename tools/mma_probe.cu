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
constexpr int kSmemBytes = kSmemBar + 64;
constexpr uint32_t kColD = 0, kColA = 256;

struct ProbeOut {
    unsigned long long cycles;
    unsigned long long ns;
};

// kMode: bit 0 = A from tensor memory, bit 1 = CTA pair
template <int kMode>
__device__ __forceinline__ void probe_body(int N, int groups, ProbeOut* out, float* dout) {
    constexpr bool kTS = (kMode & 1) != 0;
    constexpr bool kPair = (kMode & 2) != 0;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar = sbase + kSmemBar;
    const uint32_t tmem_ptr = sbase + kSmemBar + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = kPair ? cluster_ctarank() : 0u;
    const int rows_b = kPair ? N / 2 : N;  // B rows held by this CTA

    if (warp == 0) {
        if (lane == 0) {
            mbar_init(bar, 1);
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
        *reinterpret_cast<uint16_t*>(smem + kSmemA + off) = bf16_bits(a_val((int)rank * 128 + r, k));
    }
    for (int idx = threadIdx.x; idx < rows_b * 128; idx += blockDim.x) {
        const int r = idx >> 7, k = idx & 127;
        const int h = k >> 6, c = (k & 63) >> 3, e = k & 7;
        const uint32_t off = h * (rows_b * 128) + r * 128 + ((c ^ (r & 7)) * 16) + e * 2;
        *reinterpret_cast<uint16_t*>(smem + kSmemB + off) = bf16_bits(b_val((int)rank * rows_b + r, k));
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *reinterpret_cast<volatile uint32_t*>(smem + kSmemBar + 16);
    if (kTS && warp < 4) {  // A as the MMA reads it from tensor memory: lane = row, column = k / 2
        const uint32_t t_a = tbase + (static_cast<uint32_t>(warp * 32) << 16) + kColA;
        const int r = (int)rank * 128 + warp * 32 + lane;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            uint32_t v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int k = (q * 32 + i) * 2;
                v[i] = (uint32_t)bf16_bits(a_val(r, k)) | ((uint32_t)bf16_bits(a_val(r, k + 1)) << 16);
            }
            tmem_st_32x32b_x32(t_a + q * 32, v);
        }
        tmem_wait_st();
    }
    tc_fence_before();
    if constexpr (kPair) cluster_sync();
    else __syncthreads();
    tc_fence_after();

    const uint32_t idesc = umma_idesc_f16(true, kPair ? 256 : 128, N, false);
    const uint64_t a0 = umma_smem_desc_sw128(sbase + kSmemA, 16, 1024);
    const uint64_t b0 = umma_smem_desc_sw128(sbase + kSmemB, 16, 1024);
    auto issue_group = [&](bool first) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t a_off = ((k >> 2) * (128 * 128) + (k & 3) * 32) >> 4;
            const uint32_t b_off = ((k >> 2) * (rows_b * 128) + (k & 3) * 32) >> 4;
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
    auto commit = [&]() {
        if constexpr (kPair) umma_commit_2cta(bar, 3);
        else umma_commit(bar);
    };
    uint32_t phase = 0;
    // ---- 1. one K = 128 product, checked on the host ----
    if (warp == 0 && rank == 0) {
        if (elect_one()) {
            issue_group(true);
            commit();
        }
        __syncwarp();
    }
    mbar_wait(bar, phase, 1);
    phase ^= 1;
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
    // ---- 2. `groups` x 8 MMAs back to back, one commit at the end ----
    unsigned long long t0 = 0, t1 = 0, n0 = 0, n1 = 0;
    if (warp == 0 && rank == 0) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(n0));
        t0 = clock64();
        for (int g = 0; g < groups; ++g) {
            if (elect_one()) issue_group(false);
            __syncwarp();
        }
        if (elect_one()) commit();
        __syncwarp();
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
__global__ void __launch_bounds__(128, 1) probe_1cta(int N, int groups, ProbeOut* out, float* dout) {
    probe_body<kMode>(N, groups, out, dout);
}
template <int kMode>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe_2cta(int N, int groups, ProbeOut* out, float* dout) {
    probe_body<kMode>(N, groups, out, dout);
}

template <class Kern>
static void run(const char* name, Kern kern, int mode, int N, int n_sm) {
    const bool pair = (mode & 2) != 0;
    const int groups = 2000;
    const int grid = pair ? (n_sm / 2) * 2 : n_sm;
    ProbeOut* d_out;
    float* d_d;
    CK(cudaMalloc(&d_out, sizeof(ProbeOut) * grid));
    CK(cudaMemset(d_out, 0, sizeof(ProbeOut) * grid));
    CK(cudaMalloc(&d_d, sizeof(float) * 2 * 128 * 256));
    CK(cudaMemset(d_d, 0, sizeof(float) * 2 * 128 * 256));
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    for (int rep = 0; rep < 2; ++rep) {
        kern<<<grid, 128, kSmemBytes>>>(N, groups, d_out, d_d);
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
    std::vector<double> per;
    double mhz = 0;
    for (int i = 0; i < grid; i += pair ? 2 : 1) {
        per.push_back((double)out[i].cycles / (8.0 * groups));
        mhz += out[i].ns ? (double)out[i].cycles / out[i].ns * 1e3 : 0;
    }
    std::sort(per.begin(), per.end());
    const double med = per[per.size() / 2];
    const double flop_per_mma = 2.0 * (pair ? 256 : 128) * N * 16;
    printf("{\"name\": \"%s\", \"mode\": %d, \"N\": %d, \"clk_per_mma_median\": %.1f, \"min\": %.1f, "
           "\"max\": %.1f, \"ideal_clk\": %.1f, \"flop_per_clk_per_sm\": %.0f, \"sm_mhz\": %.0f, "
           "\"mismatches\": %d, \"maxerr\": %.3g}\n",
           name, mode, N, med, per.front(), per.back(), N / 2.0, flop_per_mma / med / (pair ? 2 : 1),
           mhz / per.size(), bad, maxerr);
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
        if (only >= 0 && mode != only) continue;
        for (int N : ns) {
            switch (mode) {
                case 0: run("ss_1cta", probe_1cta<0>, 0, N, n_sm); break;
                case 1: run("ts_1cta", probe_1cta<1>, 1, N, n_sm); break;
                case 2: run("ss_2cta", probe_2cta<2>, 2, N, n_sm); break;
                case 3: run("ts_2cta", probe_2cta<3>, 3, N, n_sm); break;
            }
        }
    }
    return 0;
}

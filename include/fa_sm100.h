/*
 * C ABI of libfa_sm100.so -- the B200 (sm_100a) fused attention forward.
 *
 * This is the drop-in boundary for the one hot path of sonnyli/flash_attention_from_scratch:
 * it replaces the pybind11 module `flash_attention_kernels`
 * (/root/reference/src/flash_attention.cu:34-149) that the reference's Python operator
 * `flash_attention.forward / forward_timed` (/root/reference/flash_attention/__init__.py:7-17)
 * calls.  Plain pointers and sizes only -- no torch types cross this boundary.
 *
 * Tensor contract (same as the reference launcher, flash_attention.cu:38-98):
 *   Q, K, V, O are 16-bit (fp16 or bf16) device tensors of logical shape
 *   (batch, seq_len, n_heads, d_head) with d_head == 128 contiguous; the other three strides are
 *   given in ELEMENTS (as torch's Tensor.stride()) and shared by all four tensors, exactly as the
 *   reference takes them from Q only (flash_attention.cu:84-86).  Any seq_len >= 1 is accepted
 *   (keys beyond seq_len in the last 128-block are masked to -inf); the reference requires a
 *   multiple of its B_r/B_c tile (flash_attention.cu:79-82) and the Python operator keeps that
 *   check when it is given a reference-style kernel_cfg.
 *   Unlike the reference kernel (static_kernel_configuration.cuh:146: row stride hard-coded to
 *   d_head * 16) any n_heads and any 16-byte-aligned strides are accepted.
 *
 * All functions are thread-safe and may be used on several devices from one process; the current
 * CUDA device must be the one that owns the pointers (the Python operator sets it).
 */
#ifndef FA_SM100_H_
#define FA_SM100_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* dtype codes == torch ScalarType values, as in the reference's DType enum
 * (/root/reference/py/flash_helpers/kernel_configs.py:9-14). */
#define FA_DTYPE_FP16 5
#define FA_DTYPE_BF16 15

/* status codes (0 == success); fa_last_error_string() has the text */
#define FA_OK 0
#define FA_ERR_DTYPE 1      /* "Only fp16 and bf16 are supported"            flash_attention.cu:51-52 */
#define FA_ERR_DHEAD 2      /* d_head != 128 ("Kernel configuration was not found", :60-61)        */
#define FA_ERR_SEQLEN 3     /* seq_len out of range (the B_r/B_c multiple check of flash_attention.cu:79-82
                               lives in the Python operator)                                        */
#define FA_ERR_ARG 4        /* null pointer / non-positive size / misaligned pointer or stride     */
#define FA_ERR_DEVICE 5     /* not an sm_100 device / no CUDA driver ("requires SM_80", :46-48)    */
#define FA_ERR_TENSORMAP 6  /* cuTensorMapEncodeTiled rejected the layout                          */
#define FA_ERR_LAUNCH 7     /* CUDA launch / runtime error (the reference never checks, :126)      */

/* Asynchronous launch on `stream` (a cudaStream_t; NULL == legacy default stream).
 * Replaces flash_attention_forward(..., benchmark=false), flash_attention.cu:34-135. */
int fa_fwd(const void* q, const void* k, const void* v, void* o, int batch, int seq_len,
           int n_heads, int d_head, int64_t stride_batch, int64_t stride_seq, int64_t stride_head,
           int dtype, void* stream);

/* Same, bracketed by a cudaEvent pair on `stream` and synchronised; *ms receives the kernel time.
 * Replaces the benchmark=true path, flash_attention.cu:119-132. */
int fa_fwd_timed(const void* q, const void* k, const void* v, void* o, int batch, int seq_len,
                 int n_heads, int d_head, int64_t stride_batch, int64_t stride_seq,
                 int64_t stride_head, int dtype, void* stream, float* ms);

/* End-to-end convenience for HOST buffers (contiguous (batch, seq, heads, 128)): copies Q, K, V to
 * device `device`, runs the kernel and copies O back, pipelined over the batch dimension on
 * internal streams; returns after O is complete in host memory.  Pinned host memory makes the
 * copies asynchronous; pageable memory works but serialises.  The reference has no such entry
 * (its operator only accepts CUDA tensors); this exists for end-to-end measurement. */
int fa_fwd_host(const void* q_host, const void* k_host, const void* v_host, void* o_host,
                int batch, int seq_len, int n_heads, int d_head, int dtype, int device);

/* Frees the device workspace cached by fa_fwd_host on `device` (-1: all devices). */
int fa_host_workspace_free(int device);

/* Text of the last error raised on the calling thread ("" if none). */
const char* fa_last_error_string(void);

/* Device facts the launcher relies on (cuda_utils.cuh:29-46 in the reference). */
int fa_device_info(int device, int* n_sms, int* smem_optin_bytes, int* compute_capability);

/* Static facts about the built kernel: dynamic shared memory per CTA, threads per CTA,
 * query rows per work tile (the kernel is persistent: one CTA per SM walks the tiles), TMEM columns. */
int fa_kernel_info(int* smem_bytes, int* threads, int* rows_per_cta, int* tmem_cols);

/* Number of kernel launches issued through this library since load (all threads). */
int64_t fa_launch_count(void);

/* Which kernel a launch uses.  The reference selects one of its 85 template instantiations with the
 * kernel_cfg map lookup (flash_attention.cu:59-62); this library has two machine mappings of the same
 * arithmetic (three kernels) and picks by shape:
 *   FA_MODE_AUTO     (default): the ping-pong kernel up to seq_len 1024, CTA pairs above (measured crossover) except on grids of
 *                    a few waves where half-size tiles need fewer of them (fa_pick_kernel)
 *   FA_MODE_SINGLE   one CTA per SM, work tile = 256 query rows
 *   FA_MODE_PAIR     clusters of two CTAs sharing every K/V block (tcgen05 cta_group::2), tile = 512 rows
 *   FA_MODE_PINGPONG clusters of two CTAs, one 128-row Q tile per CTA, two S accumulators (tile = 256 rows)
 * Process-wide; returns the previous mode (or -1 for an invalid argument).  The environment variable
 * FA_SM100_MODE=auto|single|pair|pingpong sets the initial value. */
#define FA_MODE_AUTO 0
#define FA_MODE_SINGLE 1
#define FA_MODE_PAIR 2
#define FA_MODE_PINGPONG 3 /* CTA pairs, one 128-row Q tile per CTA, the two softmax warpgroups alternate KV
                              blocks on two S accumulators (csrc/fa_fwd_pp_sm100.cuh); tile = 256 rows */
int fa_set_kernel_mode(int mode);

/* Same choice for the CALLING THREAD only; overrides the process-wide mode until reset with -1.
 * This is what a per-call `kernel_cfg` selection maps to (the reference looks its kernel up per call,
 * flash_attention.cu:59-62): the Python operator brackets a launch with it when kernel_cfg.cta_group is
 * 1 (single CTAs) or 2 (CTA pairs).  Returns the previous override (-1 = none), -2 for a bad argument. */
int fa_set_thread_kernel_mode(int mode);

/* FA_MODE_SINGLE / FA_MODE_PAIR / FA_MODE_PINGPONG of the calling thread's last launch (-1: none yet): what
 * AUTO actually picked, so that callers (bench.py) need not re-implement the rule. */
int fa_last_kernel(void);

/* The kernel AUTO (or the mode in force for the calling thread) would launch for this problem on a device with
 * `n_sms` SMs (<= 0: 148), without launching anything: FA_MODE_SINGLE / FA_MODE_PAIR / FA_MODE_PINGPONG.  AUTO =
 * ping-pong up to seq_len 1024; above, CTA pairs unless the grid is so small that the ping-pong kernel's half-size
 * work tiles need fewer waves (cost model in csrc/fa_api.cu: pick_kernel). */
int fa_pick_kernel(int seq_len, int batch, int n_heads, int n_sms);

/* Hit / miss counters of the per-thread cache of encoded TMA tensor maps (key: pointer, shape, strides, dtype;
 * the reference builds nothing per call, its kernel takes raw pointers: flash_attention.cu:107-110). */
int fa_tensor_map_cache_stats(int64_t* hits, int64_t* misses);

/* Bring-up entry: runs the debug instantiation with explicit descriptor knobs and a dump buffer
 * (see FwdDebug in csrc/fa_fwd_sm100.cuh).  knobs[7] = bring-up level (1 setup only, 2 TMA,
 * 3 QK^T, >= 4 everything; knobs[0..6] are ignored); `dump` and `diag` should be host-mapped
 * (pinned) so they survive a trapped kernel.  Synchronous.  Not part of the reference interface. */
int fa_fwd_debug(const void* q, const void* k, const void* v, void* o, int batch, int seq_len,
                 int n_heads, int d_head, int64_t stride_batch, int64_t stride_seq,
                 int64_t stride_head, int dtype, float* dump, const uint32_t* knobs,
                 uint32_t* diag);

#ifdef __cplusplus
}
#endif
#endif /* FA_SM100_H_ */

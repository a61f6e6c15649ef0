"""CPU/torch restatement of the attention forward the reference computes (test infrastructure).

Three views of the same function, each citing what it follows:

* `py_flash_attention`  -- the reference's own test oracle
  (/root/reference/py/flash_helpers/test/utils.py:137-162): S = QK^T/sqrt(d) via einsum in the
  input dtype (or fp32 when `upcast`), softmax over keys, O = PV.
* `sdpa_ref`            -- the north-star oracle (BASELINE.json): torch SDPA on (B,H,N,d) views.
* `blockwise_kernel_ref`-- the KERNEL's arithmetic, block by block
  (/root/reference/tools/debug/debug.py:102-151 fused-softmax branch,
  /root/reference/src/include/softmax.cuh:15-128, load_store.cuh:336-351,
  forward_kernel.cuh:150-152): exp2 with the scale folded in, fp32 row sum of UN-rounded P,
  P rounded to the 16-bit dtype before PV, fp32 accumulation, final O/l rounded to 16 bit.
  Optional lazy rescale (threshold in log2 units) restates what the B200 kernel adds.

All tensors are (batch, seq, heads, d_head) like the reference (utils.py:112-121).
"""
from __future__ import annotations

import math

import torch


def py_flash_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, upcast: bool = False):
    """utils.py:137-162.  `upcast=False` runs every op in the input dtype (16-bit einsum and
    softmax), `upcast=True` computes in fp32 and casts the result back."""
    d_head = q.shape[-1]
    dtype_og = q.dtype
    if upcast:
        q, k, v = q.float(), k.float(), v.float()
    s = torch.einsum("bqhd,bkhd->bqhk", q, k) / (d_head ** 0.5)
    p = s.softmax(dim=-1)
    out = torch.einsum("bqhk,bkhd->bqhd", p, v)
    if upcast:
        out = out.to(dtype=dtype_og)
    return out


def sdpa_ref(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, fp32: bool = False):
    """torch.nn.functional.scaled_dot_product_attention on the transposed views; result in the
    input dtype.  `fp32=True` computes in fp32 and rounds once at the end."""
    dt = q.dtype
    if fp32:
        q, k, v = q.float(), k.float(), v.float()
    o = torch.nn.functional.scaled_dot_product_attention(
        q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
    ).transpose(1, 2)
    return o.to(dt).contiguous()


def ex2_emulated_ref(x: torch.Tensor) -> torch.Tensor:
    """Bit-level restatement (fp32 torch ops) of the kernel's FMA-pipe exp2
    (flash_attention_from_scratch_b200/csrc/ptx_sm100.cuh: ex2_emulated_x2), which replaces the
    MUFU `ex2.approx` of the reference (softmax.cuh:51-64 under --use_fast_math) for a quarter of the
    elements: clamp at -127, floor via the round-down magic-number add, degree-3 polynomial for
    2^frac, integer part added into the exponent field."""
    import numpy as np

    xf = x.detach().to(torch.float32).cpu().numpy().astype(np.float32)
    xf = np.maximum(xf, np.float32(-127.0))
    # t = x + (2^23 + 2^22) rounded DOWN: exact in float64, then rounded toward -inf to fp32
    t64 = xf.astype(np.float64) + 12582912.0
    t = t64.astype(np.float32)
    t = np.where(t.astype(np.float64) > t64, np.nextafter(t, np.float32(-np.inf)), t).astype(np.float32)
    r = (t - np.float32(12582912.0)).astype(np.float32)          # floor(x) as float
    f = (xf - r).astype(np.float32)                              # fractional part in [0, 1)
    fma = lambda a, b, c_: (a.astype(np.float64) * b.astype(np.float64) + np.float64(c_)).astype(np.float32)  # noqa: E731
    p = fma(np.full_like(f, np.float32(0.07706724)), f, np.float32(0.22764498))
    p = fma(p, f, np.float32(0.69511664))
    p = fma(p, f, np.float32(1.0))
    bits = p.view(np.int32) + (t.view(np.int32) << 23)           # 2^floor(x) into the exponent
    return torch.from_numpy(bits.view(np.float32).copy()).to(x.device)


def blockwise_kernel_ref(q, k, v, block: int = 128, reverse: bool = False,
                         rescale_threshold: float = 0.0, return_stats: bool = False,
                         emulated_pairs: int = 0):
    """Block-wise online-softmax attention with the kernel's rounding points.

    reverse=True walks KV blocks N/B_c-1 .. 0 like the reference (forward_kernel.cuh:142,180);
    the B200 kernel walks 0 .. N/B_c-1.  rescale_threshold=0 rescales on every max increase
    (reference, softmax.cuh:37-49); 8.0 restates the B200 kernel's lazy rescale.  emulated_pairs=4
    additionally restates its polynomial exp2 on a quarter of the elements (production setting).
    """
    B, N, H, D = q.shape
    dt = q.dtype
    c = math.log2(math.e) / math.sqrt(D)  # forward_kernel.cuh:150-151
    qf = q.float().permute(0, 2, 1, 3)    # (B,H,N,D)
    kf = k.float().permute(0, 2, 1, 3)
    vf = v.float().permute(0, 2, 1, 3)
    m = torch.full((B, H, N, 1), float("-inf"))
    l = torch.zeros((B, H, N, 1))
    o = torch.zeros((B, H, N, D))
    blocks = list(range(0, N, block))
    if reverse:
        blocks = blocks[::-1]
    first = True
    for j0 in blocks:
        s = qf @ kf[:, :, j0:j0 + block].transpose(-1, -2)          # fp32 accumulate of 16-bit products
        m_blk = torch.maximum(m, s.max(dim=-1, keepdim=True).values)  # softmax.cuh:15-35
        if first:
            m_new = m_blk
            alpha = torch.ones_like(l)
        else:
            grow = (m_blk - m) * c
            take = grow > rescale_threshold
            m_new = torch.where(take, m_blk, m)
            alpha = torch.where(take, torch.exp2((m - m_blk) * c), torch.ones_like(l))
        p32 = torch.exp2(s * c - m_new * c)                          # softmax.cuh:51-64
        if emulated_pairs:
            # the kernel computes `emulated_pairs` of every 16 (even, odd) key pairs of the first three
            # 32-key fragments of a block with the polynomial (softmax_sm100.cuh: emulate_pair)
            xs = s * c - m_new * c
            sel = torch.zeros(block, dtype=torch.bool)
            for frag in range(min(3, block // 32)):
                for pair in range(16):
                    if (pair * emulated_pairs) % 16 < emulated_pairs:
                        sel[frag * 32 + 2 * pair: frag * 32 + 2 * pair + 2] = True
            sel = sel[: xs.shape[-1]]
            p32 = torch.where(sel, ex2_emulated_ref(xs), p32)
        l = alpha * l + p32.sum(dim=-1, keepdim=True)                # un-rounded fp32 sum, :66-83
        o = alpha * o + p32.to(dt).float() @ vf[:, :, j0:j0 + block]  # P rounded RN to 16 bit
        m = m_new
        first = False
    out = (o / l).to(dt).permute(0, 2, 1, 3).contiguous()           # softmax.cuh:107-128
    if return_stats:
        return out, m.squeeze(-1), l.squeeze(-1)
    return out


def reference_pass_criterion(out, ref16, ref32, factor: float = 2.0):
    """The reference's acceptance test (py/flash_helpers/test/test.py:58-61):
    max|out - ref16| <= 2 * max|ref16 - ref32|.  Returns (passed, diff_out, diff_ref)."""
    d_out = (out.float() - ref16.float()).abs().max().item()
    d_ref = (ref16.float() - ref32.float()).abs().max().item()
    return d_out <= factor * d_ref, d_out, d_ref

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


def blockwise_kernel_ref(q, k, v, block: int = 128, reverse: bool = False,
                         rescale_threshold: float = 0.0, return_stats: bool = False):
    """Block-wise online-softmax attention with the kernel's rounding points.

    reverse=True walks KV blocks N/B_c-1 .. 0 like the reference (forward_kernel.cuh:142,180);
    the B200 kernel walks 0 .. N/B_c-1.  rescale_threshold=0 rescales on every max increase
    (reference, softmax.cuh:37-49); 8.0 restates the B200 kernel's lazy rescale.
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

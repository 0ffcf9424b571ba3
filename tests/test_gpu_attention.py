"""Kernel-level parity of the TransformerBlock (-m gpu): the attention kernels alone, through the C-ABI
(sshslie_transformer_block), against oracle.transformer_block (the restatement of model.py:99-119) - forward and
backward, at token counts that are not multiples of the kernels' blocks (6, 35), at the training size (256) and at the
full-image inference size (4096, where the tcgen05 core of attention_tc.cu takes over; 1156 = ragged tensor-core tiles).

Tolerances: the block's input and output are bf16 tensors in the engine (2^-8 relative); q, k, v, the logits, the softmax
and the FFN are fp32 (tensor-core path: bf16 hi+lo operand pairs for the logits, bf16 probabilities).  Weights are scaled
up so that the logits spread over several units and the softmax is far from uniform."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
PREFIX = "illum_adjust_net.attn."
NAMES = ["q_linear", "k_linear", "v_linear", "ff_linear1", "ff_linear2"]


def _params(seed, gain):
    g = torch.Generator().manual_seed(seed)
    p = {}
    for nm in NAMES:
        p[PREFIX + nm + ".weight"] = torch.randn(64, 64, generator=g) * gain / 8.0
        p[PREFIX + nm + ".bias"] = torch.randn(64, generator=g) * 0.1
    return p


def _flat(p):
    return torch.cat([p[PREFIX + nm + sfx].flatten() for nm in NAMES for sfx in (".weight", ".bias")])


def _run(x, p, dy=None):
    import sshslie_b200 as S
    from gpu_util import stream
    lib = S.lib.load()
    B, C, H, W = x.shape
    nbytes = lib.sshslie_transformer_block_scratch_bytes(B, H, W)
    scratch = torch.empty(nbytes + 1024, dtype=torch.uint8, device="cuda")
    base = (scratch.data_ptr() + 1023) // 1024 * 1024
    xd, pd = x.cuda().contiguous(), _flat(p).cuda().contiguous()
    y = torch.empty_like(xd)
    dx = torch.empty_like(xd) if dy is not None else None
    dp = torch.empty(20800, device="cuda") if dy is not None else None
    dyd = dy.cuda().contiguous() if dy is not None else None
    S.lib.check(lib.sshslie_transformer_block(int(dy is not None), S.lib.ptr(xd), S.lib.ptr(pd), S.lib.ptr(y),
                                              S.lib.ptr(dyd), S.lib.ptr(dx), S.lib.ptr(dp), B, H, W,
                                              ctypes.c_void_p(base), nbytes, stream()), "sshslie_transformer_block")
    torch.cuda.synchronize()
    return y.cpu(), (dx.cpu() if dx is not None else None), (dp.cpu() if dp is not None else None)


@pytest.mark.parametrize("shape", [(2, 2, 3), (1, 5, 7), (2, 16, 16), (1, 34, 34), (2, 40, 40), (1, 64, 64)],
                         ids=["L6", "L35", "L256", "L1156_tc_ragged", "L1600_tc_ragged_b2", "L4096_tc"])
def test_transformer_block_forward(shape):
    from gpu_util import bf16_round
    from oracle import sshslie_oracle as O
    B, H, W = shape
    g = torch.Generator().manual_seed(H * 100 + W)
    x = bf16_round(torch.randn(B, 64, H, W, generator=g))
    p = _params(7, gain=2.5)
    y, _, _ = _run(x, p)
    ref = O.transformer_block(p, x)
    # logits spread: make sure the test exercises a non-trivial softmax
    t = x.reshape(B, 64, H * W).permute(0, 2, 1)
    q = torch.nn.functional.linear(t, p[PREFIX + "q_linear.weight"], p[PREFIX + "q_linear.bias"])
    k = torch.nn.functional.linear(t, p[PREFIX + "k_linear.weight"], p[PREFIX + "k_linear.bias"])
    assert float((q[..., :16] @ k[..., :16].transpose(-1, -2) / 4).std()) > 1.0
    torch.testing.assert_close(y, ref, rtol=2 ** -7, atol=2 ** -7 * float(ref.abs().max()) * 0.25)


@pytest.mark.parametrize("shape", [(2, 2, 3), (1, 5, 7), (2, 16, 16), (1, 64, 64)], ids=["L6", "L35", "L256", "L4096"])
def test_transformer_block_backward(shape):
    """dx and the ten parameter gradients of <y, dy> against torch autograd on the oracle block.  dx is a bf16 tensor in
    the engine (2^-8 relative to its largest entry) and carries the backward mask of the ReLU layer that produces the
    block's input in the network (x > 0); the weight gradients are fp32 sums."""
    from gpu_util import bf16_round
    from oracle import sshslie_oracle as O
    B, H, W = shape
    g = torch.Generator().manual_seed(H * 10 + W)
    x = bf16_round(torch.randn(B, 64, H, W, generator=g))
    dy = bf16_round(torch.randn(B, 64, H, W, generator=g))
    p = _params(9, gain=2.0)
    _, dx, dp = _run(x, p, dy)
    xr = x.clone().requires_grad_(True)
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    (O.transformer_block(pr, xr) * dy).sum().backward()
    want_dx = xr.grad * (x > 0)
    tol = 2 ** -6 * want_dx.abs() + 2 ** -7 * float(want_dx.abs().max())
    bad = (dx - want_dx).abs() > tol
    # hidden units of the FFN that sit within rounding of 0 take the other side of the ReLU than in the oracle (the
    # tensor-core forward and torch differ in the last bits of o): a handful of entries at the largest size, each bounded
    assert float(bad.float().mean()) <= (2e-4 if H * W >= 1024 else 0.0), float(bad.float().mean())
    assert float((dx - want_dx).abs().max()) <= 0.05 * float(want_dx.abs().max())
    want = torch.cat([pr[PREFIX + nm + sfx].grad.flatten() for nm in NAMES for sfx in (".weight", ".bias")])
    off = 0
    for nm in NAMES:
        for sfx, n in ((".weight", 4096), (".bias", 64)):
            a, b = dp[off:off + n], want[off:off + n]
            scale = float(b.abs().max())
            if nm == "k_linear" and sfx == ".bias":          # softmax is shift invariant: the true gradient is ~0
                assert float(a.abs().max()) <= 1e-3 * float(want.abs().max())
            else:
                atol = (2e-2 if H * W >= 1024 else 5e-3) * scale      # (ReLU flips of the FFN hidden units, see above)
                torch.testing.assert_close(a, b, rtol=5e-3, atol=atol, msg=lambda m: f"{nm}{sfx}: {m}")
            off += n


def test_transformer_block_is_bit_repeatable():
    from gpu_util import bf16_round
    g = torch.Generator().manual_seed(3)
    x = bf16_round(torch.randn(2, 64, 16, 16, generator=g))
    dy = bf16_round(torch.randn(2, 64, 16, 16, generator=g))
    p = _params(9, gain=2.0)
    a = _run(x, p, dy)
    b = _run(x, p, dy)
    assert all(torch.equal(u, v) for u, v in zip(a, b))

"""End-to-end parity of the drop-in module (-m gpu): LowLightEnhance.forward / compute_loss / backward / Adam on a
B200 against the CPU oracle on the same seeded inputs, and against the fixtures recorded from the unmodified
reference.  Tolerances (bf16 tensor-core operands, fp32 accumulate, fp32 heads and loss kernels):
  outputs R/I/S  <= 5e-3 abs,  I_delta <= 4e-3 abs
  loss terms     <= 2e-2 relative, except L_I_smooth_delta <= 5e-2 relative: that term averages |forward differences|
                 of I_delta (~1e-3), so per-pixel rounding noise of the activations feeding final_conv biases it upwards.
                 With plain bf16 storage it is +17 % (the reference under CPU bf16 autocast: +16 %); the CUDA
                 path therefore stores the four full-resolution tensors on that path as bf16 hi+lo pairs
  gradients      cosine >= 0.995 over the full 1.14M-vector and per-tensor cosine >= 0.97
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

LOSS_RTOL = {"L_I_smooth_delta": 0.02}      # measured 0.55 % at B=2 x 128 x 128 (JYU weights): the term is weighted x2000


def _coefs():
    from oracle import sshslie_oracle as O
    return {"jyu": O.JYU_COEF, "cv": O.DEFAULT_COEF}


def _model(coef, seed=41, force_simt=False, graph=False):
    import sshslie_b200 as S
    torch.manual_seed(seed)
    m = S.LowLightEnhance(input_channels=64, lr=1e-3, **coef).to("cuda")
    m.force_simt = force_simt
    m.use_cuda_graph = graph
    return m


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


@pytest.mark.parametrize("impl", ["simt", "tcgen05"])
@pytest.mark.parametrize("shape", [(2, 32), (1, 64), (2, 128)])
def test_forward_matches_oracle(shape, impl):
    from oracle import sshslie_oracle as O
    B, size = shape
    m = _model(O.JYU_COEF, force_simt=(impl == "simt"))
    x = O.synthetic_patches(B, 64, size, seed=41)
    with torch.no_grad():
        R, I, Id, S = m.forward(x.cuda())
    torch.cuda.synchronize()
    p = O.init_params(41)
    Rr, Ir, Idr, Sr = O.forward(p, x)
    assert (R.cpu() - Rr).abs().max() <= 5e-3
    assert (I.cpu() - Ir).abs().max() <= 5e-3
    assert (Id.cpu() - Idr).abs().max() <= 4e-3
    assert (S.cpu() - Sr).abs().max() <= 5e-3


@pytest.mark.parametrize("impl", ["simt", "tcgen05"])
@pytest.mark.parametrize("case", [(2, 32, "jyu"), (1, 64, "cv"), (2, 128, "jyu")])
def test_loss_and_grads_match_oracle(case, impl):
    """Two oracles (both oracle/sshslie_oracle.py, fp32 arithmetic):
      fp32    : the reference semantics.  Loss terms per LOSS_RTOL; full 1.14M-gradient cosine >= 0.995; every
                tensor carrying >= 1 % of the gradient norm has cosine >= 0.96 and norm within 15 %; at the full patch
                size every tensor carrying >= 1e-4 of the norm has its norm within 10 % (measured: 0.969 .. 1.024 - the
                decomposition bottleneck was 0.82 before the half-resolution skip, conv1's input and I got bf16 pairs).
      storage : the same computation with activations/gradients rounded to bf16 at the tensors the CUDA path
                stores in bf16 (q=bf16_storage).  This isolates implementation error from storage noise:
                loss terms within 2e-3, full cosine >= 0.9995, tensors with >= 0.5 % of the norm cosine >= 0.95
                (sign() of the L1 terms flips on rounding-level differences, so this is not bit-level either).
    Tensors whose fp32 gradient is pure noise (attn.k_linear.bias: softmax is shift invariant, |g| ~ 1e-13) must
    stay below 1e-8 of the total norm."""
    from oracle import sshslie_oracle as O
    B, size, cname = case
    coef = _coefs()[cname]
    m = _model(coef, force_simt=(impl == "simt"))
    x = O.synthetic_patches(B, 64, size, seed=41)
    m.optimizer.zero_grad()
    loss, losses = m.compute_loss(x.cuda())
    loss.backward()
    torch.cuda.synchronize()
    p = O.init_params(41)
    l32, g32, _ = O.loss_and_grads(p, x, coef)
    l16, g16, _ = O.loss_and_grads(p, x, coef, q=O.cuda_storage)
    for k in O.LOSS_KEYS:
        np.testing.assert_allclose(losses[k], l32[k], rtol=LOSS_RTOL.get(k, 2e-2), atol=1e-5, err_msg=k)
        # the storage emulation follows the CUDA path tensor by tensor (oracle.HI_LO_TENSORS); L_I_smooth_delta, a mean of
        # |forward differences| of ~1e-3, is the one term that sees the remaining accumulation-order differences
        np.testing.assert_allclose(losses[k], l16[k], rtol=2e-2 if k == "L_I_smooth_delta" else 2e-3, atol=1e-6,
                                   err_msg=k + " (storage-precision oracle)")
    np.testing.assert_allclose(float(loss.detach()), l32["total_loss"], rtol=2e-2)
    G = {k: prm.grad.detach().cpu() for k, prm in m.named_parameters()}
    cat = lambda d: torch.cat([d[k].flatten() for k in G])
    tot = float(cat(g32).norm())
    assert _cos(cat(G), cat(g32)) >= 0.995
    assert _cos(cat(G), cat(g16)) >= 0.9995
    for k in G:
        share = float(g32[k].norm()) / tot
        if share < 1e-8:
            assert float(G[k].norm()) / tot < 1e-8, k
            continue
        if share >= 0.01:
            assert _cos(G[k], g32[k]) >= 0.96, (k, _cos(G[k], g32[k]))
            np.testing.assert_allclose(float(G[k].norm()), float(g32[k].norm()), rtol=0.15, err_msg=k)
        if size >= 128 and impl == "tcgen05" and share >= 1e-4:
            np.testing.assert_allclose(float(G[k].norm()), float(g32[k].norm()), rtol=0.10, err_msg=k)
        if share >= 0.005:
            assert _cos(G[k], g16[k]) >= 0.95, (k, _cos(G[k], g16[k]))
        else:
            assert _cos(G[k], g16[k]) >= 0.7, (k, _cos(G[k], g16[k]))


def test_reference_fixture_full_size(golden_dir):
    """Same check against the fixture recorded from the UNMODIFIED reference (config_outdoor_jyu.yml, B=2, 128^2)."""
    from oracle import sshslie_oracle as O
    g = np.load(os.path.join(golden_dir, "jyu_b2_128.npz"))
    m = _model(O.JYU_COEF)
    x = O.synthetic_patches(2, 64, 128, seed=41)
    m.optimizer.zero_grad()
    loss, losses = m.compute_loss(x.cuda())
    loss.backward()
    for k in O.LOSS_KEYS:
        np.testing.assert_allclose(losses[k], float(g["loss/" + k]), rtol=LOSS_RTOL.get(k, 2e-2), atol=1e-5, err_msg=k)
    total_l2 = float(np.sqrt(sum(float(g["grad/" + k + "/l2"]) ** 2 for k, _ in m.named_parameters())))
    for k, prm in m.named_parameters():
        ref = g["grad/" + k + "/samples"]
        f = prm.grad.detach().reshape(-1).double().cpu()
        idx = torch.linspace(0, f.numel() - 1, min(64, f.numel())).long()
        got = f[idx].numpy()
        share = float(g["grad/" + k + "/l2"]) / total_l2
        if share >= 0.01:      # tensors that carry the gradient; the rest is bf16-storage noise (see test above)
            c = float(got @ ref / (np.linalg.norm(got) * np.linalg.norm(ref) + 1e-300))
            assert c >= 0.85, (k, c)       # 64 strided samples only: sampling noise on top of bf16-storage noise
    R, I, Id, S_ = m.last_outputs
    for nm, t, tol in [("R_low", R, 5e-3), ("I_low", I, 5e-3), ("I_delta", Id, 4e-3), ("S", S_, 5e-3)]:
        f = t.detach().reshape(-1).double().cpu()
        idx = torch.linspace(0, f.numel() - 1, 256).long()
        assert np.abs(f[idx].numpy() - g["out/" + nm + "/samples"]).max() <= tol, nm


def test_train_steps_track_oracle():
    """3 optimizer steps (zero_grad / compute_loss / backward / step, model.py:313-316), CUDA graph on:
    the loss trajectory follows the oracle's within 3 % and every weight stays within 3 * lr * steps."""
    from oracle import sshslie_oracle as O
    coef = O.DEFAULT_COEF
    m = _model(coef, graph=True)
    p = O.init_params(41)
    state = {}
    for step in range(4):
        x = O.synthetic_patches(1, 64, 64, seed=200 + step)
        m.optimizer.zero_grad()
        loss, losses = m.compute_loss(x.cuda())
        loss.backward()
        m.optimizer.step()
        ref_losses, grads, _ = O.loss_and_grads(p, x, coef)
        p = O.adam_step(p, grads, state, lr=1e-3)
        np.testing.assert_allclose(losses["total_loss"], ref_losses["total_loss"], rtol=3e-2, err_msg=f"step {step}")
    sd = m.state_dict()
    for k in p:
        assert (sd[k].cpu() - p[k]).abs().max() <= 3 * 1e-3 * 4 + 1e-6, k


def test_state_dict_roundtrip_and_checkpoint(tmp_path):
    import sshslie_b200 as S
    from oracle import sshslie_oracle as O
    m = _model(O.DEFAULT_COEF)
    x = O.synthetic_patches(1, 64, 32, seed=7).cuda()
    m.optimizer.zero_grad()
    loss, _ = m.compute_loss(x)
    loss.backward()
    m.optimizer.step()
    path = str(tmp_path / "ck.pth")
    m.save_checkpoint(path, 1)
    ck = torch.load(path)
    assert set(ck.keys()) == {"epoch", "model_state_dict", "optimizer_state_dict"}
    assert list(ck["model_state_dict"].keys()) == list(O.init_params(41).keys())
    m2 = _model(O.DEFAULT_COEF, seed=1)
    m2.load_checkpoint(path)
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    with torch.no_grad():
        o1 = m.forward(x)
        o1 = [t.clone() for t in o1]
        o2 = m2.forward(x)
    for a, b in zip(o1, o2):
        assert torch.equal(a, b)


def test_linearity_of_batch():
    """Size-independent property: every loss term is a mean over equal shards, so the gradient of a 2-patch batch is
    the mean of the two single-patch gradients (the basis of the data-parallel all-reduce)."""
    from oracle import sshslie_oracle as O
    coef = O.DEFAULT_COEF
    x = O.synthetic_patches(2, 64, 128, seed=5).cuda()
    m = _model(coef)
    grads = []
    for xs in (x, x[0:1], x[1:2]):
        m.optimizer.zero_grad()
        loss, _ = m.compute_loss(xs.contiguous())
        loss.backward()
        grads.append(torch.cat([p.grad.detach().flatten().clone() for p in m.parameters()]))
    mean12 = 0.5 * (grads[1] + grads[2])
    assert _cos(grads[0], mean12) >= 0.9995


def test_backward_autograd_semantics():
    """`loss.backward()` (the reference's call, model.py:315) hands over the finished gradients without the autograd
    engine; arithmetic on the loss still goes through autograd: (3 * loss).backward() gives 3x the gradients, and a
    second backward into existing .grad accumulates."""
    from oracle import sshslie_oracle as O
    m = _model(O.JYU_COEF)
    x = O.synthetic_patches(1, 64, 64, seed=5).cuda()
    m.optimizer.zero_grad()
    loss, _ = m.compute_loss(x)
    loss.backward()
    g1 = torch.cat([p.grad.detach().flatten() for p in m.parameters()]).clone()
    m.optimizer.zero_grad()
    loss, _ = m.compute_loss(x)
    (3.0 * loss).backward()
    g3 = torch.cat([p.grad.detach().flatten() for p in m.parameters()]).clone()
    # (two runs of the same step differ in the last bits of a few entries: the attention / final-conv weight gradients
    # are accumulated with fp32 atomics)
    tol = 1e-5 * float(g1.abs().max())
    torch.testing.assert_close(g3, 3.0 * g1, rtol=5e-3, atol=tol)
    loss2, _ = m.compute_loss(x)
    loss2.backward()                       # accumulates into the existing .grad like autograd would
    g4 = torch.cat([p.grad.detach().flatten() for p in m.parameters()])
    torch.testing.assert_close(g4, 4.0 * g1, rtol=5e-3, atol=tol)


def _rect_input(B, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    x = 0.12 * torch.nn.functional.interpolate(torch.rand(B, 64, 9, 9, generator=g), size=(H, W), mode="bicubic",
                                               align_corners=False) + 0.01 * torch.rand(B, 64, H, W, generator=g)
    x = x.clamp(0, 1)
    return x / x.amax(dim=(1, 2, 3), keepdim=True)


@pytest.mark.parametrize("shape", [(1, 40, 56), (3, 64, 128), (1, 24, 16)])
def test_forward_rectangular_and_ragged_tokens(shape):
    """Edge cases of the tiling: H != W, token counts that are not multiples of the attention kernels' 16-token blocks
    (40x56 -> L = 35, 24x16 -> L = 6), odd batch."""
    from oracle import sshslie_oracle as O
    B, H, W = shape
    m = _model(O.JYU_COEF)
    x = _rect_input(B, H, W, 3)
    with torch.no_grad():
        R, I, Id, S = m.forward(x.cuda())
    torch.cuda.synchronize()
    Rr, Ir, Idr, Sr = O.forward(O.init_params(41), x)
    assert (R.cpu() - Rr).abs().max() <= 5e-3
    assert (I.cpu() - Ir).abs().max() <= 5e-3
    assert (Id.cpu() - Idr).abs().max() <= 4e-3
    assert (S.cpu() - Sr).abs().max() <= 5e-3


@pytest.mark.parametrize("shape", [(2, 34, 18), (1, 130, 126), (1, 500, 500)])
def test_forward_any_even_size(shape):
    """Any even H, W: the pyramid levels are ceil(n/2) (model.py:127-129), so 130x126 runs at
    65x63 / 33x32 / 17x16 and every F.interpolate(size=skip.shape, mode='nearest') (model.py:156-169) has a non-integer
    ratio; tiles are partial along both axes.  500x500 (L = 63 x 63 = 3969 tokens) takes the tensor-core attention."""
    from oracle import sshslie_oracle as O
    B, H, W = shape
    m = _model(O.JYU_COEF)
    x = _rect_input(B, H, W, 5)
    with torch.no_grad():
        R, I, Id, S = m.forward(x.cuda())
    torch.cuda.synchronize()
    assert R.shape == (B, 64, H, W) and I.shape == (B, 1, H, W) and Id.shape == (B, 1, H, W)
    Rr, Ir, Idr, Sr = O.forward(O.init_params(41), x)
    assert (R.cpu() - Rr).abs().max() <= 5e-3
    assert (I.cpu() - Ir).abs().max() <= 5e-3
    assert (Id.cpu() - Idr).abs().max() <= 4e-3
    assert (S.cpu() - Sr).abs().max() <= 5e-3


@pytest.mark.parametrize("shape", [(2, 64, 64), (1, 40, 56)])
def test_sub_networks_are_callable(shape):
    """`decomposition_net(x) -> (R, L)` and `illum_adjust_net(I, R) -> I_delta` on their own (model.py:49-70, 143-175), the
    second on caller-provided inputs that did NOT come from the decomposition net."""
    from oracle import sshslie_oracle as O
    B, H, W = shape
    m = _model(O.JYU_COEF)
    p = O.init_params(41)
    x = _rect_input(B, H, W, 21)
    with torch.no_grad():
        R, I = m.decomposition_net(x.cuda())
    Rr, Ir = O.decomposition_net(p, x)
    assert (R.cpu() - Rr).abs().max() <= 5e-3 and (I.cpu() - Ir).abs().max() <= 5e-3
    g = torch.Generator().manual_seed(5)
    Rin = torch.rand(B, 64, H, W, generator=g)
    Iin = torch.rand(B, 1, H, W, generator=g)
    with torch.no_grad():
        Id = m.illum_adjust_net(Iin.cuda(), Rin.cuda())
    Idr = O.illum_adjust_net(p, Iin, Rin)
    assert Id.shape == (B, 1, H, W)
    assert (Id.cpu() - Idr).abs().max() <= 4e-3 * max(1.0, float(Idr.abs().max()))


def test_odd_size_is_refused():
    """The reference itself cannot run an odd image (deconv output 2*ceil(H/2) != H, torch.cat fails, model.py:55-58)."""
    from oracle import sshslie_oracle as O
    import sshslie_b200 as S
    m = _model(O.JYU_COEF)
    with pytest.raises(S.lib.SshslieError):
        with torch.no_grad():
            m.forward(torch.rand(1, 64, 33, 32, device="cuda"))


def test_loss_and_grads_rectangular_odd_batch():
    """Training step on a non-square power-of-two patch with an odd batch (B=3, 64x128): losses and gradients vs the oracle."""
    from oracle import sshslie_oracle as O
    coef = O.DEFAULT_COEF
    m = _model(coef)
    x = _rect_input(3, 64, 128, 9)
    m.optimizer.zero_grad()
    loss, losses = m.compute_loss(x.cuda())
    loss.backward()
    ref, ref_g, _ = O.loss_and_grads(O.init_params(41), x, coef)
    for k in O.LOSS_KEYS:
        np.testing.assert_allclose(losses[k], ref[k], rtol=LOSS_RTOL.get(k, 2e-2), atol=1e-5, err_msg=k)
    g = torch.cat([p.grad.detach().flatten().cpu() for p in m.parameters()])
    r = torch.cat([v.flatten() for v in ref_g.values()])
    assert _cos(g, r) >= 0.995


@pytest.mark.parametrize("shape", [(2, 96, 96), (1, 256, 256), (1, 72, 120)])
def test_train_step_other_patch_sizes(shape):
    """patch_size is a free config value (config/*.yml:12): 96 and 72x120 are not powers of two and 256 does not fit the
    shared-memory FFT, so the Fourier term runs through the DFT path; 256 also puts the transformer block at 1024 tokens
    (tensor-core attention forward, its large-L backward).  Losses and the full gradient vs the oracle."""
    from oracle import sshslie_oracle as O
    B, H, W = shape
    coef = O.JYU_COEF
    m = _model(coef)
    x = _rect_input(B, H, W, 11)
    m.optimizer.zero_grad()
    loss, losses = m.compute_loss(x.cuda())
    loss.backward()
    torch.cuda.synchronize()
    torch.set_num_threads(os.cpu_count() or 1)
    ref, ref_g, _ = O.loss_and_grads(O.init_params(41), x, coef)
    for k in O.LOSS_KEYS:
        np.testing.assert_allclose(losses[k], ref[k], rtol=LOSS_RTOL.get(k, 2e-2), atol=1e-5, err_msg=k)
    g = torch.cat([p.grad.detach().flatten().cpu() for p in m.parameters()])
    r = torch.cat([v.flatten() for v in ref_g.values()])
    assert _cos(g, r) >= 0.995
    m.optimizer.step()                                   # and the step itself runs
    torch.cuda.synchronize()


@pytest.mark.parametrize("case", ["cv_b1_128", "jyu_b2_32_trained", "cv_b1_64_trained"])
def test_reference_fixtures_other_configs(case, golden_dir):
    """The remaining fixtures recorded from the UNMODIFIED reference: the Li-et-al loss weights at full size and two sets
    of TRAINED weights (non-trivial biases / heads; the weights are re-derived with the oracle's Adam, which
    tests/test_oracle_golden.py pins to the same fixture)."""
    from oracle import sshslie_oracle as O
    g = np.load(os.path.join(golden_dir, case + ".npz"))
    batch, size, pre = int(g["meta/batch"]), int(g["meta/size"]), int(g["meta/pre_steps"])
    coef = _coefs()[str(g["meta/coef"])]
    torch.set_num_threads(os.cpu_count() or 1)
    p = O.init_params(41)
    state = {}
    for s in range(pre):
        xs = O.synthetic_patches(batch, 64, size, seed=100 + s)
        _, grads, _ = O.loss_and_grads(p, xs, coef)
        p = O.adam_step(p, grads, state, lr=1e-3)
    m = _model(coef)
    m.load_state_dict({k: v.clone() for k, v in p.items()})
    x = O.synthetic_patches(batch, 64, size, seed=41)
    m.optimizer.zero_grad()
    loss, losses = m.compute_loss(x.cuda())
    loss.backward()
    for k in O.LOSS_KEYS:
        np.testing.assert_allclose(losses[k], float(g["loss/" + k]), rtol=LOSS_RTOL.get(k, 2e-2), atol=1e-5, err_msg=k)
    R, I, Id, S_ = m.last_outputs
    for nm, t, tol in [("R_low", R, 5e-3), ("I_low", I, 5e-3), ("I_delta", Id, 4e-3), ("S", S_, 5e-3)]:
        f = t.detach().reshape(-1).double().cpu()
        idx = torch.linspace(0, f.numel() - 1, 256).long()
        assert np.abs(f[idx].numpy() - g["out/" + nm + "/samples"]).max() <= tol, nm
    total_l2 = float(np.sqrt(sum(float(g["grad/" + k + "/l2"]) ** 2 for k, _ in m.named_parameters())))
    got_l2 = float(torch.cat([q.grad.detach().flatten() for q in m.parameters()]).double().norm())
    np.testing.assert_allclose(got_l2, total_l2, rtol=0.05)


def test_engine_cache_is_bounded():
    """test_model over cubes of many sizes must not accumulate workspaces: at most `max_cached_engines` shapes stay bound,
    and a shape that was evicted is rebuilt transparently."""
    from oracle import sshslie_oracle as O
    m = _model(O.JYU_COEF)
    first = None
    with torch.no_grad():
        for H, W in [(32, 32), (32, 48), (48, 32), (40, 40), (64, 32), (32, 64), (32, 32)]:
            x = _rect_input(1, H, W, 5)
            R, I, Id, S = m.forward(x.cuda())
            if (H, W) == (32, 32):
                if first is None:
                    first = S.clone()
                else:
                    assert torch.equal(first, S)          # rebuilt engine, same deterministic forward
            assert len(m._engines) <= m.max_cached_engines
    assert len(m._engines) == m.max_cached_engines


def _flat_grads(m):
    return torch.cat([p.grad.detach().flatten() for p in m.parameters()]).clone()


def test_gradient_accumulation_over_two_batches():
    """Micro-batch accumulation exactly as autograd does it (backward; compute_loss(x2); backward -> g1 + g2): the views
    of the flat gradient buffer handed out by the first backward must survive the second compute_loss overwriting it."""
    from oracle import sshslie_oracle as O
    m = _model(O.JYU_COEF)
    x1 = O.synthetic_patches(1, 64, 64, seed=5).cuda()
    x2 = O.synthetic_patches(1, 64, 64, seed=6).cuda()
    m.optimizer.zero_grad()
    l1, _ = m.compute_loss(x1)
    l1.backward()
    g1 = _flat_grads(m)
    m.optimizer.zero_grad()
    l2, _ = m.compute_loss(x2)
    l2.backward()
    g2 = _flat_grads(m)
    assert float((g1 - g2).abs().max()) > 1e-3 * float(g1.abs().max())           # the batches really differ
    m.optimizer.zero_grad()
    la, _ = m.compute_loss(x1)
    la.backward()
    lb, _ = m.compute_loss(x2)            # overwrites the flat buffer the first gradients were views of
    lb.backward()
    torch.testing.assert_close(_flat_grads(m), g1 + g2, rtol=1e-5, atol=1e-6 * float(g1.abs().max()))


def test_zero_grad_set_to_none_false_does_not_double():
    """optimizer.zero_grad(set_to_none=False) keeps p.grad = view of the flat buffer; the next step must give g, not 2 g."""
    from oracle import sshslie_oracle as O
    m = _model(O.JYU_COEF)
    x = O.synthetic_patches(1, 64, 64, seed=5).cuda()
    m.optimizer.zero_grad()
    l, _ = m.compute_loss(x)
    l.backward()
    g = _flat_grads(m)
    for _ in range(3):
        m.optimizer.zero_grad(set_to_none=False)
        l, _ = m.compute_loss(x)
        l.backward()
        torch.testing.assert_close(_flat_grads(m), g, rtol=0, atol=0)               # bit-repeatable as well


def test_autograd_grad_on_the_loss():
    """torch.autograd.grad(loss, params) works as on the reference's loss tensor (the loss node is wired to all 46
    parameters) and leaves .grad alone."""
    from oracle import sshslie_oracle as O
    m = _model(O.JYU_COEF)
    x = O.synthetic_patches(1, 64, 64, seed=5).cuda()
    m.optimizer.zero_grad()
    l, _ = m.compute_loss(x)
    l.backward()
    g = _flat_grads(m)
    m.optimizer.zero_grad()
    l, _ = m.compute_loss(x)
    gs = torch.autograd.grad(l, list(m.parameters()))
    assert all(p.grad is None for p in m.parameters())
    torch.testing.assert_close(torch.cat([t.flatten() for t in gs]), g, rtol=0, atol=0)


def test_loss_weights_changed_after_graph_capture():
    """The captured CUDA graph bakes the loss weights into kernel arguments; changing model.c_loss_* afterwards must take
    effect on the next step (the graph is re-captured), not be silently ignored."""
    from oracle import sshslie_oracle as O
    m = _model(O.JYU_COEF, graph=True)
    x = O.synthetic_patches(1, 64, 64, seed=5).cuda()
    for _ in range(4):                                    # steps 3+ replay the graph
        m.optimizer.zero_grad()
        l, losses = m.compute_loss(x)
        l.backward()
    t_jyu = losses["total_loss"]
    m.c_loss_i_smooth_delta, m.c_loss_fourier = 20.0, 0.2     # the cv1 weights
    m.optimizer.zero_grad()
    l, losses = m.compute_loss(x)
    l.backward()
    t_cv = losses["total_loss"]
    want = (10 * losses["L_reconstruction"] + losses["L_R_fidelity"] + losses["L_I_smooth_low"]
            + 20 * losses["L_I_smooth_delta"] + 0.2 * losses["L_fourier"] + losses["L_spectral_cons"])
    assert abs(t_cv - want) <= 1e-4 * abs(want) and abs(t_cv - t_jyu) > 1e-2 * abs(t_jyu)
    m2 = _model(O.DEFAULT_COEF)
    m2.optimizer.zero_grad()
    l2, _ = m2.compute_loss(x)
    l2.backward()
    torch.testing.assert_close(_flat_grads(m), _flat_grads(m2), rtol=0, atol=0)


def test_forward_returns_fresh_tensors():
    """The reference returns new tensors from every forward(); a caller holding R across two calls must not see it
    overwritten."""
    from oracle import sshslie_oracle as O
    m = _model(O.JYU_COEF)
    xa = O.synthetic_patches(1, 64, 64, seed=1).cuda()
    xb = O.synthetic_patches(1, 64, 64, seed=2).cuda()
    with torch.no_grad():
        Ra, _, _, Sa = m.forward(xa)
        keep = Ra.clone()
        Rb, _, _, Sb = m.forward(xb)
    torch.cuda.synchronize()
    assert Ra.data_ptr() != Rb.data_ptr() and torch.equal(Ra, keep) and not torch.equal(Ra, Rb)


def test_losses_read_late_keep_their_step():
    """The loss dict is filled from an async copy made right behind its own step: reading it after LATER steps were launched
    returns that step's values (what bench.py's e2e loop relies on)."""
    from oracle import sshslie_oracle as O
    xs = [O.synthetic_patches(2, 64, 32, seed=50 + i).cuda() for i in range(3)]
    m = _model(O.JYU_COEF, graph=True)
    eager = []
    for x in xs:                                    # read immediately
        m.optimizer.zero_grad()
        _, l = m.compute_loss(x)
        eager.append(l.copy())
    late = []
    for x in xs:                                    # launch all three, read afterwards
        m.optimizer.zero_grad()
        _, l = m.compute_loss(x)
        late.append(l)
    torch.cuda.synchronize()
    for a, b in zip(eager, late):
        assert a == b.copy()
    assert eager[0]["total_loss"] != eager[1]["total_loss"]


def test_step_is_bit_repeatable():
    """The reference runs with cudnn.deterministic=True (main.py:165): two runs of the same step give bit-identical
    losses and gradients (no floating-point atomics anywhere on the training path)."""
    from oracle import sshslie_oracle as O
    outs = []
    for _ in range(2):
        m = _model(O.JYU_COEF)
        x = O.synthetic_patches(2, 64, 128, seed=41).cuda()
        m.optimizer.zero_grad()
        l, losses = m.compute_loss(x)
        l.backward()
        outs.append((_flat_grads(m), [losses[k] for k in O.LOSS_KEYS]))
    assert outs[0][1] == outs[1][1]
    assert torch.equal(outs[0][0], outs[1][0])


@pytest.mark.parametrize("staged", ["1", "0"], ids=["tma_store", "direct_store"])
def test_model_parity_with_pipelined_kernels(staged, monkeypatch):
    """The persistent pipelined gather kernel (conv_pipe.cu) takes over only from 1024 tiles per layer; here it is forced
    onto every stride-1 layer of the B=2, 128x128 step (incl. the 9x9 layer, the sigmoid head with its fp32 TMA tile
    store, split / residual / hi+lo epilogues) and must give the same parity as the halo kernels."""
    from oracle import sshslie_oracle as O
    monkeypatch.setenv("SSHSLIE_PIPE_MIN_TILES", "0")
    monkeypatch.setenv("SSHSLIE_PIPE_MAX_SLABS", "81")
    monkeypatch.setenv("SSHSLIE_PIPE_STAGED", staged)
    # stride-2 convs (parity views) and transposed convs / strided data gradients (one launch per output parity class) also
    # move to the halo-reuse kernels from 512 tiles per class: force that too
    monkeypatch.setenv("SSHSLIE_S2_MIN_TILES", "0")
    m = _model(O.JYU_COEF)
    x = O.synthetic_patches(2, 64, 128, seed=41)
    with torch.no_grad():
        R, I, Id, S = m.forward(x.cuda())
    p = O.init_params(41)
    Rr, Ir, Idr, Sr = O.forward(p, x)
    assert (R.cpu() - Rr).abs().max() <= 5e-3 and (I.cpu() - Ir).abs().max() <= 5e-3
    assert (Id.cpu() - Idr).abs().max() <= 4e-3 and (S.cpu() - Sr).abs().max() <= 5e-3
    m.optimizer.zero_grad()
    loss, losses = m.compute_loss(x.cuda())
    loss.backward()
    l32, g32, _ = O.loss_and_grads(p, x, O.JYU_COEF)
    for k in O.LOSS_KEYS:
        np.testing.assert_allclose(losses[k], l32[k], rtol=LOSS_RTOL.get(k, 2e-2), atol=1e-5, err_msg=k)
    G = torch.cat([prm.grad.detach().flatten().cpu() for prm in m.parameters()])
    Gr = torch.cat([g32[k].flatten() for k, _ in m.named_parameters()])
    assert _cos(G, Gr) >= 0.995
    # and bit-identical to the halo-kernel path?  No: accumulation order inside a tile is the same, but the two kernels
    # need not agree bitwise - only repeatability of each path is required
    m.optimizer.zero_grad()
    loss2, _ = m.compute_loss(x.cuda())
    loss2.backward()
    G2 = torch.cat([prm.grad.detach().flatten().cpu() for prm in m.parameters()])
    assert torch.equal(G, G2)

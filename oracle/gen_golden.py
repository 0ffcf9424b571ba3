"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU (build container only).

    python oracle/gen_golden.py

For every case: seed-41 default-init reference `LowLightEnhance` (optionally advanced by a few
reference Adam steps so biases/heads are non-trivial), synthetic input from
`oracle.sshslie_oracle.synthetic_patches`, then the reference's own `compute_loss` + `backward`
(+ `optimizer.step`).  Recorded, small enough to commit:
  * the 7 loss floats (model.py:566-574),
  * for each of the 4 outputs + R_enh: sum, abs-sum, and values at fixed strided sample points,
  * for each of the 46 gradients: L2 norm, sum, and up to 64 strided samples,
  * for each of the 46 weights before and after one more Adam step: L2 norm + samples,
  * sample points are `torch.linspace(0, numel-1, n).long()` (recomputed by the tests, not stored);
  * the script asserts that `oracle.init_params(41)` equals the reference's seed-41 init bit for bit.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402
from oracle import sshslie_oracle as O  # noqa: E402

CASES = {
    # name: (batch, size, coef-set, pre-steps of reference Adam before recording)
    "jyu_b2_128": dict(batch=2, size=128, coef="jyu", pre_steps=0),
    "cv_b1_128": dict(batch=1, size=128, coef="cv", pre_steps=0),
    "jyu_b2_32_trained": dict(batch=2, size=32, coef="jyu", pre_steps=3),
    "cv_b1_64_trained": dict(batch=1, size=64, coef="cv", pre_steps=2),
}
COEFS = {"jyu": O.JYU_COEF, "cv": O.DEFAULT_COEF}


def sample(t: torch.Tensor, n=64):
    f = t.detach().reshape(-1).double()
    idx = torch.linspace(0, f.numel() - 1, min(n, f.numel())).long()
    return f[idx].numpy(), idx.numpy()


def stats(prefix, t, out, n=64):
    f = t.detach().double()
    vals, idx = sample(t, n)
    out[prefix + "/sum"] = np.float64(f.sum())
    out[prefix + "/abssum"] = np.float64(f.abs().sum())
    out[prefix + "/l2"] = np.float64(f.norm())
    out[prefix + "/samples"] = vals


def run_case(name, cfg, refmodel):
    torch.manual_seed(41)                       # main.py:160-164
    coef = COEFS[cfg["coef"]]
    m = refmodel.LowLightEnhance(input_channels=64, lr=1e-3, **coef)
    out = {}
    init = O.init_params(41)
    sd0 = m.state_dict()
    assert list(sd0.keys()) == list(init.keys()), "state_dict key order differs from oracle PARAM_SPECS"
    for k in sd0:
        assert torch.equal(sd0[k], init[k]), f"oracle init_params differs from reference init at {k}"
    x = O.synthetic_patches(cfg["batch"], 64, cfg["size"], seed=41)
    for s in range(cfg["pre_steps"]):
        xs = O.synthetic_patches(cfg["batch"], 64, cfg["size"], seed=100 + s)
        m.optimizer.zero_grad()
        loss, _ = m.compute_loss(xs)
        loss.backward()
        m.optimizer.step()
    # Adam state does not travel in the fixture: restart the optimizer like a fresh run on these weights
    m.optimizer = torch.optim.Adam(m.parameters(), lr=1e-3)
    for k, v in m.state_dict().items():
        stats("w0/" + k, v, out)
    m.optimizer.zero_grad()
    loss, losses = m.compute_loss(x)            # model.py:314
    loss.backward()                             # model.py:315
    for k in O.LOSS_KEYS:
        out["loss/" + k] = np.float64(losses[k])
    with torch.no_grad():
        R, I, Id, S = m.forward(x)
        Re, _ = m.decomposition_net(S)
    for nm, t in [("R_low", R), ("I_low", I), ("I_delta", Id), ("S", S), ("R_enh", Re)]:
        stats("out/" + nm, t, out, n=256)
    for k, p in m.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        stats("grad/" + k, g, out)
    m.optimizer.step()                          # model.py:316
    for k, v in m.state_dict().items():
        stats("w1/" + k, v, out)
    out["meta/batch"] = np.int64(cfg["batch"])
    out["meta/size"] = np.int64(cfg["size"])
    out["meta/pre_steps"] = np.int64(cfg["pre_steps"])
    out["meta/coef"] = np.array(cfg["coef"])
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: total_loss={losses['total_loss']:.6f} -> {path} ({os.path.getsize(path)} bytes)")
    return m


def main():
    torch.set_num_threads(os.cpu_count())
    refmodel = ref_shims.import_reference_model()
    for name, cfg in CASES.items():
        run_case(name, cfg, refmodel)


if __name__ == "__main__":
    main()

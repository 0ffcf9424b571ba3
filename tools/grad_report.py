"""Diagnostic (GPU box): per-tensor gradient agreement of the CUDA path with the fp32 oracle and with the oracle run
at the CUDA path's storage precision (bf16 activations / gradients).  Writes gpurun_out/grad_report.txt."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sshslie_b200 as S  # noqa: E402
from oracle import sshslie_oracle as O  # noqa: E402


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(a @ b / (a.norm() * b.norm() + 1e-300))


def main():
    B, size, cname = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    simt = len(sys.argv) > 4 and sys.argv[4] == "simt"
    coef = {"jyu": O.JYU_COEF, "cv": O.DEFAULT_COEF}[cname]
    torch.manual_seed(41)
    m = S.LowLightEnhance(input_channels=64, lr=1e-3, **coef).to("cuda")
    m.use_cuda_graph = False
    m.force_simt = simt
    x = O.synthetic_patches(B, 64, size, seed=41)
    m.optimizer.zero_grad()
    loss, losses = m.compute_loss(x.cuda())
    loss.backward()
    torch.cuda.synchronize()
    p = O.init_params(41)
    l32, g32, _ = O.loss_and_grads(p, x, coef)
    l16, g16, _ = O.loss_and_grads(p, x, coef, q=O.cuda_storage)
    out = []
    out.append(f"case B={B} size={size} coef={cname} simt={simt}")
    for k in O.LOSS_KEYS:
        out.append(f"{k:20s} cuda={losses[k]:.6f} fp32={l32[k]:.6f} bf16emu={l16[k]:.6f}")
    G = {k: prm.grad.detach().cpu() for k, prm in m.named_parameters()}
    tot = torch.cat([v.flatten() for v in g32.values()]).norm()
    out.append(f"{'tensor':48s} {'|g32|/tot':>9s} {'cos(cuda,32)':>12s} {'cos(cuda,emu)':>13s} {'cos(emu,32)':>11s} {'|cuda|/|32|':>11s}")
    for k in G:
        n32 = float(g32[k].norm())
        out.append(f"{k:48s} {n32/float(tot):9.4f} {cos(G[k], g32[k]):12.4f} {cos(G[k], g16[k]):13.4f} "
                   f"{cos(g16[k], g32[k]):11.4f} {float(G[k].norm())/(n32+1e-300):11.4f}")
    cat = lambda d: torch.cat([d[k].flatten() for k in G])
    out.append(f"FULL cos(cuda,fp32)={cos(cat(G), cat(g32)):.5f} cos(cuda,emu)={cos(cat(G), cat(g16)):.5f} "
               f"cos(emu,fp32)={cos(cat(g16), cat(g32)):.5f}")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"grad_report_{B}_{size}_{cname}{'_simt' if simt else ''}.txt"), "w") as f:
        f.write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()

"""GPU box: cycles per tcgen05.mma (M=128, K=16, bf16) as a function of N, accumulator interleave and commit cadence."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sshslie_b200 as S  # noqa: E402

lib = S.lib.load()
out = torch.zeros(148, dtype=torch.int64, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
print("AOFF", os.environ.get("SSHSLIE_PROBE_AOFF"), "SBO", os.environ.get("SSHSLIE_PROBE_SBO"), "MN", os.environ.get("SSHSLIE_PROBE_MN"), "DISTINCT", os.environ.get("SSHSLIE_PROBE_DISTINCT"))
print(f"{'N':>4s} {'n_acc':>5s} {'commit_every':>12s} {'ctas':>5s} {'cycles/MMA':>10s} {'ideal(N/2)':>10s}")
for ctas in (148,):
    for N in (64, 128):
        for n_acc in (1, 4):
            if n_acc * N > 512:
                continue
            for ce in (1,):
                n = 2048
                for _ in range(2):
                    S.lib.check(lib.sshslie_umma_probe(N, n, n_acc, ce, S.lib.ptr(out), ctas, st), "probe")
                torch.cuda.synchronize()
                cyc = out[:ctas].double().mean().item() / n
                print(f"{N:4d} {n_acc:5d} {ce:12d} {ctas:5d} {cyc:10.1f} {N/2:10.1f}")

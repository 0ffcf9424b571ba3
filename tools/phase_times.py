"""GPU box: device time of CUDA-graph replays of (a) forward only, (b) phase 1 (fwd+loss+pass-2 bwd+illum bwd),
(c) phase 2 (pass-1 decomposition bwd), (d) both; plus the same eagerly.  Prints ms per call."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sshslie_b200 as S  # noqa: E402
from oracle import sshslie_oracle as O  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
size = int(sys.argv[2]) if len(sys.argv) > 2 else 128
torch.manual_seed(41)
m = S.LowLightEnhance(input_channels=64, lr=1e-3, **O.JYU_COEF).to("cuda")
m.use_cuda_graph = False
x = O.synthetic_patches(B, 64, size, seed=41).cuda()
m._ensure_flat()
eng = m._engine(x, train=True)
engf = m._engine(x, train=False)
m._stage_input(eng, x)
m._stage_input(engf, x)
lib = S.lib.load()


def fwd():
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    S.lib.check(lib.sshslie_forward(engf.handle, S.lib.ptr(engf.x), S.lib.ptr(m._flat), None, None, None, None, st),
                "fwd")


def timeit(fn, n=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def graphed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g.replay


cases = [("forward", fwd), ("phase1", lambda: m._launch_loss_and_grad(eng, 1)),
         ("phase2", lambda: m._launch_loss_and_grad(eng, 2)), ("both", lambda: m._launch_loss_and_grad(eng, 3))]
print(f"B={B} size={size} PDL={os.environ.get('SSHSLIE_PDL', '1')}")
for name, fn in cases:
    te = timeit(fn)
    tg = timeit(graphed(fn))
    print(f"{name:8s} eager {te:.3f} ms   graph {tg:.3f} ms")

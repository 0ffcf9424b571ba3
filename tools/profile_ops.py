"""GPU box: per-launch-group device times of one training step (cudaEvent pairs, eager), averaged over a few steps.
Usage: python tools/profile_ops.py [B] [size] [out-file]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sshslie_b200 as S  # noqa: E402
from oracle import sshslie_oracle as O  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
TRAIN = os.environ.get("FORWARD_ONLY", "0") != "1"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 128
out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", f"profile_ops_b{B}_{size}.txt")
torch.manual_seed(41)
m = S.LowLightEnhance(input_channels=64, lr=1e-3, **O.JYU_COEF).to("cuda")
x = O.synthetic_patches(B, 64, size, seed=41).cuda()
acc = {}
order = []
reps = 5
for r in range(reps + 1):
    rows = m.profile_step(x, train=TRAIN)
    if r == 0:
        continue
    for i, (name, ms, fl, by) in enumerate(rows):
        key = (i, name)
        if key not in acc:
            acc[key] = [0.0, fl, by]
            order.append(key)
        acc[key][0] += ms / reps
total = sum(v[0] for v in acc.values())
lines = [f"# B={B} {size}x{size}: {len(order)} launch groups, sum of device time {total:.3f} ms/step (eager, event-timed)",
         f"{'#':>3s} {'ms':>8s} {'share':>6s} {'TFLOP/s':>8s}  name"]
for key in order:
    ms, fl, by = acc[key]
    tf = fl / (ms * 1e-3) / 1e12 if fl and ms > 0 else 0
    lines.append(f"{key[0]:3d} {ms:8.4f} {100*ms/total:5.1f}% {tf:8.1f}  {key[1]}")
os.makedirs(os.path.dirname(out), exist_ok=True)
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))

"""torchrun helper (2+ GPUs): data-parallel gradient == oracle gradient on the concatenated global batch.
    torchrun --nproc-per-node 2 tests/dp_check.py
Every loss term is a mean over equal shards, so mean-of-rank-gradients is exact."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sshslie_b200 as S  # noqa: E402
from oracle import sshslie_oracle as O  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
coef = O.DEFAULT_COEF                     # config_indoor_li_et_al_cv1.yml weights, per-rank batch 1
torch.manual_seed(41)
m = S.LowLightEnhance(lr=1e-3, **coef).to("cuda")
m.enable_data_parallel()
xs = [O.synthetic_patches(1, 64, 128, seed=41 + r) for r in range(world)]
for it in range(4):                        # iterations 3+ run through the two captured graphs
    m.optimizer.zero_grad()
    loss, losses = m.compute_loss(xs[rank].cuda())
    loss.backward()
torch.cuda.synchronize()
g = torch.cat([p.grad.detach().flatten() for p in m.parameters()]).cpu().double()
if rank == 0:
    xg = torch.cat(xs, 0)
    ref_l, ref_g, _ = O.loss_and_grads(O.init_params(41), xg, coef)
    r = torch.cat([v.flatten() for v in ref_g.values()]).double()
    cos = float(g @ r / (g.norm() * r.norm()))
    rel = abs(losses["total_loss"] - ref_l["total_loss"]) / ref_l["total_loss"]
    print(f"DP_CHECK world={world} grad_cos={cos:.5f} norm_ratio={float(g.norm()/r.norm()):.4f} loss_rel={rel:.2e}")
    assert cos >= 0.995 and rel < 2e-2
# all ranks must hold identical averaged gradients
t = torch.cat([p.grad.detach().flatten() for p in m.parameters()])
lo, hi = t.clone(), t.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN)
dist.all_reduce(hi, op=dist.ReduceOp.MAX)
assert torch.equal(lo, hi), "ranks disagree on the averaged gradient"
m.optimizer.step()
dist.barrier()
torch.cuda.synchronize()
if rank == 0:
    print("DP_CHECK ok", flush=True)
sys.stdout.flush()
# leave without tearing the communicator down: the captured step graph still holds its NCCL kernels, and
# destroy_process_group / interpreter shutdown can block on that
os._exit(0)

"""world_size-2 CPU (gloo) test of the data-parallel host logic: bucket layout + all-reduce + mean, on the same code the
NCCL path runs (sshslie_b200/parallel.py), and bench.py's reference arm under a 2-rank launch."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sshslie_b200 as S
    from sshslie_b200 import parallel as P
    total, offs, sizes = S.lib.param_table(64)
    dec, ill = P.bucket_ranges(offs, sizes)
    assert dec == (0, 849121) and ill == (849121, 1141922)          # SURVEY.md §8e parameter split
    g = torch.Generator().manual_seed(100 + rank)
    # the layout LowLightEnhance._dp_step exchanges: gradients, then the 8 loss scalars in the same allocation, so that
    # the losses ride in the illum bucket's all-reduce and ONE kernel turns sums into means
    store = torch.cat([torch.randn(total, generator=g), torch.full((8,), float(rank + 1))])
    flat, losses = store[:total], store[total:]
    mine = flat.clone()
    P.allreduce_bucket(store, (ill[0], total + 8))                    # bucket 1 + losses first (ready mid-backward)
    assert torch.equal(flat[:dec[1]], mine[:dec[1]])                  # decomposition slice untouched so far
    P.allreduce_bucket(store, dec)
    P.finish_mean(store, None, world)
    others = [torch.randn(total, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
    ref = sum(others) / world
    torch.testing.assert_close(flat, ref, rtol=1e-6, atol=1e-6)
    assert abs(float(losses[0]) - 1.5) < 1e-6
    # what LowLightEnhance._dp_step does instead of the rescaling pass: every rank differentiates loss / world (loss weights
    # c_loss_* / world, parallel.dp_loss_weights), so the all-reduce(SUM) is the mean already.  The gradient is linear in
    # the weights; the logged total rides along pre-scaled, the six raw term values arrive as sums and are divided when
    # the host reads them (LazyLosses term_scale)
    from sshslie_b200.model import LazyLosses, LOSS_KEYS
    w = [10.0, 1.0, 1.0, 2000.0, 20.0, 1.0]
    ws = P.dp_loss_weights(w, world)
    terms = torch.tensor([0.1 * (rank + 1) * (k + 1) for k in range(6)])         # this rank's raw term values
    grad_k = [torch.randn(16, generator=torch.Generator().manual_seed(7 * r + k)) for r in range(world) for k in range(6)]
    my = sum(ws[k] * grad_k[rank * 6 + k] for k in range(6))                     # d(loss_r / world)
    store2 = torch.cat([my, torch.tensor([float(sum(ws[k] * terms[k] for k in range(6)))]), terms, torch.zeros(1)])
    dist.all_reduce(store2, op=dist.ReduceOp.SUM)
    want = sum(w[k] * grad_k[r * 6 + k] for r in range(world) for k in range(6)) / world
    torch.testing.assert_close(store2[:16], want, rtol=1e-5, atol=1e-5)
    ll = LazyLosses(store2[16:23].clone(), 1.0 / world)
    mean_terms = [sum(0.1 * (r + 1) * (k + 1) for r in range(world)) / world for k in range(6)]
    assert abs(ll["total_loss"] - sum(w[k] * mean_terms[k] for k in range(6))) < 1e-3
    for k in range(6):
        assert abs(ll[LOSS_KEYS[k + 1]] - mean_terms[k]) < 1e-6
    assert "total_loss" in ll and ll.get("L_fourier") is not None and len(ll.copy()) == 7
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_bucketed_allreduce_mean_gloo(tmp_path):
    mp.spawn(_worker, args=(2, 29531, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_bench_reference_arm_two_ranks():
    """`bench.py --impl reference` under a 2-rank launch: rank 0 alone runs and prints ONE JSON line."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "0"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "patches/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["steps"] == 1 and d["warmup"] == 0                      # the arm honours --steps / --warmup

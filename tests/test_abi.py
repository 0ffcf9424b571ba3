"""CPU-side checks of the C-ABI boundary (no GPU): the library loads, exports every symbol the header declares,
the flat parameter table matches the reference's state_dict, host-only entry points behave, errors are loud."""
import ctypes
import os
import re

import pytest
import torch

import sshslie_b200 as S
from oracle import sshslie_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "sshslie_b200.h")).read()
    declared = sorted(set(re.findall(r"SSHSLIE_API\s+[\w\s\*]+?\b(sshslie_\w+)\s*\(", header)))
    assert declared == sorted(S.lib.EXPORTS)
    lib = S.lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.sshslie_version() >= 100


def test_param_table_is_reference_state_dict_order():
    total, offs, sizes = S.lib.param_table(64)
    p = O.init_params(41)
    assert total == 1141922 == sum(v.numel() for v in p.values())
    assert sizes == [v.numel() for v in p.values()]
    assert offs == [sum(sizes[:i]) for i in range(len(sizes))]


def test_module_surface_matches_reference():
    torch.manual_seed(41)
    m = S.LowLightEnhance(input_channels=64, lr=1e-3, lr_update_factor=0.1, lr_update_period=250)
    sd = m.state_dict()
    ref = O.init_params(41)
    assert list(sd.keys()) == list(ref.keys())
    for k in ref:
        assert torch.equal(sd[k], ref[k]), k                      # same seed -> same init as the reference
    assert hasattr(m, "scheduler") and m.adaptive_lr
    for attr in ("decomposition_net", "illum_adjust_net", "optimizer", "freeze_decom_epochs", "eval_metrics",
                 "all_epoch_losses", "train_model", "test_model", "evaluate_model", "save_checkpoint",
                 "load_checkpoint", "compute_loss", "forward"):
        assert hasattr(m, attr), attr
    osd = m.optimizer.state_dict()
    assert set(osd.keys()) == {"state", "param_groups"} and osd["param_groups"][0]["lr"] == 1e-3


def test_engine_plan_is_host_only():
    lib = S.lib.load()
    h = ctypes.c_void_p()
    assert lib.sshslie_engine_create(ctypes.byref(h), 2, 64, 128, 128, S.lib.FLAG_TRAIN) == 0
    train_bytes = lib.sshslie_engine_workspace_bytes(h)
    lib.sshslie_engine_destroy(h)
    assert lib.sshslie_engine_create(ctypes.byref(h), 1, 64, 512, 512, 0) == 0
    infer_bytes = lib.sshslie_engine_workspace_bytes(h)
    lib.sshslie_engine_destroy(h)
    assert 50e6 < train_bytes < 2e9 and 100e6 < infer_bytes < 8e9
    # any even size plans for inference, any multiple of 8 for training (256: the Fourier term's DFT planes are in the workspace)
    for args in [(1, 64, 130, 126, 0), (1, 64, 500, 500, 0), (1, 64, 256, 256, S.lib.FLAG_TRAIN), (2, 64, 96, 72, S.lib.FLAG_TRAIN)]:
        assert lib.sshslie_engine_create(ctypes.byref(h), *args) == 0, args
        assert lib.sshslie_engine_workspace_bytes(h) > 0
        lib.sshslie_engine_destroy(h)


@pytest.mark.parametrize("args", [(2, 32, 128, 128, 0), (2, 64, 131, 128, 0), (0, 64, 128, 128, 0), (1, 64, 14, 16, 0),
                                  (1, 64, 132, 128, S.lib.FLAG_TRAIN), (1, 64, 2048, 128, S.lib.FLAG_TRAIN)])
def test_engine_rejects_unsupported_shapes(args):
    lib = S.lib.load()
    h = ctypes.c_void_p()
    assert lib.sshslie_engine_create(ctypes.byref(h), *args) == -1
    assert b"sshslie_engine_create" in lib.sshslie_last_error()


def test_cpu_tensors_fail_loudly():
    m = S.LowLightEnhance()
    with pytest.raises(S.lib.SshslieError):
        m.compute_loss(torch.rand(1, 64, 32, 32))
    with pytest.raises(S.lib.SshslieError):
        m.illum_adjust_net(torch.rand(1, 1, 16, 16), torch.rand(1, 64, 16, 16))


def test_bench_reference_arm_prints_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) needs no GPU and prints ONE JSON line with the
    contract keys; ours refuses to run without a CUDA device instead of falling back."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-1000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "train_patches_per_sec" and line["value"] > 0
    from oracle import make_ref
    want_kind = "reference" if make_ref.ref_dir() else "port"      # oracle/_ref is placed by __graft_entry__.build()
    assert line["cpu_baseline"]["kind"] == want_kind and line["cpu_baseline"]["cores"] >= 1
    assert line["steps"] == 1 and line["warmup"] == 1              # --steps / --warmup are honoured
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0
    if not torch.cuda.is_available():
        ours = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                              capture_output=True, text=True, timeout=600, cwd=ROOT)
        assert ours.returncode != 0 and "CUDA" in (ours.stderr + ours.stdout)

"""-m gpu: the reference's entry point contract end to end — `main.py --config config/config.yml --phase train` on synthetic
.mat cubes (BASELINE.json configs[0], run on the GPU because this implementation has no CPU path), then `--phase test`-style
inference through test_model on a 64x64 cube, and a 512x512 full-image forward (configs[3])."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_cubes(d, n, h, w, seed):
    import scipy.io as sio
    os.makedirs(d, exist_ok=True)
    rng = np.random.default_rng(seed)
    for i in range(n):
        yy, xx = np.mgrid[0:h, 0:w]
        scene = 0.5 + 0.5 * np.sin(xx / 23.0 + i) * np.cos(yy / 31.0)
        spec = 0.4 + 0.6 * rng.random(64)
        hi = 238 + 3000 * scene[..., None] * spec[None, None, :]
        low = 238 + 0.1 * (hi - 238) + rng.normal(0, 3, hi.shape)
        sio.savemat(os.path.join(d, f"cube_{i}.mat"), {"data": low.astype(np.float32)})


def test_main_train_one_epoch(tmp_path):
    data = tmp_path / "data" / "low"
    _write_cubes(str(data / "train"), 4, 160, 144, 41)
    os.makedirs(str(data / "eval"), exist_ok=True)
    env = dict(os.environ, PYTHONPATH=ROOT)
    cmd = [sys.executable, os.path.join(ROOT, "main.py"), "--config", os.path.join(ROOT, "config", "config.yml"),
           "--model_name", "outdoor", "--phase", "train", "--epoch", "2", "--eval_every_epoch", "1",
           "--train_data", str(data / "train"), "--eval_data", str(data / "eval")]
    out = subprocess.run(cmd, cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "Epoch [2/2] Batch [2/2]" in out.stdout and "Job finished" in out.stdout
    losses = [float(l.split("Loss:")[1]) for l in out.stdout.splitlines() if "Batch [" in l]
    assert len(losses) == 4 and all(np.isfinite(losses))
    ck = list((tmp_path / "checkpoint" / "outdoor").glob("Decomposition_*/model_epoch_latest.pth"))
    assert len(ck) == 1
    sd = torch.load(str(ck[0]), map_location="cpu")
    assert set(sd) == {"epoch", "model_state_dict", "optimizer_state_dict"} and len(sd["model_state_dict"]) == 46


def test_test_model_writes_outputs(tmp_path):
    import scipy.io as sio
    import sshslie_b200 as S
    from oracle import sshslie_oracle as O
    torch.manual_seed(41)
    m = S.LowLightEnhance(global_min=238.0, global_max=4095.0, time_stamp="t").to("cuda")
    ck_dir = tmp_path / "ck"
    os.makedirs(ck_dir)
    # one step so that the optimizer state exists, then checkpoint
    x = O.synthetic_patches(1, 64, 64, seed=3).cuda()
    m.optimizer.zero_grad()
    loss, _ = m.compute_loss(x)
    loss.backward()
    m.optimizer.step()
    m.save_checkpoint(str(ck_dir / "model_epoch_latest.pth"), 1)
    cube = O.synthetic_patches(1, 64, 64, seed=9)[0].permute(1, 2, 0).numpy()
    out_dir = tmp_path / "out"
    os.makedirs(out_dir)
    m.test_model(str(ck_dir), [cube], ["scene.mat"], str(out_dir), save_reflectance=True, save_illumination=True,
                 save_i_delta=True)
    S_np = sio.loadmat(str(out_dir / "scene.mat"))["data"]
    assert S_np.shape == (64, 64, 64) and np.isfinite(S_np).all()
    with torch.no_grad():
        _, _, _, S_ref = O.forward({k: v.cpu() for k, v in m.state_dict().items()}, torch.from_numpy(cube).permute(2, 0, 1)[None])
    ref = S_ref[0].permute(1, 2, 0).numpy() * (4095.0 - 238.0) + 238.0            # model.py:423-424
    assert np.abs(S_np - ref).max() <= 5e-3 * (4095.0 - 238.0)
    for suffix in ("_R_low", "_I_low", "_I_delta"):
        assert (out_dir / "artifacts" / f"scene{suffix}.mat").exists()


def test_full_image_inference_512():
    """BASELINE.json configs[3]: forward on a 1x64x512x512 cube (L = 4096 attention tokens), vs the oracle."""
    import sshslie_b200 as S
    from oracle import sshslie_oracle as O
    torch.manual_seed(41)
    m = S.LowLightEnhance().to("cuda")
    x = O.synthetic_patches(1, 64, 512, seed=41)
    with torch.no_grad():
        R, I, Id, S_ = m.forward(x.cuda())
    torch.cuda.synchronize()
    torch.set_num_threads(os.cpu_count() or 1)
    Rr, Ir, Idr, Sr = O.forward(O.init_params(41), x)
    assert (R.cpu() - Rr).abs().max() <= 5e-3
    assert (I.cpu() - Ir).abs().max() <= 5e-3
    assert (Id.cpu() - Idr).abs().max() <= 4e-3
    assert (S_.cpu() - Sr).abs().max() <= 5e-3


def test_evaluate_model_computes_metrics(tmp_path):
    """In-training evaluation (model.py:342-403): outputs are written and PSNR / SSIM / SAM against the label cubes land in
    `eval_metrics[epoch]`, equal to the oracle's restatement of the metrics on the written file."""
    import scipy.io as sio
    import sshslie_b200 as S
    from oracle import sshslie_oracle as O
    torch.manual_seed(41)
    gmin, gmax = 238.0, 4095.0
    m = S.LowLightEnhance(global_min=gmin, global_max=gmax, time_stamp="t").to("cuda")
    cube = O.synthetic_patches(1, 64, 64, seed=9)[0].permute(1, 2, 0).numpy()
    label_dir = tmp_path / "label"
    os.makedirs(label_dir)
    rng = np.random.default_rng(2)
    label = (gmin + (gmax - gmin) * np.clip(cube * 3.0 + 0.02 * rng.standard_normal(cube.shape), 0, 1)).astype(np.float32)
    sio.savemat(str(label_dir / "scene.mat"), {"data": label})
    out_dir = tmp_path / "eval"
    m.evaluate_model([cube], ["scene.mat"], str(out_dir), 3, str(label_dir))
    assert set(m.eval_metrics[3]) == {"psnr", "ssim", "sam"}
    pred = torch.from_numpy(sio.loadmat(str(out_dir / "epoch_3" / "scene.mat"))["data"])
    lab = torch.from_numpy(label)
    np.testing.assert_allclose(m.eval_metrics[3]["psnr"], float(O.psnr(pred, lab, gmax)), rtol=1e-4)
    np.testing.assert_allclose(m.eval_metrics[3]["sam"], float(O.sam(pred, lab)), rtol=1e-3)
    np.testing.assert_allclose(m.eval_metrics[3]["ssim"], float(O.ssim(pred, lab, gmax)), rtol=1e-3)

#!/usr/bin/env python
"""Entry point with the reference's CLI and config contract (reference main.py:16-90, 147-276):

    python main.py --config config/config.yml --model_name outdoor [--<any key> value ...]

Precedence CLI > YAML > built-in default; same keys as the reference's config/*.yml.  The model is the B200-native
drop-in (`sshslie_b200.LowLightEnhance`); experiment tracking (mlflow), torchinfo and plotting are optional extras
that are skipped when not installed.  Needs a CUDA device: there is no CPU path.
"""
import argparse
import os
import random
import sys
import traceback
from datetime import datetime
from glob import glob

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEFAULTS = dict(
    use_gpu=1, seed_value=41, gpu_idx='0', gpu_mem=0.8, decom=0, mat_key='data', channels=64, global_min=0.,
    global_max=1., normalization='global_normalization', batch_size=1, patch_size=128, start_lr=0.001,
    lr_update_factor=1, lr_update_period=400, train_data='./data/train/low', eval_data='./data/eval/low',
    test_data='./data/test/low', label_dir='./data/test/high', phase='train_and_test', epoch=400,
    eval_every_epoch=200, plot_every_epoch=200, c_loss_reconstruction=10., c_loss_r_fidelity=1.,
    c_loss_i_smooth_low=1., c_loss_i_smooth_delta=20., c_loss_fourier=0.2, c_loss_spectral_cons=1.,
    alpha_i_smooth_low=1., alpha_i_smooth_delta=10., save_reflectance=False, save_illumination=False,
    save_i_delta=False, model_name='no_name_model', pretrained_model='', freeze_decom_epochs=0,
    test_model_dir='')      # phase=test only: checkpoint directory to load ('' = the newest Decomposition_* of this model)


def parse_args(argv=None):
    ap = argparse.ArgumentParser(description="SS-HSLIE on B200: YAML config + command-line overrides")
    ap.add_argument('--config', type=str, default='./config/config.yml')
    for key, val in DEFAULTS.items():
        ap.add_argument(f'--{key}', type=type(val), default=None)
    args = ap.parse_args(argv)
    with open(args.config) as fh:
        cfg = yaml.safe_load(fh) or {}
    for key, val in DEFAULTS.items():
        if getattr(args, key) is None:
            setattr(args, key, cfg.get(key, val))
    args.timestamp = datetime.now().strftime('%Y%m%d_%H%M%S')
    args.full_model_name = f"{args.model_name}_{args.timestamp}"
    args.model_ckpt_dir = './checkpoint/' + args.model_name
    args.eval_result_dir = './results/eval_results_' + args.full_model_name
    args.test_result_dir = './results/test_results_' + args.full_model_name
    # the reference reads 'decomposition_<ts>' but writes 'Decomposition_<ts>' (SURVEY A.1); use the written name.
    # A standalone `--phase test` run has a fresh timestamp, so that directory cannot exist: take --test_model_dir, else
    # the newest Decomposition_* directory of this model.
    if not args.test_model_dir:
        args.test_model_dir = os.path.join(args.model_ckpt_dir, 'Decomposition_' + args.timestamp)
        if args.phase == 'test':
            found = sorted(glob(os.path.join(args.model_ckpt_dir, 'Decomposition_*')), key=os.path.getmtime)
            if found:
                args.test_model_dir = found[-1]
    return args


def build_model(args, device):
    from sshslie_b200 import LowLightEnhance
    return LowLightEnhance(
        input_channels=args.channels, lr=args.start_lr, lr_update_factor=args.lr_update_factor,
        lr_update_period=args.lr_update_period, time_stamp=args.timestamp,
        c_loss_reconstruction=args.c_loss_reconstruction, c_loss_r_fidelity=args.c_loss_r_fidelity,
        c_loss_i_smooth_low=args.c_loss_i_smooth_low, c_loss_i_smooth_delta=args.c_loss_i_smooth_delta,
        c_loss_fourier=args.c_loss_fourier, c_loss_spectral_cons=args.c_loss_spectral_cons,
        alpha_i_smooth_low=args.alpha_i_smooth_low, alpha_i_smooth_delta=args.alpha_i_smooth_delta, device=device,
        global_min=args.global_min, global_max=args.global_max, save_reflectance=args.save_reflectance,
        save_illumination=args.save_illumination, save_i_delta=args.save_i_delta).to(device)


def run_train(model, args):
    model.train_model(train_data_path=args.train_data, eval_data_path=args.eval_data, batch_size=args.batch_size,
                      patch_size=args.patch_size, num_epochs=args.epoch, start_lr=args.start_lr,
                      ckpt_dir=args.model_ckpt_dir, eval_result_dir=args.eval_result_dir,
                      eval_every_epoch=args.eval_every_epoch, label_dir=args.label_dir,
                      plot_every_epoch=args.plot_every_epoch)


def run_test(model, args):
    from sshslie_b200.utils import load_hsi
    os.makedirs(args.test_result_dir, exist_ok=True)
    names = sorted(glob(os.path.join(args.test_data, '*.*')))
    cubes = [load_hsi(n, matContentHeader=args.mat_key, normalization=args.normalization, max_val=args.global_max,
                      min_val=args.global_min) for n in names]
    model.test_model(model_dir=args.test_model_dir, test_low_data=cubes, test_low_data_names=names,
                     save_dir=args.test_result_dir, save_reflectance=args.save_reflectance,
                     save_illumination=args.save_illumination, save_i_delta=args.save_i_delta)


def main(args):
    random.seed(args.seed_value)
    np.random.seed(args.seed_value)
    torch.manual_seed(args.seed_value)
    if not (args.use_gpu and torch.cuda.is_available()):
        raise SystemExit("sshslie_b200 needs a CUDA device (use_gpu=1 on a B200); the hot path has no CPU implementation")
    idx = int(str(args.gpu_idx).split(',')[0] or 0)               # the reference sets CUDA_VISIBLE_DEVICES from gpu_idx
    device = torch.device(f'cuda:{idx}' if idx < torch.cuda.device_count() else 'cuda:0')
    torch.cuda.set_device(device)
    model = build_model(args, device)
    if args.pretrained_model and os.path.exists(args.pretrained_model):
        ckpt = torch.load(args.pretrained_model, map_location=device)
        model.load_state_dict(ckpt.get('model_state_dict', ckpt))
        model.freeze_decom_epochs = args.freeze_decom_epochs
    try:
        if args.phase in ('train', 'train_and_test'):
            run_train(model, args)
        if args.phase in ('test', 'train_and_test'):
            run_test(model, args)
    except Exception:
        traceback.print_exc()
        raise
    print("Job finished...")


if __name__ == '__main__':
    main(parse_args())

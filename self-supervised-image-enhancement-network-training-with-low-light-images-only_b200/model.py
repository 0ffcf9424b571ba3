"""Drop-in replacement for the reference's `model.py` hot path (the nn.Module surface of SURVEY.md §8b).

Same constructor signature, attribute names, state_dict keys, `forward` / `compute_loss` return contract,
`train_model` / `evaluate_model` / `test_model` / `save_checkpoint` / `load_checkpoint` entry points as
/root/reference/model.py:177-607.  All arithmetic of forward, loss, backward and Adam runs in the CUDA library
behind include/sshslie_b200.h; this file is host glue only and has no PyTorch compute fallback.

Differences a caller can observe: parameters live in ONE flat fp32 buffer (each nn.Parameter
is a view, so `state_dict()` is unchanged); `compute_loss` runs forward AND backward in the same call (the
returned loss still supports `loss.backward()`, which just hands the finished gradients to autograd);
`self.optimizer` is a fused single-launch Adam with torch.optim.Adam's state_dict layout.
"""
import ctypes
import os
import time
from glob import glob

import numpy as np
import torch
import torch.nn as nn

from . import lib as L

LOSS_KEYS = ["total_loss", "L_reconstruction", "L_R_fidelity", "L_I_smooth_low", "L_I_smooth_delta",
             "L_fourier", "L_spectral_cons"]


def conv(in_channels, out_channels, kernel_size, stride=1, padding=None, activation=True):
    """Parameter container with the reference's module structure (model.py:17-23) -> keys '<name>.0.weight'."""
    if padding is None:
        padding = (kernel_size - 1) // 2
    layers = [nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding)]
    if activation:
        layers.append(nn.ReLU(inplace=True))
    return nn.Sequential(*layers)


class DecompositionNet(nn.Module):
    """Parameters of model.py:25-47.  Calling it runs the CUDA forward of the owning LowLightEnhance."""

    def __init__(self, in_channels, channel=64, kernel_size=3):
        super().__init__()
        self.in_channels, self.channel, self.kernel_size = in_channels, channel, kernel_size
        self.conv0 = conv(in_channels, channel // 2, kernel_size, activation=True)
        self.shallow_conv = conv(in_channels, channel, kernel_size * 3, activation=False)
        self.conv1 = conv(channel, channel, kernel_size, activation=True)
        self.conv2 = conv(channel, channel * 2, kernel_size, stride=2, activation=True)
        self.conv3 = conv(channel * 2, channel * 2, kernel_size, activation=True)
        self.deconv = nn.Sequential(
            nn.ConvTranspose2d(channel * 2, channel, kernel_size, stride=2, padding=(kernel_size - 1) // 2,
                               output_padding=1),
            nn.ReLU(inplace=True))
        self.conv5 = conv(channel + channel, channel, kernel_size, activation=True)
        self.conv7 = conv(channel + channel // 2, channel, kernel_size, activation=False)
        self.recon = nn.Conv2d(channel, in_channels + 1, kernel_size, stride=1, padding=(kernel_size - 1) // 2)
        self._owner = None

    def forward(self, x):
        if self._owner is None:
            raise L.SshslieError("DecompositionNet must be owned by a LowLightEnhance to run")
        R, I, _, _ = self._owner[0].forward(x)
        return R, I


class TransformerBlock(nn.Module):
    """Parameters of model.py:87-97."""

    def __init__(self, channels, num_heads=4, head_dim=16, ff_dim=64):
        super().__init__()
        self.num_heads, self.head_dim, self.total_dim = num_heads, head_dim, num_heads * head_dim
        self.q_linear = nn.Linear(channels, self.total_dim)
        self.k_linear = nn.Linear(channels, self.total_dim)
        self.v_linear = nn.Linear(channels, self.total_dim)
        self.ff_linear1 = nn.Linear(self.total_dim, ff_dim)
        self.ff_linear2 = nn.Linear(ff_dim, channels)


class IllumAdjustmentNet(nn.Module):
    """Parameters of model.py:121-141 (use_transformer=True, the only configuration any caller builds)."""

    def __init__(self, in_channels, channel=64, kernel_size=3):
        super().__init__()
        self.conv0 = conv(in_channels + 1, channel, kernel_size, activation=False)
        self.conv1 = conv(channel, channel, kernel_size, stride=2, activation=True)
        self.conv2 = conv(channel, channel, kernel_size, stride=2, activation=True)
        self.conv3 = conv(channel, channel, kernel_size, stride=2, activation=True)
        self.attn = TransformerBlock(channel)
        self.deconv1 = conv(channel, channel, kernel_size, activation=True)
        self.deconv2 = conv(channel, channel, kernel_size, activation=True)
        self.deconv3 = conv(channel, channel, kernel_size, activation=True)
        self.feature_fusion = conv(channel * 3, channel, 1, activation=False)
        self.final_conv = nn.Conv2d(channel, 1, 3, stride=1, padding=1)
        self._owner = None

    def forward(self, I, R):
        """model.py:143-175: I (B,1,H,W), R (B,C,H,W) -> I_delta (B,1,H,W), through the owner's engine of that shape."""
        owner = getattr(self, "_owner", None)
        if owner is None:
            raise L.SshslieError("IllumAdjustmentNet must be owned by a LowLightEnhance to run")
        return owner[0].illum_forward(I, R)


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam (defaults) over the owner's flat buffers: one kernel launch per step (model.py:213,316).

    `state_dict()` has torch.optim.Adam's layout: per-parameter 'step', 'exp_avg', 'exp_avg_sq' (views of the
    flat moment buffers), so reference checkpoints load and ours load in the reference.
    """

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, owner=None):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False, foreach=None,
                        capturable=False, differentiable=False, fused=None)
        super().__init__(params, defaults)
        self._owner = [owner]
        self._step = 0
        self._state_ready = False
        self.grad_scale = 1.0

    def _ensure_state(self):
        """(Re)attach the per-parameter state entries to the flat moment buffers."""
        own = self._owner[0]
        own._ensure_flat()
        if own._flat_m is None or own._flat_m.device != own._flat.device:
            own._flat_m = torch.zeros_like(own._flat)
            own._flat_v = torch.zeros_like(own._flat)
            self._state_ready = False
        if self._state_ready and self._state_flat is own._flat_m:
            return
        for p, (off, n) in zip(own._plist, own._pranges):
            st = self.state[p]
            if "exp_avg" in st and st["exp_avg"].data_ptr() != own._flat_m.data_ptr() + 4 * off:
                own._flat_m[off:off + n].copy_(st["exp_avg"].reshape(-1))          # loaded from a checkpoint
                own._flat_v[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                self._step = max(self._step, int(st.get("step", 0)))
            st["exp_avg"] = own._flat_m[off:off + n].view(p.shape)
            st["exp_avg_sq"] = own._flat_v[off:off + n].view(p.shape)
            st["step"] = torch.tensor(float(self._step))
        self._state_ready = True
        self._state_flat = own._flat_m

    def zero_grad(self, set_to_none=True):
        own = self._owner[0]
        if set_to_none:
            for p in own._plist:
                p.grad = None
            own._grads_live = False
        else:
            super().zero_grad(set_to_none=False)
            # parameters that still hold the views of the flat gradient buffer now hold zeros: the next compute_loss may
            # overwrite the buffer and backward() leaves those views in place (0 + g = g)
            own._grads_live = False

    def _launch(self, lo, hi):
        own = self._owner[0]
        group = self.param_groups[0]
        b1, b2 = group["betas"]
        stream = ctypes.c_void_p(torch.cuda.current_stream(own._flat.device).cuda_stream)
        L.check(L.load().sshslie_adam_step(
            ctypes.c_void_p(own._flat.data_ptr() + 4 * lo), ctypes.c_void_p(own._flat_grad.data_ptr() + 4 * lo),
            ctypes.c_void_p(own._flat_m.data_ptr() + 4 * lo), ctypes.c_void_p(own._flat_v.data_ptr() + 4 * lo),
            hi - lo, float(group["lr"]), float(b1), float(b2), float(group["eps"]), self._step,
            float(self.grad_scale), stream), "sshslie_adam_step")

    @torch.no_grad()
    def step(self, closure=None):
        own = self._owner[0]
        self._ensure_state()
        self._step += 1
        views = own._grad_views
        if views is not None and all(p.grad is v for p, v in zip(own._plist, views)):
            self._launch(0, own._nparams)            # every gradient is the finished view of the flat buffer
            return None
        # general path: contiguous ranges of parameters that received a gradient (frozen decomposition_net -> tail)
        ranges = []
        for p, (off, n) in zip(own._plist, own._pranges):
            if p.grad is None:
                continue
            if p.grad.data_ptr() != own._flat_grad.data_ptr() + 4 * off:
                own._flat_grad[off:off + n].copy_(p.grad.reshape(-1))
            if ranges and ranges[-1][1] == off:
                ranges[-1][1] = off + n
            else:
                ranges.append([off, off + n])
        for lo, hi in ranges:
            self._launch(lo, hi)
        return None

    def state_dict(self):
        if self._owner[0]._plist[0].is_cuda:
            self._ensure_state()
            for st in self.state.values():
                st["step"] = torch.tensor(float(self._step))
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._step = 0
        for st in self.state.values():
            if "step" in st:
                self._step = max(self._step, int(st["step"]))
        self._state_ready = False
        self._ensure_state()


class _LossFn(torch.autograd.Function):
    """compute_loss already ran forward AND backward on the GPU.  The returned loss is wired to all 46 parameters, so
    the general autograd entry points work as on the reference's loss tensor - `torch.autograd.grad(loss, params)`,
    `(2 * loss).backward()`, accumulation through AccumulateGrad - with the finished gradient (times the incoming
    scalar, scaled on the device) as the result.  The plain `loss.backward()` of model.py:315 takes the shortcut in
    `_fast_backward` instead and never reaches this node."""

    @staticmethod
    def forward(ctx, owner_box, *params):
        ctx.owner_box = owner_box
        return owner_box[0]._losses_dev[0].clone()

    @staticmethod
    def backward(ctx, gout):
        own = ctx.owner_box[0]
        scaled = own._flat_grad * gout.reshape(())      # a fresh tensor: later steps do not overwrite what autograd holds
        grads = tuple(scaled[off:off + n].view(p.shape) if need else None
                      for p, (off, n), need in zip(own._plist, own._pranges, ctx.needs_input_grad[1:]))
        return (None,) + grads


def _hand_over(own, views):
    """p.grad <- finished gradient.  `views` alias the flat gradient buffer that every compute_loss overwrites, so:
      * no gradient yet                      -> p.grad = view (no copy; FusedAdam.step sees `p.grad is view`);
      * p.grad already IS that view           -> nothing to do: zero_grad(set_to_none=False) kept the view and the step
                                                 that just ran refilled it (adding the buffer to itself would give 2 g);
      * any other tensor (accumulated grads) -> a fresh tensor p.grad + view, as autograd's AccumulateGrad would."""
    for p, v in zip(own._plist, views):
        if not p.requires_grad:
            continue
        g = p.grad
        if g is None:
            p.grad = v
        elif g is v or (g.data_ptr() == v.data_ptr() and g.shape == v.shape):
            pass
        else:
            p.grad = g + v
    own._grads_live = True


def _fast_backward(own, total):
    """`loss.backward()` exactly as the reference calls it (model.py:315): the gradients are already finished, so the
    call only hands the views of the flat gradient buffer to the parameters - no autograd engine, no host sync."""
    def backward(gradient=None, retain_graph=None, create_graph=False, inputs=None):
        if gradient is not None or create_graph or inputs is not None:
            return torch.Tensor.backward(total, gradient, retain_graph, create_graph, inputs)
        _hand_over(own, own._grad_views)
    return backward


class LazyLosses(dict):
    """dict[str, float] whose values are read back from the device on first access.

    Given a CUDA tensor, construction enqueues an asynchronous copy of the seven floats into pinned host memory and records
    an event; the first access waits for THAT event only - not for work enqueued afterwards - so a loop that logs the losses
    of step i after launching step i+1 never stalls the device (the reference's `.item()` calls, model.py:566-574, stall
    it once per step)."""

    def __init__(self, dev_tensor, term_scale=1.0):
        super().__init__()
        self._dev = dev_tensor
        self._event = None
        if dev_tensor.is_cuda:
            host = torch.empty(dev_tensor.shape, dtype=dev_tensor.dtype, pin_memory=True)
            host.copy_(dev_tensor.detach(), non_blocking=True)
            self._event = torch.cuda.Event()
            self._event.record()
            self._dev = host
        self._ready = False
        self._term_scale = term_scale          # data parallel: the six term values arrive as sums over the ranks

    def _sync(self):
        if not self._ready:
            if self._event is not None:
                self._event.synchronize()
            vals = self._dev.detach().cpu().tolist()
            for i, (k, v) in enumerate(zip(LOSS_KEYS, vals)):
                dict.__setitem__(self, k, v if i == 0 else v * self._term_scale)
            self._ready = True

    def __getitem__(self, k):
        self._sync()
        return dict.__getitem__(self, k)

    def keys(self):
        self._sync()
        return dict.keys(self)

    def items(self):
        self._sync()
        return dict.items(self)

    def values(self):
        self._sync()
        return dict.values(self)

    def __iter__(self):
        self._sync()
        return dict.__iter__(self)

    def __len__(self):
        return len(LOSS_KEYS)

    def __repr__(self):
        self._sync()
        return dict.__repr__(self)

    def __contains__(self, k):
        return k in LOSS_KEYS

    def get(self, k, default=None):
        self._sync()
        return dict.get(self, k, default)

    def copy(self):
        self._sync()
        return dict(self)

    def __eq__(self, other):
        self._sync()
        return dict.__eq__(self, other)

    def __ne__(self, other):
        return not self.__eq__(other)

    def __bool__(self):
        return True

    def __reduce__(self):          # pickle / copy.deepcopy see the plain dict of floats the reference returns
        self._sync()
        return (dict, (dict(self),))


class DevicePatchSampler:
    """Resident training cubes + one-kernel crop/augment/transposition (sshslie_gather_patches).

    Draws (x, y, mode) per sample with numpy in the reference's order (model.py:306-308), so a seeded run picks the same
    patches as the reference; only the pixel work moves to the GPU."""

    def __init__(self, cubes_hwc, batch_size, patch_size, channels, device):
        self.cubes = [torch.from_numpy(np.ascontiguousarray(c, dtype=np.float32)).to(device) for c in cubes_hwc]
        self.shapes = [c.shape for c in cubes_hwc]
        self.B, self.ps, self.C, self.device = batch_size, patch_size, channels, device
        self.ptrs_host = torch.empty(batch_size, dtype=torch.int64).pin_memory()
        self.meta_host = torch.empty(batch_size, 5, dtype=torch.int32).pin_memory()
        self.ptrs = torch.empty(batch_size, dtype=torch.int64, device=device)
        self.meta = torch.empty(batch_size, 5, dtype=torch.int32, device=device)
        self.out = torch.empty(batch_size, channels, patch_size, patch_size, device=device)
        self._copied = None                    # event: the pinned staging buffers have been read by the last H2D

    def sample(self, batch_id):
        n = len(self.cubes)
        if self._copied is not None:
            self._copied.synchronize()         # do not overwrite pinned metadata an in-flight copy still reads
        for i in range(self.B):
            idx = (batch_id * self.B + i) % n
            h, w, _ = self.shapes[idx]
            x = np.random.randint(0, h - self.ps)
            y = np.random.randint(0, w - self.ps)
            mode = np.random.randint(0, 8)
            self.ptrs_host[i] = self.cubes[idx].data_ptr()
            self.meta_host[i] = torch.tensor([h, w, x, y, mode], dtype=torch.int32)
        return self.gather()

    def gather(self):
        self.ptrs.copy_(self.ptrs_host, non_blocking=True)
        self.meta.copy_(self.meta_host, non_blocking=True)
        if self._copied is None:
            self._copied = torch.cuda.Event()
        self._copied.record()
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        L.check(L.load().sshslie_gather_patches(L.ptr(self.ptrs), L.ptr(self.meta), L.ptr(self.out), self.B, self.C,
                                                self.ps, stream), "sshslie_gather_patches")
        return self.out


class _Engine:
    """One bound sshslie_engine + its workspace and static I/O buffers for a (B, H, W, train) shape."""

    def __init__(self, device, B, C, H, W, train, force_simt=False):
        lib = L.load()
        self.B, self.C, self.H, self.W, self.train = B, C, H, W, train
        self.handle = ctypes.c_void_p()
        flags = (L.FLAG_TRAIN if train else 0) | (L.FLAG_FORCE_SIMT if force_simt else 0)
        L.check(lib.sshslie_engine_create(ctypes.byref(self.handle), B, C, H, W, flags), "sshslie_engine_create")
        nbytes = lib.sshslie_engine_workspace_bytes(self.handle)
        self.workspace = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
        base = (self.workspace.data_ptr() + 1023) // 1024 * 1024
        stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        L.check(lib.sshslie_engine_bind(self.handle, ctypes.c_void_p(base), nbytes, stream), "sshslie_engine_bind")
        self.x = torch.empty(B, C, H, W, device=device)
        if train:          # static output buffers of the (graph-captured) training step; forward() returns fresh tensors
            self.R = torch.empty(B, C, H, W, device=device)
            self.I = torch.empty(B, 1, H, W, device=device)
            self.Id = torch.empty(B, 1, H, W, device=device)
            self.S = torch.empty(B, C, H, W, device=device)
        self.graph = None
        self.graph_cfg = None                  # loss weights baked into the captured kernel arguments
        self.calls = 0

    def __del__(self):
        try:
            if self.handle:
                L.load().sshslie_engine_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class LowLightEnhance(nn.Module):
    def __init__(self, input_channels=64, lr=1e-3, lr_update_factor=1, lr_update_period=None, time_stamp=None,
                 c_loss_reconstruction=10, c_loss_r_fidelity=1, c_loss_i_smooth_low=1, c_loss_i_smooth_delta=20,
                 c_loss_fourier=0.2, c_loss_spectral_cons=1, alpha_i_smooth_low=1, alpha_i_smooth_delta=10,
                 device=torch.device("cpu"), global_min=None, global_max=None,
                 save_reflectance=False, save_illumination=False, save_i_delta=False):
        super().__init__()
        if input_channels != 64:
            raise L.SshslieError("the CUDA path is built for 64 spectral bands (every reference config)")
        self.input_channels = input_channels
        self.device = device
        self.time_stamp = time_stamp
        self.c_loss_reconstruction = c_loss_reconstruction
        self.c_loss_r_fidelity = c_loss_r_fidelity
        self.c_loss_i_smooth_low = c_loss_i_smooth_low
        self.c_loss_i_smooth_delta = c_loss_i_smooth_delta
        self.c_loss_fourier = c_loss_fourier
        self.c_loss_spectral_cons = c_loss_spectral_cons
        self.alpha_i_smooth_low = alpha_i_smooth_low
        self.alpha_i_smooth_delta = alpha_i_smooth_delta
        self.lr = lr
        self.lr_update_factor = lr_update_factor
        self.lr_update_period = lr_update_period
        self.adaptive_lr = abs(self.lr_update_factor - 1) > 1e-6          # model.py:207-208
        self.global_min, self.global_max = global_min, global_max
        self.save_reflectance = save_reflectance
        self.save_illumination = save_illumination
        self.save_i_delta = save_i_delta
        self.eval_metrics = {}

        self.decomposition_net = DecompositionNet(in_channels=input_channels)
        self.illum_adjust_net = IllumAdjustmentNet(in_channels=input_channels)
        self.decomposition_net._owner = [self]
        self.illum_adjust_net._owner = [self]       # (a list: the owner must not become a sub-module of its own child)

        self._plist = list(self.parameters())
        total, offs, sizes = L.param_table(input_channels)
        assert len(self._plist) == L.NUM_PARAM_TENSORS and [p.numel() for p in self._plist] == sizes
        self._pranges = list(zip(offs, sizes))
        self._nparams = total
        self._flat = self._flat_grad = self._flat_m = self._flat_v = None
        self._grad_views = None
        self._engines = {}
        self._losses_dev = None
        self.use_cuda_graph = True
        self.max_cached_engines = 4
        self.force_simt = False
        self.dp_group = None                 # set by enable_data_parallel()
        self._box = [self]

        self.optimizer = FusedAdam(self.parameters(), lr=self.lr, owner=self)
        self.freeze_decom_epochs = 0
        if self.adaptive_lr:
            self.scheduler = torch.optim.lr_scheduler.StepLR(self.optimizer, step_size=self.lr_update_period,
                                                             gamma=self.lr_update_factor)
        self.all_epoch_losses = {k: [] for k in LOSS_KEYS}

    # ------------------------------------------------------------------ flat parameter storage
    def _ensure_flat(self):
        p0 = self._plist[0]
        ok = (self._flat is not None and self._flat.device == p0.device
              and p0.data_ptr() == self._flat.data_ptr()
              and self._plist[-1].data_ptr() == self._flat.data_ptr() + 4 * self._pranges[-1][0])
        if ok:
            return
        if not p0.is_cuda:
            raise L.SshslieError("LowLightEnhance parameters must be on a CUDA device (call .to('cuda')); "
                                 "there is no CPU implementation of the hot path")
        with torch.no_grad():
            flat = torch.empty(self._nparams, dtype=torch.float32, device=p0.device)
            for p, (off, n) in zip(self._plist, self._pranges):
                flat[off:off + n].copy_(p.detach().reshape(-1).float())
                p.data = flat[off:off + n].view(p.shape)
        self._flat = flat
        # gradients and the 8 loss scalars share one allocation: under data parallelism the losses ride in the same
        # all-reduce as the illum_adjust_net bucket (the slice right before them) and one kernel turns sums into means
        self._grad_store = torch.zeros(self._nparams + 8, dtype=torch.float32, device=p0.device)
        self._flat_grad = self._grad_store[:self._nparams]
        self._grad_views = [self._flat_grad[off:off + n].view(p.shape) for p, (off, n) in
                            zip(self._plist, self._pranges)]
        self._losses_dev = self._grad_store[self._nparams:]
        self._dp_ranges = None
        self._engines = {}

    def _cfg_tuple(self):
        return (float(self.c_loss_reconstruction), float(self.c_loss_r_fidelity),
                float(self.c_loss_i_smooth_low), float(self.c_loss_i_smooth_delta),
                float(self.c_loss_fourier), float(self.c_loss_spectral_cons),
                float(self.alpha_i_smooth_low), float(self.alpha_i_smooth_delta))

    def _cfg(self):
        return L.LossCfg(*self._cfg_tuple())

    def _graph_current(self, eng):
        """The captured graph(s) carry the loss weights as kernel arguments: a change of model.c_loss_* / alpha_* after
        capture (loss-weight schedules, sweeps on one model object) drops them, and the step is captured again."""
        if eng.graph is not None and eng.graph_cfg != self._cfg_tuple():
            eng.graph = None
            eng.calls = max(eng.calls, 3)      # already warm: capture again on this very call
        return eng.graph is not None

    def _keep_accumulated_grads(self):
        """compute_loss overwrites the flat gradient buffer.  Gradients a previous backward() handed out as views of it
        and that were not cleared since (micro-batch accumulation: backward; compute_loss(x2); backward) are detached
        into tensors of their own first, so that the next backward() adds g2 to g1 instead of to itself."""
        if not getattr(self, "_grads_live", False) or self._grad_views is None:
            return
        for p, v in zip(self._plist, self._grad_views):
            if p.grad is v:
                p.grad = v.clone()
        self._grads_live = False

    def _engine(self, x, train):
        B, C, H, W = x.shape
        key = (B, C, H, W, train, self.force_simt)
        eng = self._engines.pop(key, None)
        if eng is None:
            # a bound engine owns its workspace (~1 GB for a 512x512 cube): keep the few most recently used shapes only,
            # so that test_model over cubes of many different sizes does not accumulate workspaces
            while len(self._engines) >= self.max_cached_engines:
                self._engines.pop(next(iter(self._engines)))
            eng = _Engine(self._flat.device, B, C, H, W, train, self.force_simt)
        self._engines[key] = eng               # (re)insert as most recently used
        return eng

    def _stage_input(self, eng, x):
        if x.dtype != torch.float32:
            x = x.float()
        ev = getattr(x, "_sshslie_ready", None)
        if ev is not None:                     # a batch from prefetch(): its H2D copy ran on the copy stream
            torch.cuda.current_stream(self._flat.device).wait_event(ev)
        eng.x.copy_(x, non_blocking=True)      # H2D (or D2D) into the engine's static input buffer

    def prefetch(self, x_host):
        """Start the host->device copy of a (pinned) batch on a copy stream and return the device tensor; pass it to
        compute_loss / forward later.  Two staging buffers per shape alternate, so the copy of batch i+1 overlaps the
        step on batch i (the reference's loop pays torch.from_numpy(...).to(device) serially, model.py:312).  At most
        one prefetched batch per shape may be outstanding besides the one being consumed."""
        self._ensure_flat()
        dev = self._flat.device
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._stage_bufs = {}
        key = (tuple(x_host.shape), x_host.dtype)
        slot = self._stage_bufs.get(key)
        if slot is None:
            slot = self._stage_bufs[key] = {"bufs": [torch.empty(x_host.shape, dtype=x_host.dtype, device=dev)
                                                     for _ in range(2)], "used": [None, None], "i": 0}
        k = slot["i"]
        slot["i"] = k ^ 1
        buf = slot["bufs"][k]
        with torch.cuda.stream(self._copy_stream):
            if slot["used"][k] is not None:    # the step that last consumed this buffer has read it
                self._copy_stream.wait_event(slot["used"][k])
            buf.copy_(x_host, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        out = buf.view(buf.shape)              # a fresh tensor object carrying this copy's event
        out._sshslie_ready = ev
        out._sshslie_slot = (slot, k)
        return out

    def _release_staged(self, x):
        ref = getattr(x, "_sshslie_slot", None)
        if ref is not None:                    # the consumer's copy out of the staging buffer is enqueued: mark it reusable
            slot, k = ref
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self._flat.device))
            slot["used"][k] = ev

    # ------------------------------------------------------------------ hot path
    def forward(self, input_low):
        """model.py:229-234 -> (R_low, I_low, I_delta, S), fp32 (N,C,H,W)/(N,1,H,W) on the module's device."""
        self._ensure_flat()
        eng = self._engine(input_low, train=False)
        self._stage_input(eng, input_low)
        self._release_staged(input_low)
        lib = L.load()
        stream = ctypes.c_void_p(torch.cuda.current_stream(self._flat.device).cuda_stream)
        # fresh result tensors per call, like the reference: a caller may hold R across two forward() calls
        B, C, H, W = eng.x.shape
        dev = eng.x.device
        R, S = torch.empty(B, C, H, W, device=dev), torch.empty(B, C, H, W, device=dev)
        I, Id = torch.empty(B, 1, H, W, device=dev), torch.empty(B, 1, H, W, device=dev)
        L.check(lib.sshslie_forward(eng.handle, L.ptr(eng.x), L.ptr(self._flat), L.ptr(R), L.ptr(I),
                                    L.ptr(Id), L.ptr(S), stream), "sshslie_forward")
        return R, I, Id, S

    def illum_forward(self, I, R):
        """`self.illum_adjust_net(I, R)` (model.py:232): the illumination net alone on caller-provided I and R."""
        self._ensure_flat()
        if not (I.is_cuda and R.is_cuda):
            raise L.SshslieError("sshslie_b200 runs on CUDA tensors only (no CPU implementation of the hot path)")
        if I.dim() != 4 or R.dim() != 4 or I.shape[1] != 1 or I.shape[0] != R.shape[0] or I.shape[2:] != R.shape[2:]:
            raise L.SshslieError(f"illum_adjust_net: expected I (B,1,H,W) and R (B,C,H,W), got {tuple(I.shape)} {tuple(R.shape)}")
        eng = self._engine(R, train=False)
        lib = L.load()
        stream = ctypes.c_void_p(torch.cuda.current_stream(self._flat.device).cuda_stream)
        Ic, Rc = I.float().contiguous(), R.float().contiguous()
        Id = torch.empty_like(Ic)
        L.check(lib.sshslie_illum_forward(eng.handle, L.ptr(Ic), L.ptr(Rc), L.ptr(self._flat), L.ptr(Id), stream),
                "sshslie_illum_forward")
        return Id

    def _launch_loss_and_grad(self, eng, phase_mask=3):
        lib = L.load()
        stream = ctypes.c_void_p(torch.cuda.current_stream(self._flat.device).cuda_stream)
        cfg = self._cfg()
        L.check(lib.sshslie_loss_and_grad(eng.handle, L.ptr(eng.x), L.ptr(self._flat), ctypes.byref(cfg),
                                          L.ptr(self._flat_grad), L.ptr(self._losses_dev), L.ptr(eng.R),
                                          L.ptr(eng.I), L.ptr(eng.Id), L.ptr(eng.S), phase_mask, stream),
                "sshslie_loss_and_grad")

    def compute_loss(self, input_low):
        """model.py:544-575 -> (total_loss 0-dim tensor supporting .backward(), dict of 7 floats)."""
        self._ensure_flat()
        eng = self._engine(input_low, train=True)
        self._keep_accumulated_grads()
        self._stage_input(eng, input_low)
        self._release_staged(input_low)
        dp = self.dp_group is not None
        if dp:
            self._dp_step(eng)
        elif self.use_cuda_graph:
            eng.calls += 1
            if not self._graph_current(eng) and eng.calls >= 3:       # two eager warm-up steps, then capture
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._launch_loss_and_grad(eng)
                eng.graph = g
                eng.graph_cfg = self._cfg_tuple()
            if eng.graph is not None:
                eng.graph.replay()
            else:
                self._launch_loss_and_grad(eng)
        else:
            self._launch_loss_and_grad(eng)
        self.last_outputs = (eng.R, eng.I, eng.Id, eng.S)
        if torch.is_grad_enabled():
            total = _LossFn.apply(self._box, *self._plist)
            total.backward = _fast_backward(self, total)       # instance attribute shadows Tensor.backward
        else:
            total = self._losses_dev[0].clone()
        return total, LazyLosses(self._losses_dev[:7], 1.0 / self._dp_world if dp else 1.0)

    def profile_step(self, input_low, train=True):
        """One eager training step (train=False: forward only) with device timing per launch group: list of
        (name, ms, flops, bytes)."""
        self._ensure_flat()
        eng = self._engine(input_low, train=train)
        self._stage_input(eng, input_low)
        lib = L.load()
        stream = ctypes.c_void_p(torch.cuda.current_stream(self._flat.device).cuda_stream)
        cfg = self._cfg()
        n = lib.sshslie_profile_step(eng.handle, L.ptr(eng.x), L.ptr(self._flat), ctypes.byref(cfg),
                                     L.ptr(self._flat_grad), L.ptr(self._losses_dev), stream)
        if n < 0:
            L.check(n, "sshslie_profile_step")
        rows = []
        buf = ctypes.create_string_buffer(256)
        ms, fl, by = ctypes.c_float(), ctypes.c_double(), ctypes.c_double()
        for i in range(n):
            lib.sshslie_profile_row(i, buf, 256, ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(by))
            rows.append((buf.value.decode(), ms.value, fl.value, by.value))
        return rows

    # ------------------------------------------------------------------ data parallel
    def enable_data_parallel(self, group=None):
        """Average gradients over `group` (NCCL) inside compute_loss; the illum_adjust_net bucket is reduced
        while the first-pass decomposition backward still runs (phase split of sshslie_loss_and_grad)."""
        import torch.distributed as dist
        self.dp_group = group if group is not None else dist.group.WORLD
        self._dp_world = dist.get_world_size(self.dp_group)
        self._dp_stream = torch.cuda.Stream(device=self._plist[0].device)
        # identical weights on every rank
        self._ensure_flat()
        dist.broadcast(self._flat, src=dist.get_global_rank(self.dp_group, 0), group=self.dp_group)

    def _dp_cfg(self):
        """Loss weights of one rank's step under data parallelism: c_loss_* / world, so that the all-reduce(SUM) of the
        per-rank gradients IS the gradient of the global-batch loss (every term is a mean over equal shards) and no
        kernel has to rescale 1.14 M floats afterwards.  The alphas are not weights and stay as they are."""
        c = list(self._cfg_tuple())
        for k in range(6):
            c[k] = c[k] / self._dp_world
        return L.LossCfg(*c)

    def _dp_body(self, eng, cur):
        """One data-parallel step on stream `cur` (eager or under CUDA-graph capture): phase 1 -> all-reduce of the
        illum_adjust_net bucket + loss scalars on a side stream, overlapping phase 2 -> all-reduce of the decomposition
        bucket -> join."""
        from . import parallel as P
        dec, ill = self._dp_ranges
        lib = L.load()
        cfg = self._dp_cfg()

        def phase(mask):
            L.check(lib.sshslie_loss_and_grad(eng.handle, L.ptr(eng.x), L.ptr(self._flat), ctypes.byref(cfg),
                                              L.ptr(self._flat_grad), L.ptr(self._losses_dev), L.ptr(eng.R),
                                              L.ptr(eng.I), L.ptr(eng.Id), L.ptr(eng.S), mask,
                                              ctypes.c_void_p(cur.cuda_stream)), "sshslie_loss_and_grad")
        phase(1)                                     # fwd + loss + pass-2 bwd + illum bwd
        self._dp_stream.wait_stream(cur)
        with torch.cuda.stream(self._dp_stream):
            # illum_adjust_net bucket + the loss scalars behind it, one call; overlaps the pass-1 backward below
            P.allreduce_bucket(self._grad_store, (ill[0], self._nparams + 8), self.dp_group)
        phase(2)                                     # pass-1 decomposition backward
        P.allreduce_bucket(self._grad_store, dec, self.dp_group)
        cur.wait_stream(self._dp_stream)

    def _dp_step(self, eng):
        from . import parallel as P
        if self._dp_ranges is None:
            self._dp_ranges = P.bucket_ranges([o for o, _ in self._pranges], [n for _, n in self._pranges])
        cur = torch.cuda.current_stream(self._flat.device)
        eng.calls += 1
        one_graph = os.environ.get("SSHSLIE_DP_GRAPH", "1") != "0"
        if self.use_cuda_graph and one_graph:
            # the whole step - both phases AND the two NCCL all-reduces - is ONE graph: no host launch gaps between the
            # phase graphs and the collectives (those gaps were the scaling tail at 8 GPUs)
            if not self._graph_current(eng) and eng.calls >= 3:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._dp_body(eng, torch.cuda.current_stream(self._flat.device))
                eng.graph = g
                eng.graph_cfg = self._cfg_tuple()
            if eng.graph is not None:
                eng.graph.replay()
            else:
                self._dp_body(eng, cur)
            return
        self._dp_body(eng, cur)

    # ------------------------------------------------------------------ loops (host glue, model.py:236-443)
    def train_model(self, train_data_path, eval_data_path, batch_size, patch_size, num_epochs, start_lr, ckpt_dir,
                    eval_result_dir, eval_every_epoch, label_dir, plot_every_epoch=10):
        from .utils import load_hsi, data_augmentation
        ckpt_dir = os.path.join(ckpt_dir, 'Decomposition_' + str(self.time_stamp))
        os.makedirs(ckpt_dir, exist_ok=True)
        os.makedirs(eval_result_dir, exist_ok=True)
        train_files = sorted(glob(os.path.join(train_data_path, "*.mat")))
        train_low_data = [load_hsi(f, matContentHeader='data', normalization='global_normalization',
                                   max_val=self.global_max, min_val=self.global_min) for f in train_files]
        eval_files = sorted(glob(os.path.join(eval_data_path, "*.mat")))
        eval_low_data = [load_hsi(f, matContentHeader='data', normalization='global_normalization',
                                  max_val=self.global_max, min_val=self.global_min) for f in eval_files]
        num_batches = len(train_low_data) // batch_size
        dev = self._plist[0].device
        # crop + augmentation + HWC->NCHW run on the device from resident cubes (x, y, mode still drawn by numpy in the
        # reference's order, model.py:306-308); the reference's per-batch numpy work and 4 MiB/patch H2D are gone
        sampler = DevicePatchSampler(train_low_data, batch_size, patch_size, self.input_channels, dev)
        for epoch in range(num_epochs):
            if getattr(self, 'freeze_decom_epochs', 0) > 0:
                if epoch < self.freeze_decom_epochs:
                    for p in self.decomposition_net.parameters():
                        p.requires_grad = False
                    print(f"Epoch {epoch+1}: DecompositionNet frozen")
                elif epoch == self.freeze_decom_epochs:
                    for p in self.decomposition_net.parameters():
                        p.requires_grad = True
                    self.optimizer = FusedAdam(self.parameters(), lr=self.optimizer.param_groups[0]['lr'], owner=self)
                    self._flat_m = None
                    if self.adaptive_lr:
                        self.scheduler = torch.optim.lr_scheduler.StepLR(
                            self.optimizer, step_size=self.lr_update_period, gamma=self.lr_update_factor)
                    print(f"Epoch {epoch+1}: DecompositionNet unfrozen")
            cur = {k: 0 for k in LOSS_KEYS}
            count = 0
            for batch_id in range(num_batches):
                batch = sampler.sample(batch_id)
                self.optimizer.zero_grad()
                loss, batch_losses = self.compute_loss(batch)
                loss.backward()
                self.optimizer.step()
                for k in LOSS_KEYS:
                    cur[k] += batch_losses[k]
                count += 1
                print(f"Epoch [{epoch+1}/{num_epochs}] Batch [{batch_id+1}/{num_batches}] "
                      f"Loss: {batch_losses['total_loss']:.6f}")
            for k in LOSS_KEYS:
                self.all_epoch_losses[k].append(cur[k] / count if count > 0 else 0)
            avg = cur['total_loss'] / count if count > 0 else 0
            if (epoch + 1) % eval_every_epoch == 0:
                self.evaluate_model(eval_low_data, eval_files, eval_result_dir, epoch + 1, label_dir)
                self.save_checkpoint(os.path.join(ckpt_dir, f"model_epoch_{epoch+1}.pth"), epoch + 1)
                self.save_checkpoint(os.path.join(ckpt_dir, "model_epoch_latest.pth"), epoch + 1)
            if self.adaptive_lr:
                self.scheduler.step()
            print(f"Epoch [{epoch+1}/{num_epochs}] Average Loss: {avg:.6f}")

    def _to_hwc_host(self, t, denorm=False):
        """(1,C,H,W) device tensor -> (H,W,C) float32 numpy array: transposition (and, for S, the de-normalisation of
        model.py:423-424) run in one kernel on the device; the host only receives the finished cube."""
        _, C, H, W = t.shape
        out = torch.empty(H, W, C, dtype=torch.float32, device=t.device)
        scale, offset, apply = 1.0, 0.0, 0
        if denorm and self.global_min is not None and self.global_max is not None:
            scale = float(np.float32(self.global_max - self.global_min))
            offset, apply = float(np.float32(self.global_min)), 1
        stream = ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)
        L.check(L.load().sshslie_denorm_hwc(L.ptr(t.contiguous()), L.ptr(out), C, H, W, scale, offset, apply, stream),
                "sshslie_denorm_hwc")
        return out.cpu().numpy()

    def _save_outputs(self, out_dir, filename, R, I, Id, S, save_r, save_i, save_d):
        from .utils import save_hsi
        save_hsi(os.path.join(out_dir, filename), self._to_hwc_host(S, denorm=True))
        art = os.path.join(out_dir, 'artifacts')
        os.makedirs(art, exist_ok=True)
        stem = filename.split('.')[0]
        if save_r:
            save_hsi(os.path.join(art, stem + '_R_low.mat'), self._to_hwc_host(R))
        if save_i:
            save_hsi(os.path.join(art, stem + '_I_low.mat'), self._to_hwc_host(I))
        if save_d:
            save_hsi(os.path.join(art, stem + '_I_delta.mat'), self._to_hwc_host(Id))

    def evaluate_model(self, eval_low_data, eval_files, eval_result_dir, epoch, label_dir):
        if len(eval_low_data) <= 0:
            print(f"--- No files found for evaluation. Skipping evaluation for epoch {epoch} ---")
            return
        from . import metrics as M
        from .utils import load_hsi
        epoch_dir = os.path.join(eval_result_dir, f'epoch_{epoch}')
        os.makedirs(epoch_dir, exist_ok=True)
        acc, n = [0.0, 0.0, 0.0], 0
        with torch.no_grad():
            for idx, low_im in enumerate(eval_low_data):
                filename = os.path.basename(eval_files[idx])
                x = torch.from_numpy(low_im).unsqueeze(0).permute(0, 3, 1, 2)
                R, I, Id, S = self.forward(x)
                self._save_outputs(epoch_dir, filename, R, I, Id, S,
                                   self.save_reflectance, self.save_illumination, self.save_i_delta)
                # PSNR / SSIM / SAM against the label cube of the same name (model.py:390-397 -> metrics.calc_metrics with
                # data_max = global_max), on the device.  The reference reads its own output back with key 'ref' although
                # save_hsi wrote 'data' and so cannot get this far; here the metrics are computed whenever
                # the label file exists.
                label_path = os.path.join(label_dir, filename) if label_dir else None
                if label_path and os.path.exists(label_path) and self.global_max is not None:
                    label = torch.from_numpy(load_hsi(label_path, matContentHeader='data'))
                    pred = torch.from_numpy(self._to_hwc_host(S, denorm=True))
                    ps, sa = M.psnr_sam(pred, label, self.global_max)
                    acc[0] += ps
                    acc[1] += M.ssim(pred, label, self.global_max)
                    acc[2] += sa
                    n += 1
        if n > 0:
            self.eval_metrics[epoch] = {"psnr": acc[0] / n, "ssim": acc[1] / n, "sam": acc[2] / n}
            print(f"--- Evaluation for epoch {epoch}: PSNR {acc[0] / n:.4f} SSIM {acc[1] / n:.4f} SAM {acc[2] / n:.4f} ---")

    def test_model(self, model_dir, test_low_data, test_low_data_names, save_dir, save_reflectance=False,
                   save_illumination=False, save_i_delta=False):
        self.load_checkpoint(os.path.join(model_dir, 'model_epoch_latest.pth'))
        total = 0.0
        with torch.no_grad():
            for idx in range(len(test_low_data)):
                filename = os.path.basename(test_low_data_names[idx])
                x = torch.from_numpy(test_low_data[idx]).unsqueeze(0).permute(0, 3, 1, 2)
                torch.cuda.synchronize()
                t0 = time.time()
                R, I, Id, S = self.forward(x)
                torch.cuda.synchronize()                      # the reference's timer lacks this
                rt = time.time() - t0
                total += rt
                self._save_outputs(save_dir, filename, R, I, Id, S, save_reflectance, save_illumination,
                                   save_i_delta)
                print(f"Processed {filename} in {rt:.4f} seconds.")
        n = len(test_low_data)
        print(f"Average run time: {total / n if n else 0:.4f} seconds.")

    def save_checkpoint(self, path, epoch):
        self.optimizer._ensure_state()
        torch.save({'epoch': epoch, 'model_state_dict': self.state_dict(),
                    'optimizer_state_dict': self.optimizer.state_dict()}, path)
        print(f"Checkpoint saved at {path}")

    def load_checkpoint(self, path):
        ckpt = torch.load(path, map_location=self._plist[0].device)
        self.load_state_dict(ckpt['model_state_dict'])
        self.optimizer.load_state_dict(ckpt['optimizer_state_dict'])
        print(f"Loaded checkpoint from {path}")

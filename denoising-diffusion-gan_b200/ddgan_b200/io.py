"""Checkpoint and image-output formats of the reference loop.

* `content.pth` (ddgan.py:545-560): {'epoch', 'global_step', 'args', 'netG_dict', 'optimizerG', 'netD_dict', 'optimizerD',
  ['schedulerG', 'schedulerD',] 'emaG'} -- written / read here for a `train.Trainer`, with the optimiser entries in
  torch.optim.Adam's own state_dict layout and the EMA as {parameter name: CPU tensor} (ema.py:81-95), so files travel both ways
  between this framework and the reference scripts.
* `netG_<epoch>.pth` (ddgan.py:561-566): the generator's state_dict with the EMA weights swapped in.
* sampled images (test_ddgan.py:190-201): the reference maps to [0, 1], then saves one PNG (and optionally one .npy) per image
  with a device->host transfer and a float->byte conversion each.  Here the whole batch is quantised on the GPU by one kernel
  (ddg_images_to_u8, same rounding as torchvision.utils.save_image), leaves the device as bytes in one pinned copy and is encoded
  by a small thread pool.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import _lib


# ------------------------------------------------------------------------------------------------------------------
# images
# ------------------------------------------------------------------------------------------------------------------
def images_to_uint8(x: torch.Tensor, to_range_0_1: bool = True, out: torch.Tensor = None) -> torch.Tensor:
    """x [N, C, H, W] fp32 on the GPU -> uint8 [N, H, W, C]: clamp(((x + 1) / 2) * 255 + 0.5) truncated, i.e. to_range_0_1
    (test_ddgan.py:149) followed by save_image's `mul(255).add_(0.5).clamp_(0, 255).to(uint8)`."""
    _lib.require_cuda_f32(x)
    x = x.contiguous()
    n, c, h, w = x.shape
    if out is None:
        out = torch.empty(n, h, w, c, dtype=torch.uint8, device=x.device)
    s, b = (0.5, 0.5) if to_range_0_1 else (1.0, 0.0)
    _lib.check(_lib.lib().ddg_images_to_u8(_lib.ptr(x), _lib.ptr(out), n, c, h, w, s, b, _lib.stream()), 'images_to_u8')
    return out


class ImageWriter:
    """Batched replacement of the per-image save loop of test_ddgan.py:190-201.  write(batch, start_index) quantises on the GPU,
    copies the bytes to a pinned buffer and hands the encoding ('png' through PIL, or 'npy') to worker threads."""

    def __init__(self, save_dir, fmt='png', workers=8, save_npy=False):
        self.dir = save_dir
        os.makedirs(save_dir, exist_ok=True)
        self.fmt = fmt
        self.save_npy = save_npy
        self.pool = ThreadPoolExecutor(max_workers=workers)
        self.pending = []
        self._pinned = None

    def _encode(self, arr, index):
        path = os.path.join(self.dir, f'{index}.{self.fmt}')
        if self.fmt == 'npy':
            np.save(path, arr)
            return path
        from PIL import Image
        a = arr[..., 0] if arr.shape[-1] == 1 else arr
        Image.fromarray(a).save(path)
        return path

    def write(self, batch: torch.Tensor, start_index: int = 0):
        u8 = images_to_uint8(batch)
        if self._pinned is None or self._pinned.shape != u8.shape:
            self.flush()
            self._pinned = torch.empty(u8.shape, dtype=torch.uint8).pin_memory()
        else:
            self.flush()                                   # the previous batch's encoders still read the pinned buffer
        self._pinned.copy_(u8, non_blocking=True)
        if self.save_npy:                                  # test_ddgan.py:185-188 saves the [0, 1] float image as well
            f = ((batch.detach() + 1.0) / 2.0).cpu().numpy()
            for j in range(f.shape[0]):
                np.save(os.path.join(self.dir, f'{start_index + j}.npy'), f[j])
        torch.cuda.current_stream().synchronize()
        host = self._pinned.numpy()
        for j in range(host.shape[0]):
            self.pending.append(self.pool.submit(self._encode, host[j], start_index + j))

    def flush(self):
        for f in self.pending:
            f.result()
        self.pending = []

    def close(self):
        self.flush()
        self.pool.shutdown()


# ------------------------------------------------------------------------------------------------------------------
# checkpoints
# ------------------------------------------------------------------------------------------------------------------
def _plain_state_dict(net):
    return {k: v.detach().clone() for k, v in net.state_dict().items()}


def save_checkpoint(path, trainer, epoch, global_step, args=None, schedulerG=None, schedulerD=None):
    """ddgan.py:545-560 (`content.pth`)."""
    a = args if args is not None else trainer.args
    content = {'epoch': epoch, 'global_step': global_step, 'args': dict(vars(a)) if not isinstance(a, dict) else dict(a),
               'netG_dict': _plain_state_dict(trainer.netG), 'optimizerG': trainer.optG.state_dict(),
               'netD_dict': _plain_state_dict(trainer.netD), 'optimizerD': trainer.optD.state_dict()}
    if schedulerG is not None:
        content['schedulerG'] = schedulerG.state_dict()
    if schedulerD is not None:
        content['schedulerD'] = schedulerD.state_dict()
    ema = trainer.ema_state_dict()
    if ema:
        content['emaG'] = ema
    torch.save(content, path)
    return content


def load_checkpoint(path, trainer, map_location=None, schedulerG=None, schedulerD=None):
    """ddgan.py:373-410 (resume): accepts files written by the reference (DistributedDataParallel's 'module.' prefix included)."""
    ck = torch.load(path, map_location=map_location, weights_only=False)
    strip = lambda sd: {(k[len('module.'):] if k.startswith('module.') else k): v for k, v in sd.items()}
    trainer.netG.load_state_dict(strip(ck['netG_dict']))
    trainer.netD.load_state_dict(strip(ck['netD_dict']))
    trainer.optG.load_state_dict(ck['optimizerG'])
    trainer.optD.load_state_dict(ck['optimizerD'])
    if schedulerG is not None and 'schedulerG' in ck:
        schedulerG.load_state_dict(ck['schedulerG'])
    if schedulerD is not None and 'schedulerD' in ck:
        schedulerD.load_state_dict(ck['schedulerD'])
    if 'emaG' in ck:
        if trainer.fused_optim:
            trainer.optG.load_ema_state_dict(ck['emaG'])
        elif trainer.ema is not None:
            trainer.ema.load_state_dict(ck['emaG'])
    return ck['epoch'], ck['global_step'], ck


def save_generator(path, trainer):
    """ddgan.py:561-566: netG_<epoch>.pth holds the EMA weights (swap in, save, swap back)."""
    trainer.swap_parameters_with_ema(store_params_in_ema=True)
    try:
        torch.save(_plain_state_dict(trainer.netG), path)
    finally:
        trainer.swap_parameters_with_ema(store_params_in_ema=True)

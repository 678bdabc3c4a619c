"""FIR resampling layers on the public NCHW surface (reference: score_sde/models/up_or_down_sampling.py:149-262).
Inside the fused engines these are folded into PNHWC kernels; the functions here serve external callers and tests."""
import numpy as np
import torch

from score_sde.op import upfirdn2d


def _setup_kernel(k):
    k = np.asarray(k, dtype=np.float32)
    if k.ndim == 1:
        k = np.outer(k, k)
    k /= np.sum(k)
    assert k.ndim == 2 and k.shape[0] == k.shape[1]
    return k


def upsample_2d(x, k=None, factor=2, gain=1):
    assert isinstance(factor, int) and factor >= 1
    k = _setup_kernel([1] * factor if k is None else k) * (gain * (factor ** 2))
    p = k.shape[0] - factor
    return upfirdn2d(x, torch.tensor(k, device=x.device), up=factor, pad=((p + 1) // 2 + factor - 1, p // 2))


def downsample_2d(x, k=None, factor=2, gain=1):
    assert isinstance(factor, int) and factor >= 1
    k = _setup_kernel([1] * factor if k is None else k) * gain
    p = k.shape[0] - factor
    return upfirdn2d(x, torch.tensor(k, device=x.device), down=factor, pad=((p + 1) // 2, p // 2))


def conv_downsample_2d(x, w, k=None, factor=2, gain=1):
    """FIR (pad = (p+1)//2, p//2 with p = (len(k) - factor) + (convW - 1)) followed by the stride-`factor` convolution, both on
    this library's kernels: the FIR output is written space-to-depth and the strided 3x3 conv runs as a stride-1 2x2-tap
    tensor-core conv (ddgan_b200.train_graph.conv_downsample_pnhwc).  Supported: the configuration the reference uses,
    k = [1, 3, 3, 1], factor 2, 3x3 weights, even H and W (up_or_down_sampling.py:75-79)."""
    assert isinstance(factor, int) and factor >= 1
    _outC, _inC, convH, convW = w.shape
    assert convW == convH
    kk = _setup_kernel([1] * factor if k is None else k) * gain
    ref = _setup_kernel([1, 3, 3, 1])
    if factor != 2 or convW != 3 or kk.shape != ref.shape or not np.allclose(kk, ref) or x.shape[2] % 2 or x.shape[3] % 2:
        raise NotImplementedError('conv_downsample_2d: only k=[1,3,3,1], factor=2, 3x3 weights, even sizes (the reference configuration)')
    from ddgan_b200 import train_graph as TG, ops
    n, c, h, wd = x.shape
    xp = TG.ToPnhwcFn.apply(x, ops.pad_c(c))
    y = TG.conv_downsample_pnhwc(xp, w, None, n, h, wd)
    return TG.FromPnhwcFn.apply(y, _outC)


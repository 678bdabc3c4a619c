"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/*.h declares
(no compute calls without a GPU), argument validation fails loudly, and the ctypes mirror matches the header."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'ddgan_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(ddg_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from ddgan_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = _lib.lib()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/ddgan_b200.h but not exported'
    assert set(_lib.exported_symbols()) == set(names), set(_lib.exported_symbols()) ^ set(names)
    assert lib.ddg_version() == 100


def test_host_side_helpers_without_gpu():
    from ddgan_b200 import _lib
    lib = _lib.lib()
    # out size arithmetic of upfirdn2d.py:111-112
    assert lib.ddg_upfirdn2d_out_size(32, 1, 2, 1, 1, 4) == 16
    assert lib.ddg_upfirdn2d_out_size(16, 2, 1, 2, 1, 4) == 32
    assert lib.ddg_upfirdn2d_out_size(32, 1, 1, 2, 2, 4) == 33
    # tiny levels take narrow tiles (more CTAs, shorter MMA chains); one wave of 128-row tiles switches to N = 256
    assert lib.ddg_conv_tile_n(256, 1000) == 64 and lib.ddg_conv_tile_n(256, 6400) == 128 and lib.ddg_conv_tile_n(256, 64 * 34 * 34) == 256
    assert lib.ddg_conv_tile_n(3, 10 ** 6) == 16 and lib.ddg_conv_tile_n(64, 10 ** 6) == 64 and lib.ddg_conv_tile_n(128, 10 ** 6) == 128
    # 3x3 conv 256->256: 8 k-blocks x 9 taps, 2 n-tiles, hi+lo planes of 32x128 bf16
    assert lib.ddg_conv_packed_bytes(256, 72, 32, 3, 128) == 2 * 72 * 32 * 128 * 2 * 2
    assert lib.ddg_conv_packed_bytes(256, 72, 32, 3, 256) == 1 * 72 * 32 * 256 * 2 * 2


def test_bad_arguments_fail_loudly():
    from ddgan_b200 import _lib
    lib = _lib.lib()
    rc = lib.ddg_upfirdn2d(None, None, None, 1, 4, 4, 4, 4, 1, 1, 1, 1, 0, 0, 0, 0, None)
    assert rc < 0 and b'upfirdn2d' in lib.ddg_last_error()
    rc = lib.ddg_conv2d_fwd(ctypes.byref(_lib.ConvDesc()), None)
    assert rc < 0


def test_ops_refuse_cpu_tensors():
    import torch
    from ddgan_b200 import ops
    with pytest.raises(RuntimeError):
        ops.upfirdn2d_raw(torch.zeros(1, 4, 4), torch.ones(2, 2), 1, 1, 1, 1, 0, 0, 0, 0)


def test_frame_pool_hands_out_fresh_tensor_objects():
    """ops.FramePool (training graphs inside Trainer.step): one cursor per shape rewound every step, distinct buffers within a step,
    the same storage for the same call of the next step -- and a NEW tensor object every time.  Reusing the object itself would keep
    last step's autograd history and hook table: a hook registered on it again never reaches the new grad_fn (that is how the
    data-parallel early all-reduce stopped firing from the second step on; found by tests/test_multigpu_gpu.py)."""
    import torch
    from ddgan_b200 import ops
    pool = ops.FramePool()
    made = []

    def make():
        made.append(torch.zeros(2, 3))
        return made[-1]
    fired = []
    ptrs = []
    for step in range(3):
        pool.begin_step()
        a = pool.get('k', make)
        b = pool.get('k', make)
        pool.end_step()
        assert a.data_ptr() != b.data_ptr()                       # distinct within a step
        ptrs.append((a.data_ptr(), b.data_ptr()))
        assert a.grad_fn is None and not a.requires_grad            # no history carried over
    # the way train_graph uses the pool: a custom Function fetches its output buffer inside forward, the caller hooks the output.
    # (With the stored object itself returned, only the first step's hook fires.)
    class Fill(torch.autograd.Function):
        @staticmethod
        def forward(ctx, w):
            out = pool.get('f', make)
            out.copy_(w)
            return out

        @staticmethod
        def backward(ctx, g):
            return g
    for step in range(3):
        pool.begin_step()
        w = torch.ones(2, 3, requires_grad=True)
        h = Fill.apply(w)
        h.register_hook(lambda g, s=step: fired.append(s))
        (h * 2).sum().backward()
        pool.end_step()
        assert torch.equal(w.grad, torch.full((2, 3), 2.0))
    assert len(made) == 3 and ptrs[0] == ptrs[1] == ptrs[2]         # same storage step after step (2 of key 'k', 1 of key 'f')
    assert fired == [0, 1, 2]
    assert not pool.active

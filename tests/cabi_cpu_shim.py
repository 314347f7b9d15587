"""TEST-ONLY stand-in for a handful of C-ABI entry points, computing on HOST memory with torch.

Purpose: the Python host side above the C ABI (which kernel is called with which pointer, leading dimension, epilogue
and accumulation flag; which gradient lands on which Parameter) can be exercised in the CPU test tier, where no GPU is
visible.  It restates what include/b200rec.h DOCUMENTS for each entry point, nothing more; the kernels themselves are
only ever validated on the GPU (tests/test_gpu_*.py).  The product never imports this module and has no CPU path:
`b200rec.hstu.HSTU.forward/predict` still raise on non-CUDA tensors; the tests below call internal host methods directly.
"""
import contextlib
import ctypes

import numpy as np
import torch

from b200rec import _lib as L


def _f32(ptr, rows, cols, ld=None):
    """A torch view (sharing memory) of the fp32 matrix [rows, cols] with leading dimension ld at address ptr."""
    ld = cols if ld is None else ld
    n = (rows - 1) * ld + cols if rows > 0 else 0
    if n == 0:
        return torch.empty((rows, cols))
    buf = (ctypes.c_float * n).from_address(ptr)
    arr = np.frombuffer(buf, dtype=np.float32, count=n)
    return torch.from_numpy(np.lib.stride_tricks.as_strided(arr, (rows, cols), (ld * 4, 4)))


def _silu_grad(z):
    s = torch.sigmoid(z)
    return s * (1 + z * (1 - s))


def _call(name, *a):
    if name == "b200rec_layernorm_fwd":          # (x, T, D, eps, y, y_dtype, mean, rstd, stream)
        x, T, D, eps, y, y_dt, mean, rstd, _ = a
        assert y_dt == L.F32
        xv = _f32(x, T, D)
        mu = xv.mean(1)
        var = xv.var(1, unbiased=False)
        rs = (var + eps).rsqrt()
        _f32(y, T, D).copy_((xv - mu[:, None]) * rs[:, None])
        _f32(mean, 1, T).copy_(mu[None])
        _f32(rstd, 1, T).copy_(rs[None])
    elif name == "b200rec_layernorm_bwd":        # (dy, dy_dtype, ldy, x, mean, rstd, T, D, residual_grad, dx, dx_act, stream)
        dy, dy_dt, ldy, x, mean, rstd, T, D, res, dx, dx_act, _ = a
        assert dy_dt == L.F32 and dx_act is None
        g, xv = _f32(dy, T, D, ldy), _f32(x, T, D)
        mu, rs = _f32(mean, 1, T)[0], _f32(rstd, 1, T)[0]
        xh = (xv - mu[:, None]) * rs[:, None]
        out = (g - g.mean(1, keepdim=True) - xh * (g * xh).mean(1, keepdim=True)) * rs[:, None]
        if res is not None:
            out = out + _f32(res, T, D)
        _f32(dx, T, D).copy_(out)
    elif name == "b200rec_resblock_bwd":         # (d_hd, z, act_dtype, T, H, D, dz, dy, stream)
        d_hd, z, a_dt, T, H, D, dz, dy, _ = a
        assert a_dt == L.F32
        g = _f32(d_hd, T, H * D)
        if z is not None:
            _f32(dz, T, H * D).copy_(g * _silu_grad(_f32(z, T, H * D)))
        _f32(dy, T, D).copy_(g.view(T, H, D).sum(1))
    elif name in EXTRA:                          # entry points served by a host build of the kernel's own source
        EXTRA[name](*a)
    else:
        raise AssertionError(f"cabi_cpu_shim: {name} is not restated")


EXTRA = {}


def _mat(t, rows, cols, ld, major):
    """Logical [rows, cols] operand: major 0 = stored [rows, cols] with ld; major 1 = stored [cols, rows] with ld."""
    assert t.dtype == torch.float32
    return _f32(t.data_ptr(), rows, cols, ld) if major == 0 else _f32(t.data_ptr(), cols, rows, ld).t()


def _gemm(A, B, C_out, M, N, K, *, lda, ldb, ldc, a_major=0, b_major=0, epilogue=L.EPI_STORE, alpha=1.0, alpha_dev=None,
          bias=None, resid=None, ldr=0, C2=None, ldc2=0, n_split=0, **kw):
    assert alpha_dev is None and not kw
    acc = alpha * (_mat(A, M, K, lda, a_major) @ _mat(B, N, K, ldb, b_major).t())
    C = _f32(C_out.data_ptr(), M, N, ldc)
    if epilogue == L.EPI_STORE:
        C.copy_(acc)
    elif epilogue == L.EPI_ACCUM:
        C.add_(acc)
    elif epilogue == L.EPI_BIAS_RESID:           # C = acc + bias[n] + resid[m, n]
        v = acc
        if bias is not None:
            v = v + _f32(bias.data_ptr(), 1, N)
        if resid is not None:
            v = v + _f32(resid.data_ptr(), M, N, ldr)
        C.copy_(v)
    elif epilogue == L.EPI_RESBLOCK:             # z = acc + bias[n]; C2 = z; C = resid[m, n % n_split] + silu(z)
        z = acc + _f32(bias.data_ptr(), 1, N)
        _f32(C2.data_ptr(), M, N, ldc2).copy_(z)
        ns = n_split if n_split > 0 else N
        r = _f32(resid.data_ptr(), M, ns, ldr)
        C.copy_(r.repeat(1, N // ns) + torch.nn.functional.silu(z))
    else:
        raise AssertionError(f"cabi_cpu_shim: epilogue {epilogue} is not restated")


def _colsum(x, rows, cols, ldx, out, accumulate=False):
    assert x.dtype == torch.float32
    s = _f32(x.data_ptr(), rows, cols, ldx).sum(0)
    o = _f32(out.data_ptr(), 1, cols)
    o.copy_(o[0] + s if accumulate else s)


@contextlib.contextmanager
def installed():
    """Routes b200rec._lib.call / gemm / colsum / stream to the host restatements inside the block."""
    saved = (L.call, L.gemm, L.colsum, L.stream)
    L.call, L.gemm, L.colsum, L.stream = _call, _gemm, _colsum, (lambda: None)
    try:
        yield
    finally:
        L.call, L.gemm, L.colsum, L.stream = saved

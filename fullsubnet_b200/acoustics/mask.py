"""Drop-in for audio_zen/acoustics/mask.py (build_complex_ideal_ratio_mask :7-29,
compress_cIRM :32-44, decompress_cIRM :47-64, complex_mul :67-70)."""
from __future__ import annotations

import torch

from .. import _lib
from ..constant import EPSILON  # noqa: F401  (mask.py:4; the kernel hard-codes 2**-23)


def _ew(fn_name, x, *scalars):
    x = _lib.require_cuda(x, fn_name)
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(getattr(_lib.load(), fn_name)(x.data_ptr(), out.data_ptr(), x.numel(), *scalars,
                                                 _lib.stream_ptr(x.device)))
    return out


def build_complex_ideal_ratio_mask(noisy_real, noisy_imag, clean_real, clean_imag) -> torch.Tensor:
    """[B,F,T] x4 -> compressed cIRM [B,F,T,2]  (mask.py:7-29)."""
    nr, ni, cr, ci = (_lib.require_cuda(t, "cIRM input") for t in (noisy_real, noisy_imag, clean_real, clean_imag))
    assert nr.shape == ni.shape == cr.shape == ci.shape
    out = torch.empty(*nr.shape, 2, dtype=torch.float32, device=nr.device)
    with torch.cuda.device(nr.device):
        _lib.check(_lib.load().fsn_build_cirm(nr.data_ptr(), ni.data_ptr(), cr.data_ptr(), ci.data_ptr(),
                                              out.data_ptr(), nr.numel(), _lib.stream_ptr(nr.device)))
    return out


def compress_cIRM(mask, K=10, C=0.1):
    """mask.py:32-44 (tensor branch)."""
    return _ew("fsn_compress_cirm", mask, float(K), float(C))


def decompress_cIRM(mask, K=10, limit=9.9):
    """mask.py:47-64"""
    return _ew("fsn_decompress_cirm", mask, float(K), float(limit))


def complex_mul(noisy_r, noisy_i, mask_r, mask_i):
    """mask.py:67-70"""
    r = noisy_r * mask_r - noisy_i * mask_i
    i = noisy_r * mask_i + noisy_i * mask_r
    return r, i

"""Drop-in for audio_zen/acoustics/feature.py (hot-path functions only: stft :9-50,
istft :53-91, mag_phase :94-96, drop_band :309-345).  Device work: libfsn_b200 (fsn_stft,
fsn_istft, fsn_drop_band)."""
from __future__ import annotations

import torch

from .. import _lib


def stft(y, n_fft, hop_length, win_length):
    """[B,T] or [B,C,T] wave -> (mag, phase, real, imag), each [B,F,T] / [B,C,F,T]  (feature.py:9-50)."""
    num_dims = y.dim()
    assert num_dims == 2 or num_dims == 3, "Only support 2D or 3D Input"
    batch_size = y.shape[0]
    num_samples = y.shape[-1]
    y = _lib.require_cuda(y, "stft input")
    if num_dims == 3:
        y = y.reshape(-1, num_samples)
    B = y.shape[0]
    F, T = n_fft // 2 + 1, 1 + num_samples // hop_length
    out = torch.empty(4, B, F, T, dtype=torch.float32, device=y.device)
    with torch.cuda.device(y.device):
        _lib.check(_lib.load().fsn_stft(y.data_ptr(), B, num_samples, n_fft, hop_length, win_length,
                                        out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), out[3].data_ptr(),
                                        None, 0, _lib.stream_ptr(y.device)))
    mag, phase, real, imag = out[0], out[1], out[2], out[3]
    if num_dims == 3:
        mag, phase, real, imag = (t.reshape(batch_size, -1, F, T) for t in (mag, phase, real, imag))
    return mag, phase, real, imag


def istft(features, n_fft, hop_length, win_length, length=None, input_type="complex"):
    """[B,F,T] spectrum -> [B,L] wave, torch.istft semantics (feature.py:53-91)."""
    if input_type == "real_imag":
        assert isinstance(features, tuple) or isinstance(features, list)
        real, imag = features
        real = _lib.require_cuda(real, "istft real")
        imag = _lib.require_cuda(imag, "istft imag")
        cstride = 1
    elif input_type == "complex":
        assert torch.is_complex(features), "The input feature is not complex."
        if not features.is_cuda:
            raise RuntimeError("fullsubnet_b200: istft input must be a CUDA tensor; this package has no CPU path.")
        ri = torch.view_as_real(features.to(torch.complex64).contiguous())  # [B,F,T,2] interleaved
        real, imag = ri[..., 0], ri[..., 1]
        cstride = 2
    elif input_type == "mag_phase":
        assert isinstance(features, tuple) or isinstance(features, list)
        mag, phase = features
        mag = _lib.require_cuda(mag, "istft mag")
        real, imag = mag * torch.cos(phase), mag * torch.sin(phase)  # feature.py:78
        cstride = 1
    else:
        raise NotImplementedError("Only 'real_imag', 'complex', and 'mag_phase' are supported.")
    assert real.dim() == 3, "istft expects [B, F, T]"
    B, F, T = real.shape
    assert F == n_fft // 2 + 1, f"istft: F = {F} != n_fft // 2 + 1"
    out_len = int(length) if length is not None else hop_length * (T - 1)
    wav = torch.empty(B, out_len, dtype=torch.float32, device=real.device)
    with torch.cuda.device(real.device):
        _lib.check(_lib.load().fsn_istft(real.data_ptr(), imag.data_ptr(), cstride, None, B, T, n_fft, hop_length,
                                         win_length, out_len, wav.data_ptr(), _lib.stream_ptr(real.device)))
    return wav


def mag_phase(complex_tensor):
    """feature.py:94-96 (API helper, not on the timed path)."""
    return torch.abs(complex_tensor), torch.angle(complex_tensor)


def drop_band(input, num_groups=2):
    """[B,C,F,T] -> [B,C,F//num_groups,T] with the reference's batch re-ordering (feature.py:309-345)."""
    batch_size, C, num_freqs, T = input.shape
    assert batch_size > num_groups, (
        f"Batch size = {batch_size}, num_groups = {num_groups}. "
        "The batch size should larger than the num_groups.")
    if num_groups <= 1:
        return input
    x = _lib.require_cuda(input, "drop_band input")
    out = torch.empty(batch_size, C, num_freqs // num_groups, T, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().fsn_drop_band(x.data_ptr(), out.data_ptr(), batch_size, C, num_freqs, T, num_groups,
                                             _lib.stream_ptr(x.device)))
    return out

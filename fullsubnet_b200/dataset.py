"""Training-data mixing on the device: the arithmetic of ``Dataset.snr_mix``
(recipes/dns_interspeech_2020/dataset_train.py:136-199) for a batch of (clean, noise) pairs, so that an 8-GPU trainer
does not need the 16-48 CPU dataloader workers per GPU the reference's on-the-fly mixing would take.
File selection, cropping and the random draws stay on the host (cheap); they are passed in as tensors."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib


def snr_mix(clean_y: torch.Tensor, noise_y: torch.Tensor, snr, target_dB_FS: float, noisy_target_dB_FS,
            rir: Optional[torch.Tensor] = None, rir_len: Optional[torch.Tensor] = None, eps: float = 1e-6):
    """clean_y, noise_y [B,L] (CUDA float32); snr, noisy_target_dB_FS: [B] (the values the reference draws with
    ``random.choice(snr_list)`` and ``np.random.randint(target - floating, target + floating)``); rir [B,Lr] with
    rir_len [B] int32 (0 = no reverberation for that clip) or None.  Returns (noisy_y, clean_y), both [B,L]."""
    clean_y = _lib.require_cuda(clean_y, "clean_y")
    noise_y = _lib.require_cuda(noise_y, "noise_y")
    assert clean_y.shape == noise_y.shape and clean_y.dim() == 2, "Inequality: clean / noise shapes"
    B, L = clean_y.shape
    dev = clean_y.device
    snr_t = torch.as_tensor(snr, dtype=torch.float32, device=dev).reshape(-1).expand(B).contiguous()
    nt_t = torch.as_tensor(noisy_target_dB_FS, dtype=torch.float32, device=dev).reshape(-1).expand(B).contiguous()
    lib = _lib.load()
    with torch.cuda.device(dev):
        st = _lib.stream_ptr(dev)
        if rir is not None:
            rir = _lib.require_cuda(rir, "rir")
            assert rir.dim() == 2 and rir.shape[0] == B
            rl = None if rir_len is None else rir_len.to(device=dev, dtype=torch.int32).contiguous()
            rev = torch.empty_like(clean_y)
            _lib.check(lib.fsn_rir_convolve(clean_y.data_ptr(), rir.data_ptr(), _lib.ptr(rl), B, L, rir.shape[1],
                                            rev.data_ptr(), st))
            clean_y = rev
        noisy, clean = torch.empty_like(clean_y), torch.empty_like(clean_y)
        _lib.check(lib.fsn_snr_mix(clean_y.data_ptr(), noise_y.data_ptr(), snr_t.data_ptr(), nt_t.data_ptr(),
                                   float(target_dB_FS), float(eps), B, L, noisy.data_ptr(), clean.data_ptr(), st))
    return noisy, clean

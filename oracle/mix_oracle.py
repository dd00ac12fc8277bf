"""TEST INFRASTRUCTURE ONLY - numpy restatement of Dataset.snr_mix
(recipes/dns_interspeech_2020/dataset_train.py:136-199) and the helpers it calls
(audio_zen/acoustics/feature.py:99-114: norm_amplitude, tailor_dB_FS, is_clipped), with the reference's random draws
turned into arguments.  Pinned by tests/golden/mix.npz (oracle/make_golden_mix.py runs the unmodified reference)."""
from __future__ import annotations

import numpy as np


def snr_mix(clean_y, noise_y, snr, target_dB_FS, noisy_target_dB_FS, rir=None, eps=1e-6):
    clean_y = np.asarray(clean_y, dtype=np.float32).copy()
    noise_y = np.asarray(noise_y, dtype=np.float32).copy()
    L = len(clean_y)
    if rir is not None:  # dataset_train.py:161 (fftconvolve, first L samples) as a direct convolution in float64
        clean_y = np.convolve(clean_y.astype(np.float64), np.asarray(rir, dtype=np.float64))[:L].astype(np.float32)

    def norm_and_tailor(y):  # feature.py:99-111
        y = y / (np.max(np.abs(y)) + eps)
        rms = np.sqrt(np.mean(y ** 2))
        return y * (10 ** (target_dB_FS / 20) / (rms + eps))

    clean_y = norm_and_tailor(clean_y)
    noise_y = norm_and_tailor(noise_y)
    clean_rms = np.sqrt(np.mean(clean_y ** 2))
    noise_rms = np.sqrt(np.mean(noise_y ** 2))
    noise_y = noise_y * (clean_rms / (10 ** (snr / 20)) / (noise_rms + eps))  # dataset_train.py:171-172
    noisy_y = clean_y + noise_y
    k = 10 ** (noisy_target_dB_FS / 20) / (np.sqrt(np.mean(noisy_y ** 2)) + eps)  # :182-183
    noisy_y, clean_y = noisy_y * k, clean_y * k
    if np.any(np.abs(noisy_y) > 0.999):  # :186-189
        s = np.max(np.abs(noisy_y)) / (0.99 - eps)
        noisy_y, clean_y = noisy_y / s, clean_y / s
    return noisy_y.astype(np.float32), clean_y.astype(np.float32)

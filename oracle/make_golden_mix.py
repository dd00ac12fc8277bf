"""TEST INFRASTRUCTURE ONLY.  ``tests/golden/mix.npz``: the UNMODIFIED ``Dataset.snr_mix``
(recipes/dns_interspeech_2020/dataset_train.py:136-199) on CPU for a few (clean, noise, snr, rir) cases; the
``np.random.randint`` draw of the noisy target dBFS inside it is pinned by a patched generator and stored.
Run:  python oracle/make_golden_mix.py
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def main():
    from make_golden import import_reference
    import_reference()
    from dataset_train import Dataset  # the reference class (recipes/dns_interspeech_2020 is on sys.path)
    from oracle import fullsubnet_oracle as O
    rng = np.random.default_rng(17)
    res, cases = {}, []
    L = 12000
    for i, (snr, draw, rir_len, gain) in enumerate([(5, -30, 0, 1.0), (-5, -17, 0, 1.0), (20, -16, 600, 1.0),
                                                    (0, -16, 2500, 30.0)]):
        clean = O.make_noisy(1, L, seed=60 + i, speechlike=True)[0].numpy() * np.float32(gain)
        if i == 1:  # impulsive clip: crest factor high enough that the mixture exceeds 0.999 -> the anti-clipping branch
            clean[::1500] += np.float32(40.0)
        noise = (0.3 * rng.standard_normal(L)).astype(np.float32)
        rir = None
        if rir_len:
            t = np.arange(rir_len)
            rir = (rng.standard_normal(rir_len) * np.exp(-t / (rir_len / 6.0))).astype(np.float32)
            rir[0] = 1.0
        orig = np.random.randint
        np.random.randint = lambda *a, **k: draw  # the only draw on this path: noisy_target_dB_FS (rir is 1-D)
        try:
            noisy, clean_out = Dataset.snr_mix(clean.copy(), noise.copy(), snr, -25, 10, rir=None if rir is None else rir.copy())
        finally:
            np.random.randint = orig
        res[f"c{i}_clean"], res[f"c{i}_noise"] = clean, noise
        res[f"c{i}_rir"] = rir if rir is not None else np.zeros(0, np.float32)
        res[f"c{i}_noisy"], res[f"c{i}_clean_out"] = noisy.astype(np.float32), clean_out.astype(np.float32)
        cases.append((snr, draw, rir_len))
        print(i, "snr", snr, "draw", draw, "rir", rir_len, "max|noisy|", float(np.abs(noisy).max()))
    res["cases"] = np.asarray(cases, dtype=np.int64)
    out = os.path.join(ROOT, "tests", "golden", "mix.npz")
    np.savez_compressed(out, **res)
    print(out, os.path.getsize(out))


if __name__ == "__main__":
    main()

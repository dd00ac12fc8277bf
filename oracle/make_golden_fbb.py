"""TEST INFRASTRUCTURE ONLY.  ``tests/golden/fullband_baseline.npz``: the UNMODIFIED upstream fullband_baseline Model
(recipes/dns_interspeech_2020/fullband_baseline/model.py) on CPU: small (F=33, H=32, ReLU, cumulative norm) and
full-size (F=257, H=512, offline norm) configurations.   Run:  python oracle/make_golden_fbb.py
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def main():
    from make_golden import REF, import_reference
    from oracle import fullband_baseline_oracle as BO
    from oracle import fullsubnet_oracle as O
    feature, _, _, _ = import_reference()
    spec = importlib.util.spec_from_file_location(
        "fbb_model", os.path.join(REF, "recipes", "dns_interspeech_2020", "fullband_baseline", "model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.set_num_threads(8)
    res = {}
    cfgs = {"small": dict(BO.DEFAULT_FBB_ARGS, num_freqs=33, hidden_size=32, output_activate_function="ReLU",
                          norm_type="cumulative_laplace_norm"),
            "full": dict(BO.DEFAULT_FBB_ARGS)}
    for tag, a in cfgs.items():
        sd = BO.make_fbb_state_dict(seed=11, args=a)
        m = mod.Model(**a).eval()
        assert [k for k, _ in BO.fbb_state_dict_shapes(a)] == list(m.state_dict().keys())
        m.load_state_dict(sd, strict=True)
        n_fft = 64 if tag == "small" else 512
        y = O.make_noisy(3, 1200 if tag == "small" else 5000, seed=19, speechlike=True)
        mag = feature.stft(y, n_fft, n_fft // 2, n_fft)[0].unsqueeze(1)
        with torch.no_grad():
            out = m(mag)
        res[tag + "_mag"], res[tag + "_out"] = mag.numpy(), out.numpy()
        print(tag, tuple(out.shape), float(out.abs().max()))
    out = os.path.join(ROOT, "tests", "golden", "fullband_baseline.npz")
    np.savez_compressed(out, **res)
    print(out, os.path.getsize(out))


if __name__ == "__main__":
    main()

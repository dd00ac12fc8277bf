"""TEST INFRASTRUCTURE ONLY.  ``tests/golden/improved_960.npz``: the UNMODIFIED upstream improved_fullsubnet Model
with the reference's own 48 kHz example arguments (improved_fullsubnet/model.py:603-620, n_fft = 960, hop = 480) on
CPU, plus reference STFT / iSTFT vectors at that transform size.   Run:  python oracle/make_golden_imp960.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def main():
    from make_golden import import_reference
    from oracle import fullsubnet_oracle as O
    from oracle import improved_fullsubnet_oracle as IO
    feature, _, _, _ = import_reference()
    from improved_fullsubnet.model import Model as ImpModel
    torch.set_num_threads(8)
    args = dict(IO.ARGS_48K_960)
    sd = IO.make_improved_state_dict(seed=5, args=args)
    m = ImpModel(**args).eval()
    assert [k for k, _ in IO.improved_state_dict_shapes(args)] == list(m.state_dict().keys())
    m.load_state_dict(sd, strict=True)
    y = O.make_noisy(2, 12000, seed=17, speechlike=True)
    with torch.no_grad():
        wav = m(y)
    mag, _, re, im = feature.stft(y, 960, 480, 960)
    back = feature.istft((re * 0.5 - im * 0.25, im * 0.5 + re * 0.25), 960, 480, 960, length=12000, input_type="real_imag")
    out = os.path.join(ROOT, "tests", "golden", "improved_960.npz")
    np.savez_compressed(out, y=y.numpy(), wav=wav.numpy(), mag=mag.numpy(), real=re.numpy(), imag=im.numpy(),
                        back=back.numpy())
    print(out, os.path.getsize(out), float(wav.abs().max()), sum(v.numel() for v in sd.values()))


if __name__ == "__main__":
    main()

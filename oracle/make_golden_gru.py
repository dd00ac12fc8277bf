"""TEST INFRASTRUCTURE ONLY.  ``tests/golden/model_gru.npz``: the UNMODIFIED upstream fullsubnet Model with
``sequence_model="GRU"`` (audio_zen/model/module/sequence_model.py:59-66) on CPU: a small model
(B=1 and B=3 with drop_band) and the full-size architecture through Inferencer.full_band_crm_mask on 2 x 0.5 s.
Run:  python oracle/make_golden_gru.py
"""
from __future__ import annotations

import os
import sys
from functools import partial

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def main():
    from make_golden import import_reference
    from oracle import fullsubnet_oracle as O
    feature, mask, Model, Inferencer = import_reference()
    torch.set_num_threads(8)
    small = dict(num_freqs=33, look_ahead=2, sequence_model="GRU", fb_num_neighbors=0, sb_num_neighbors=3,
                 fb_output_activate_function="ReLU", sb_output_activate_function=False,
                 fb_model_hidden_size=32, sb_model_hidden_size=24, norm_type="offline_laplace_norm",
                 num_groups_in_drop_band=2, weight_init=False)
    sd = O.make_state_dict(seed=7, args=small)
    model = Model(**small).eval()
    model.load_state_dict(sd, strict=True)
    ys = O.make_noisy(3, 1200, seed=4, speechlike=True)
    mag = feature.stft(ys, 64, 32, 64)[0]
    res = {"small_mag": mag.numpy()}
    with torch.no_grad():
        res["small_b1"] = model(mag[:1].unsqueeze(1)).numpy()
        res["small_g2"] = model(mag.unsqueeze(1)).numpy()
    full = dict(O.DEFAULT_MODEL_ARGS, sequence_model="GRU")
    sdf = O.make_state_dict(seed=0, args=full, sb_fc_gain=60.0)
    model = Model(**full).eval()
    model.load_state_dict(sdf, strict=True)
    inf = Inferencer.__new__(Inferencer)
    inf.model, inf.device = model, torch.device("cpu")
    inf.torch_stft = partial(feature.stft, n_fft=512, hop_length=256, win_length=512)
    inf.torch_istft = partial(feature.istft, n_fft=512, hop_length=256, win_length=512)
    y = O.make_noisy(2, 8000, seed=9, speechlike=True)
    with torch.no_grad():
        magf = feature.stft(y, 512, 256, 512)[0]
        res["full_crm"] = torch.cat([model(magf[i:i + 1].unsqueeze(1)) for i in range(2)], 0).numpy()
        res["full_wav"] = np.stack([inf.full_band_crm_mask(y[i:i + 1], {}) for i in range(2)], 0)
    res["full_y"] = y.numpy()
    print("full crm range", res["full_crm"].min(), res["full_crm"].max())
    out = os.path.join(ROOT, "tests", "golden", "model_gru.npz")
    np.savez_compressed(out, **res)
    print(out, os.path.getsize(out))


if __name__ == "__main__":
    main()

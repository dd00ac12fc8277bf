"""TEST INFRASTRUCTURE ONLY.  Generates ``tests/golden/train_{small,full}.npz``: one optimisation step of
recipes/dns_interspeech_2020/fullsubnet/trainer.py:41-68 run with the UNMODIFIED upstream Model / stft / cIRM /
drop_band from ``/root/reference`` on CPU (no DDP, AMP off - SURVEY 8c recipe), torch.nn.MSELoss
(audio_zen/loss.py:4), clip_grad_norm_(10) and Adam(lr 1e-3, betas (0.9, 0.999)) (train.py:55-59).

Run:  python oracle/make_golden_train.py
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

SMALL = dict(num_freqs=33, look_ahead=2, sequence_model="LSTM", fb_num_neighbors=0, sb_num_neighbors=3,
             fb_output_activate_function="ReLU", sb_output_activate_function=False,
             fb_model_hidden_size=32, sb_model_hidden_size=24, norm_type="offline_laplace_norm",
             num_groups_in_drop_band=2, weight_init=False)
SUBSAMPLE = 97  # full-size gradients are stored as every 97th element + the per-tensor L2 norm


def reference_step(feature, mask, Model, args, sd, noisy, clean, n_fft, hop, steps=2):
    model = Model(**args).train()
    model.load_state_dict(sd, strict=True)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.999))
    loss_fn = torch.nn.MSELoss()
    out = {}
    for it in range(steps):
        opt.zero_grad()
        noisy_mag, _, nr, ni = feature.stft(noisy, n_fft, hop, n_fft)
        _, _, cr, ci = feature.stft(clean, n_fft, hop, n_fft)
        cirm = mask.build_complex_ideal_ratio_mask(nr, ni, cr, ci)
        cirm = feature.drop_band(cirm.permute(0, 3, 1, 2), model.num_groups_in_drop_band).permute(0, 2, 3, 1)
        crm = model(noisy_mag.unsqueeze(1)).permute(0, 2, 3, 1)
        loss = loss_fn(cirm, crm)
        loss.backward()
        if it == 0:
            out["cirm"] = cirm.detach().numpy().copy()
            out["crm"] = crm.detach().numpy().copy()
            out["grads"] = {k: p.grad.detach().numpy().copy() for k, p in model.named_parameters()}
        out[f"loss{it}"] = float(loss)
        out[f"gnorm{it}"] = float(torch.nn.utils.clip_grad_norm_(model.parameters(), 10))
        opt.step()
        out[f"params{it}"] = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
    return out


def main():
    from make_golden import import_reference
    from oracle import fullsubnet_oracle as O
    feature, mask, Model, _ = import_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    torch.set_num_threads(8)

    # small model, 5 clips (odd: groups of 3 and 2), G = 2
    sd = O.make_state_dict(seed=7, args=SMALL, sb_fc_gain=8.0)
    noisy = O.make_noisy(5, 1200, seed=21, speechlike=True)
    clean = 0.5 * O.make_noisy(5, 1200, seed=22, speechlike=True)
    r = reference_step(feature, mask, Model, SMALL, sd, noisy, clean, 64, 32)
    print("small: loss", r["loss0"], r["loss1"], "gnorm", r["gnorm0"], r["gnorm1"])
    np.savez_compressed(
        os.path.join(out_dir, "train_small.npz"), noisy=noisy.numpy(), clean=clean.numpy(), cirm=r["cirm"], crm=r["crm"],
        loss=np.array([r["loss0"], r["loss1"]]), gnorm=np.array([r["gnorm0"], r["gnorm1"]]),
        **{"grad." + k: v for k, v in r["grads"].items()},
        **{"p0." + k: v for k, v in r["params0"].items()}, **{"p1." + k: v for k, v in r["params1"].items()})

    # full-size model (config 3 architecture), 4 clips x 0.25 s
    full = dict(O.DEFAULT_MODEL_ARGS)
    full["weight_init"] = False
    sd = O.make_state_dict(seed=0, args=full, sb_fc_gain=40.0)
    noisy = O.make_noisy(4, 4000, seed=31, speechlike=True)
    clean = 0.5 * O.make_noisy(4, 4000, seed=32, speechlike=True)
    r = reference_step(feature, mask, Model, full, sd, noisy, clean, 512, 256, steps=1)
    print("full: loss", r["loss0"], "gnorm", r["gnorm0"])
    np.savez_compressed(
        os.path.join(out_dir, "train_full.npz"), noisy=noisy.numpy(), clean=clean.numpy(),
        loss=np.array([r["loss0"]]), gnorm=np.array([r["gnorm0"]]),
        **{"gsub." + k: v.reshape(-1)[::SUBSAMPLE] for k, v in r["grads"].items()},
        **{"gl2." + k: np.array(np.sqrt((v.astype(np.float64) ** 2).sum())) for k, v in r["grads"].items()},
        **{"psub." + k: v.reshape(-1)[::SUBSAMPLE] for k, v in r["params0"].items()})
    for f in ("train_small.npz", "train_full.npz"):
        print(f, os.path.getsize(os.path.join(out_dir, f)))


if __name__ == "__main__":
    main()

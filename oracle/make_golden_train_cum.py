"""TEST INFRASTRUCTURE ONLY.  ``tests/golden/train_cum_small.npz``: two optimisation steps of the UNMODIFIED reference
(trainer.py:41-68 arithmetic, make_golden_train.reference_step) with ``norm_type="cumulative_laplace_norm"`` - the
norm two shipped TOMLs train with (fullsubnet/train_cumulativeLaplaceNorm.toml:82).  Run: python oracle/make_golden_train_cum.py
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
    from make_golden_train import SMALL, reference_step
    from oracle import fullsubnet_oracle as O
    feature, mask, Model, _ = import_reference()
    torch.set_num_threads(8)
    args = dict(SMALL, norm_type="cumulative_laplace_norm")
    sd = O.make_state_dict(seed=7, args=args, sb_fc_gain=8.0)
    noisy = O.make_noisy(5, 1200, seed=21, speechlike=True)
    clean = 0.5 * O.make_noisy(5, 1200, seed=22, speechlike=True)
    r = reference_step(feature, mask, Model, args, sd, noisy, clean, 64, 32)
    print("small cum: loss", r["loss0"], r["loss1"], "gnorm", r["gnorm0"], r["gnorm1"])
    out = os.path.join(ROOT, "tests", "golden", "train_cum_small.npz")
    np.savez_compressed(
        out, noisy=noisy.numpy(), clean=clean.numpy(), crm=r["crm"],
        loss=np.array([r["loss0"], r["loss1"]]), gnorm=np.array([r["gnorm0"], r["gnorm1"]]),
        **{"grad." + k: v for k, v in r["grads"].items()},
        **{"p0." + k: v for k, v in r["params0"].items()}, **{"p1." + k: v for k, v in r["params1"].items()})
    print(out, os.path.getsize(out))


if __name__ == "__main__":
    main()

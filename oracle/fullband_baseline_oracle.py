"""TEST INFRASTRUCTURE ONLY - CPU restatement of recipes/dns_interspeech_2020/fullband_baseline/model.py:8-68
(SURVEY 8f rank 3), pinned against the unmodified reference by tests/golden/fullband_baseline.npz
(oracle/make_golden_fbb.py)."""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from . import fullsubnet_oracle as O

DEFAULT_FBB_ARGS = dict(num_freqs=257, hidden_size=512, sequence_model="LSTM", output_activate_function=False,
                        look_ahead=2, norm_type="offline_laplace_norm", weight_init=False)


def fbb_forward(noisy_mag: torch.Tensor, sd: Dict[str, torch.Tensor], args: Optional[dict] = None) -> torch.Tensor:
    """model.py:46-68: noisy_mag [B,1,F,T] -> [B,2,F,T]."""
    a = dict(DEFAULT_FBB_ARGS)
    a.update(args or {})
    norm = {"offline_laplace_norm": O.offline_laplace_norm, "cumulative_laplace_norm": O.cumulative_laplace_norm}[a["norm_type"]]
    assert noisy_mag.dim() == 4
    la = a["look_ahead"]
    x = torch.nn.functional.pad(noisy_mag, [0, la])
    B, C, F, T = x.shape
    assert C == 1
    x = norm(x).reshape(B, F, T)
    o = O.lstm_stack(x.permute(0, 2, 1), sd, "fullband_model.sequence_model.", num_layers=3)
    o = o @ sd["fullband_model.fc_output_layer.weight"].t() + sd["fullband_model.fc_output_layer.bias"]
    act = a["output_activate_function"]
    if act:
        o = {"ReLU": torch.relu, "Tanh": torch.tanh, "ReLU6": lambda v: torch.clamp(v, 0, 6)}[act](o)
    return o.permute(0, 2, 1).reshape(B, 2, F, T)[:, :, :, la:]


def fbb_state_dict_shapes(args: Optional[dict] = None):
    a = dict(DEFAULT_FBB_ARGS)
    a.update(args or {})
    F, H = a["num_freqs"], a["hidden_size"]
    out = []
    for l in range(3):
        k = F if l == 0 else H
        out += [(f"fullband_model.sequence_model.weight_ih_l{l}", (4 * H, k)),
                (f"fullband_model.sequence_model.weight_hh_l{l}", (4 * H, H)),
                (f"fullband_model.sequence_model.bias_ih_l{l}", (4 * H,)),
                (f"fullband_model.sequence_model.bias_hh_l{l}", (4 * H,))]
    return out + [("fullband_model.fc_output_layer.weight", (2 * F, H)), ("fullband_model.fc_output_layer.bias", (2 * F,))]


def make_fbb_state_dict(seed: int = 0, args: Optional[dict] = None) -> Dict[str, torch.Tensor]:
    rng = np.random.default_rng(seed)
    sd = {}
    for name, shape in fbb_state_dict_shapes(args):
        k = 1.0 / math.sqrt(shape[0] // 4) if "sequence_model" in name else 1.0 / math.sqrt(
            shape[1] if len(shape) == 2 else sd[name.replace("bias", "weight")].shape[1])
        sd[name] = torch.from_numpy(rng.uniform(-k, k, size=shape).astype(np.float32))
    return sd

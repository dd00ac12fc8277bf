"""TEST INFRASTRUCTURE ONLY - CPU fp32 restatement of recipes/dns_interspeech_2020/fast_fullsubnet/model.py
(BASELINE config 4).  Pinned against the unmodified reference through
``oracle/make_golden.py`` -> ``tests/golden/fast_small.npz`` / ``fast_full.npz``."""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from .fullsubnet_oracle import freq_unfold, lstm_stack, offline_laplace_norm

DEFAULT_FAST_ARGS = dict(  # fast_fullsubnet/inference.toml:30-38
    look_ahead=2, shrink_size=2, sequence_model="LSTM", num_mels=64, encoder_input_size=257,
    bottleneck_hidden_size=384, bottleneck_num_layers=2, noisy_input_num_neighbors=5,
    encoder_output_num_neighbors=0, norm_type="offline_laplace_norm", weight_init=False,
)


def melscale_fbanks(n_freqs: int, n_mels: int, sample_rate: int = 16000, f_min: float = 0.0,
                    f_max: float = 8000.0) -> torch.Tensor:
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk') restated
    (the buffer ``mel_scale.fb`` [n_freqs, n_mels] of fast_fullsubnet/model.py:57-63)."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


def real_time_downsampling(x: torch.Tensor, shrink: int) -> torch.Tensor:
    """fast_fullsubnet/model.py:108-129: first frame alone, then means of blocks of `shrink` frames; the last
    (possibly shorter) block is averaged over its own length."""
    first = x[..., 0:1]
    rest = x[..., 1:]
    n = rest.shape[-1]
    nfull = (n - 1) // shrink  # blocks before the last one (all of full size)
    parts = [first]
    if nfull > 0:
        parts.append(rest[..., : nfull * shrink].reshape(*rest.shape[:-1], nfull, shrink).mean(dim=-1))
    parts.append(rest[..., nfull * shrink:].mean(dim=-1, keepdim=True))
    return torch.cat(parts, dim=-1)


def real_time_upsampling(x: torch.Tensor, shrink: int, target_len: int) -> torch.Tensor:
    """fast_fullsubnet/model.py:131-140: repeat every frame `shrink` times, crop to target_len."""
    return x.repeat_interleave(shrink, dim=-1)[..., :target_len]


def _seq(x, sd, prefix, layers, act, has_fc):
    """SequenceModel (audio_zen/model/module/sequence_model.py:106-125), x [B,F,T] -> [B,F_out,T]."""
    o = lstm_stack(x.permute(0, 2, 1), sd, prefix + "sequence_model.", num_layers=layers)
    if has_fc:
        o = o @ sd[prefix + "fc_output_layer.weight"].t() + sd[prefix + "fc_output_layer.bias"]
    if act == "ReLU":
        o = torch.relu(o)
    return o.permute(0, 2, 1)


def fast_model_forward(mix_mag: torch.Tensor, sd: Dict[str, torch.Tensor], args: Optional[dict] = None,
                       return_intermediates: bool = False):
    """fast_fullsubnet/model.py:143-202.  mix_mag [B,1,F,T] -> [B,2,F,T] (no drop_band in this model)."""
    a = dict(DEFAULT_FAST_ARGS)
    a.update(args or {})
    assert mix_mag.dim() == 4
    la, S = a["look_ahead"], a["shrink_size"]
    Nn, Ne, M = a["noisy_input_num_neighbors"], a["encoder_output_num_neighbors"], a["num_mels"]
    x = torch.nn.functional.pad(mix_mag, [0, la])
    B, C, F, T = x.shape
    assert C == 1
    mel = (x.transpose(-1, -2) @ sd["mel_scale.fb"]).transpose(-1, -2)  # [B,1,M,T]  (model.py:166)
    enc_in = offline_laplace_norm(mel).reshape(B, -1, T)
    e1 = _seq(enc_in, sd, "encoder.0.", 1, None, False)
    enc_out = _seq(e1, sd, "encoder.1.", 1, "ReLU", True).reshape(B, 1, -1, T)  # [B,1,M,T]
    mel_unf = freq_unfold(mel, Nn).reshape(B, M, 2 * Nn + 1, T)
    enc_unf = freq_unfold(enc_out, Ne).reshape(B, M, 2 * Ne + 1, T)
    bn_in = torch.cat([mel_unf, enc_unf], dim=2)
    K = bn_in.shape[2]
    bn_shr = offline_laplace_norm(real_time_downsampling(bn_in, S))
    bn_shr_rows = bn_shr.reshape(B * M, K, -1)
    bn_out = _seq(bn_shr_rows, sd, "bottleneck.", a["bottleneck_num_layers"], "ReLU", True)  # [B*M,1,Ts]
    bn_out = bn_out.reshape(B, M, 1, -1).permute(0, 2, 1, 3)
    bn_up = real_time_upsampling(bn_out, S, T)  # [B,1,M,T]
    dec_in = torch.cat([enc_out, bn_up], dim=2).reshape(B, -1, T)
    d1 = _seq(dec_in, sd, "decoder_lstm.0.", 1, None, False)
    d2 = _seq(d1, sd, "decoder_lstm.1.", 1, None, True)  # [B, 2F, T]
    out = d2.reshape(B, 2, F, T)[:, :, :, la:]
    if return_intermediates:
        return out, dict(mel=mel, enc_out=enc_out, bn_shr=bn_shr, bn_up=bn_up)
    return out


def fast_state_dict_shapes(args: Optional[dict] = None):
    a = dict(DEFAULT_FAST_ARGS)
    a.update(args or {})
    M, F, Hb = a["num_mels"], a["encoder_input_size"], a["bottleneck_hidden_size"]
    K = (2 * a["noisy_input_num_neighbors"] + 1) + (2 * a["encoder_output_num_neighbors"] + 1)

    def lstm(pre, l, In, H):
        return [(f"{pre}sequence_model.weight_ih_l{l}", (4 * H, In)), (f"{pre}sequence_model.weight_hh_l{l}", (4 * H, H)),
                (f"{pre}sequence_model.bias_ih_l{l}", (4 * H,)), (f"{pre}sequence_model.bias_hh_l{l}", (4 * H,))]

    out = lstm("encoder.0.", 0, 64, 384)
    out += lstm("encoder.1.", 0, 384, 257) + [("encoder.1.fc_output_layer.weight", (64, 257)),
                                             ("encoder.1.fc_output_layer.bias", (64,))]
    out += [("mel_scale.fb", (F, M))]
    for l in range(a["bottleneck_num_layers"]):
        out += lstm("bottleneck.", l, K if l == 0 else Hb, Hb)
    out += [("bottleneck.fc_output_layer.weight", (1, Hb)), ("bottleneck.fc_output_layer.bias", (1,))]
    out += lstm("decoder_lstm.0.", 0, 64 + 64, 512)
    out += lstm("decoder_lstm.1.", 0, 512, 512) + [("decoder_lstm.1.fc_output_layer.weight", (2 * F, 512)),
                                                  ("decoder_lstm.1.fc_output_layer.bias", (2 * F,))]
    return out


def make_fast_state_dict(seed: int = 0, args: Optional[dict] = None) -> Dict[str, torch.Tensor]:
    """PyTorch-default-like uniform init from numpy PCG64 (version-stable); mel filterbank from the formula."""
    a = dict(DEFAULT_FAST_ARGS)
    a.update(args or {})
    rng = np.random.default_rng(seed)
    sd = {}
    for name, shape in fast_state_dict_shapes(a):
        if name == "mel_scale.fb":
            sd[name] = melscale_fbanks(a["encoder_input_size"], a["num_mels"])
            continue
        if "sequence_model" in name:
            k = 1.0 / math.sqrt(shape[0] // 4)
        elif name.endswith("fc_output_layer.weight"):
            k = 1.0 / math.sqrt(shape[1])
        else:
            k = 1.0 / math.sqrt(sd[name.replace("bias", "weight")].shape[1])
        sd[name] = torch.from_numpy(rng.uniform(-k, k, size=shape).astype(np.float32))
    return sd

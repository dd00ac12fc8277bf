"""TEST INFRASTRUCTURE ONLY - CPU restatement of one optimisation step of
recipes/dns_interspeech_2020/fullsubnet/trainer.py:41-68, pinned against the unmodified
reference by tests/golden/train_{small,full}.npz (oracle/make_golden_train.py).

Two restatements:
  * ``train_step``      - forward = oracle.model_forward, gradients by torch autograd, then clip_grad_norm_ and Adam
                          restated by hand (torch/nn/utils/clip_grad.py, torch/optim/adam.py single-tensor path);
  * ``manual_backward`` - the hand-derived BPTT the CUDA kernels implement (time-major saved activations, closed-form
                          gradient of the second laplace norm, drop_band as a row map).  No autograd.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import fullsubnet_oracle as O

PARAM_ORDER = [k for k, _ in O.state_dict_shapes()]


def targets(noisy, clean, G, n_fft=512, hop=256, win=512):
    """trainer.py:46-54: noisy magnitude and the compressed, drop_band'ed cIRM target [B,F',T,2]."""
    nm, _, nr, ni = O.stft(noisy, n_fft, hop, win)
    _, _, cr, ci = O.stft(clean, n_fft, hop, win)
    cirm = O.build_complex_ideal_ratio_mask(nr, ni, cr, ci)
    cirm = O.drop_band(cirm.permute(0, 3, 1, 2), G).permute(0, 2, 3, 1)
    return nm, cirm


def loss_and_grads(noisy_mag, cirm, sd, args=None):
    """trainer.py:56-63 without autocast: MSE(cIRM, cRM) and d loss / d parameter by autograd."""
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    crm = O.model_forward(noisy_mag.unsqueeze(1), p, args).permute(0, 2, 3, 1)
    loss = torch.mean((cirm - crm) ** 2)  # audio_zen/loss.py:4
    loss.backward()
    return loss.detach(), {k: v.grad for k, v in p.items()}, crm.detach()


def clip_coef(grads: Dict[str, torch.Tensor], max_norm: float):
    """torch.nn.utils.clip_grad_norm_ (trainer.py:65-67): total L2 norm, coef = min(1, max_norm / (norm + 1e-6))."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    return total, torch.clamp(max_norm / (total + 1e-6), max=1.0)


def adam_update(sd, grads, state, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
    """torch.optim.Adam defaults (train.py:55-59): no weight decay, no amsgrad."""
    step = state.get("step", 0) + 1
    b1, b2 = betas
    new_sd, m, v = {}, {}, {}
    for k in sd:
        g = grads[k]
        m[k] = state["m"][k] * b1 + (1 - b1) * g if "m" in state else (1 - b1) * g
        v[k] = state["v"][k] * b2 + (1 - b2) * g * g if "v" in state else (1 - b2) * g * g
        bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
        denom = v[k].sqrt() / np.sqrt(bc2) + eps
        new_sd[k] = sd[k] - (lr / bc1) * m[k] / denom
    return new_sd, dict(step=step, m=m, v=v)


def train_step(noisy, clean, sd, args=None, state=None, n_fft=512, hop=256, win=512, max_norm=10.0, lr=1e-3):
    a = dict(O.DEFAULT_MODEL_ARGS)
    a.update(args or {})
    nm, cirm = targets(noisy, clean, a["num_groups_in_drop_band"], n_fft, hop, win)
    loss, grads, crm = loss_and_grads(nm, cirm, sd, a)
    gnorm, coef = clip_coef(grads, max_norm)
    clipped = {k: g * coef for k, g in grads.items()}
    new_sd, new_state = adam_update(sd, clipped, state or {}, lr=lr)
    return dict(loss=loss, grads=grads, gnorm=gnorm, sd=new_sd, state=new_state, cirm=cirm, crm=crm)


# --------------------------------------------------------------------------------------------------------------
def _lstm_fwd_save(X, w_ih, w_hh, b_ih, b_hh):
    """X [Tp,R,K] -> saved gates (post-activation) [Tp,R,4H], cell [Tp,R,H], hidden [Tp,R,H]."""
    Tp, R, _ = X.shape
    H = w_hh.shape[1]
    G, Cc, Hh = torch.zeros(Tp, R, 4 * H), torch.zeros(Tp, R, H), torch.zeros(Tp, R, H)
    h, c = torch.zeros(R, H), torch.zeros(R, H)
    for t in range(Tp):
        z = X[t] @ w_ih.T + h @ w_hh.T + b_ih + b_hh
        i, f, g, o = z.split(H, dim=1)
        i, f, g, o = torch.sigmoid(i), torch.sigmoid(f), torch.tanh(g), torch.sigmoid(o)
        c = f * c + i * g
        h = o * torch.tanh(c)
        G[t], Cc[t], Hh[t] = torch.cat([i, f, g, o], 1), c, h
    return G, Cc, Hh


def _lstm_bwd(G, Cc, Hh, X, w_ih, w_hh, dH_above, need_dx):
    """BPTT of one layer; G is overwritten with the pre-activation gate gradients.  Returns dW_ih, dW_hh, db, dX."""
    Tp, R, H4 = G.shape
    H = H4 // 4
    dh_rec, dc = torch.zeros(R, H), torch.zeros(R, H)
    dX = torch.zeros_like(X) if need_dx else None
    for t in range(Tp - 1, -1, -1):
        i, f, g, o = G[t].split(H, dim=1)
        dh = dH_above[t] + dh_rec
        tc = torch.tanh(Cc[t])
        dc_tot = dc + dh * o * (1 - tc * tc)
        c_prev = Cc[t - 1] if t > 0 else torch.zeros(R, H)
        dG = torch.cat([dc_tot * g * i * (1 - i), dc_tot * c_prev * f * (1 - f), dc_tot * i * (1 - g * g),
                        dh * tc * o * (1 - o)], 1)
        dc = dc_tot * f
        G[t] = dG
        dh_rec = dG @ w_hh
        if need_dx:
            dX[t] = dG @ w_ih
    flat = G.reshape(Tp * R, H4)
    dW_ih = flat.T @ X.reshape(Tp * R, -1)
    dW_hh = G[1:].reshape(-1, H4).T @ Hh[:-1].reshape(-1, H)
    return dW_ih, dW_hh, flat.sum(0), dX


def manual_backward(noisy_mag, cirm, sd, args=None):
    """Forward + backward exactly as libfsn_b200's training path does it.  noisy_mag [B,F,T], cirm [B',F',T,2]."""
    a = dict(O.DEFAULT_MODEL_ARGS)
    a.update(args or {})
    la, Ns, Nf, G = a["look_ahead"], a["sb_num_neighbors"], a["fb_num_neighbors"], a["num_groups_in_drop_band"]
    assert Nf == 0
    B, F, T = noisy_mag.shape
    Tp = T + la
    magT = torch.nn.functional.pad(noisy_mag, [0, la]).permute(2, 0, 1).contiguous()  # [Tp,B,F]
    inv1 = 1.0 / (magT.mean(dim=(0, 2)) + 1e-5)  # [B]
    Xfb = magT * inv1[None, :, None]
    fb = {k[len("fb_model."):]: v for k, v in sd.items() if k.startswith("fb_model.")}
    sb = {k[len("sb_model."):]: v for k, v in sd.items() if k.startswith("sb_model.")}

    def lw(d, l):
        return [d[f"sequence_model.{n}_l{l}"] for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]

    fG0, fC0, fH0 = _lstm_fwd_save(Xfb, *lw(fb, 0))
    fG1, fC1, fH1 = _lstm_fwd_save(fH0, *lw(fb, 1))
    z = fH1 @ fb["fc_output_layer.weight"].T + fb["fc_output_layer.bias"]
    relu = a["fb_output_activate_function"] == "ReLU"
    fbz = torch.relu(z) if relu else z  # [Tp,B,F]
    cnt = torch.from_numpy(O.reflect_count(F, Ns)).float()
    Ksb = 2 * Ns + 2
    cnt2 = float(F * Ksb * Tp)
    mu2 = ((magT * cnt[None, None, :]).sum(dim=(0, 2)) + fbz.sum(dim=(0, 2))) / cnt2
    inv2 = 1.0 / (mu2 + 1e-5)
    # row map (drop_band, feature.py:332-345); B == 1 -> identity (model.py:114)
    if B > 1 and G > 1:
        sb_, sf_ = O.drop_band_index_map(B, F, G)  # clip of output clip b', frequencies of (b', f')
        bsel, fsel = np.repeat(sb_, sf_.shape[1]), sf_.reshape(-1)
    else:
        bsel, fsel = np.repeat(np.arange(B), F), np.tile(np.arange(F), B)
    R = len(bsel)
    Fsub = R // B
    idx = np.abs(fsel[:, None] + np.arange(-Ns, Ns + 1)[None, :])
    idx = np.where(idx > F - 1, 2 * (F - 1) - idx, idx)  # reflect, no edge repeat
    bs, fs = torch.from_numpy(bsel), torch.from_numpy(fsel)
    Xsb = torch.cat([magT[:, bs[:, None], torch.from_numpy(idx)], fbz[:, bs, fs][..., None]], -1) * inv2[bs][None, :, None]
    sG0, sC0, sH0 = _lstm_fwd_save(Xsb, *lw(sb, 0))
    sG1, sC1, sH1 = _lstm_fwd_save(sH0, *lw(sb, 1))
    out = sH1 @ sb["fc_output_layer.weight"].T + sb["fc_output_layer.bias"]  # [Tp,R,2]
    crm = out[la:].reshape(T, B, Fsub, 2).permute(1, 2, 0, 3)  # [B',F',T,2]
    diff = crm - cirm
    loss = (diff ** 2).mean()
    # ---- backward
    dout = torch.zeros(Tp, R, 2)
    dout[la:] = (2.0 / diff.numel() * diff).permute(2, 0, 1, 3).reshape(T, R, 2)
    grads = {}
    grads["sb_model.fc_output_layer.weight"] = dout.reshape(-1, 2).T @ sH1.reshape(Tp * R, -1)
    grads["sb_model.fc_output_layer.bias"] = dout.reshape(-1, 2).sum(0)
    dH1 = dout @ sb["fc_output_layer.weight"]
    w = lw(sb, 1)
    dWi, dWh, db, dH0 = _lstm_bwd(sG1, sC1, sH1, sH0, w[0], w[1], dH1, True)
    grads.update({"sb_model.sequence_model.weight_ih_l1": dWi, "sb_model.sequence_model.weight_hh_l1": dWh,
                  "sb_model.sequence_model.bias_ih_l1": db, "sb_model.sequence_model.bias_hh_l1": db})
    w = lw(sb, 0)
    dWi, dWh, db, dX = _lstm_bwd(sG0, sC0, sH0, Xsb, w[0], w[1], dH0, True)
    grads.update({"sb_model.sequence_model.weight_ih_l0": dWi, "sb_model.sequence_model.weight_hh_l0": dWh,
                  "sb_model.sequence_model.bias_ih_l0": db, "sb_model.sequence_model.bias_hh_l0": db})
    # second norm: X = raw * inv2[b]  ->  d raw = dX * inv2,  d mu2[b] = -inv2[b] * sum(dX * X)
    dot = torch.zeros(B).index_add_(0, bs, (dX * Xsb).sum(dim=(0, 2)))
    dmu2 = -inv2 * dot
    dfbz = torch.zeros(Tp, B, F)
    dfbz[:, bs, fs] = dX[:, :, Ksb - 1] * inv2[bs][None, :]
    dfbz += (dmu2 / cnt2)[None, :, None]
    dz = dfbz * (fbz > 0) if relu else dfbz
    grads["fb_model.fc_output_layer.weight"] = dz.reshape(Tp * B, F).T @ fH1.reshape(Tp * B, -1)
    grads["fb_model.fc_output_layer.bias"] = dz.reshape(Tp * B, F).sum(0)
    dfH1 = dz @ fb["fc_output_layer.weight"]
    w = lw(fb, 1)
    dWi, dWh, db, dfH0 = _lstm_bwd(fG1, fC1, fH1, fH0, w[0], w[1], dfH1, True)
    grads.update({"fb_model.sequence_model.weight_ih_l1": dWi, "fb_model.sequence_model.weight_hh_l1": dWh,
                  "fb_model.sequence_model.bias_ih_l1": db, "fb_model.sequence_model.bias_hh_l1": db})
    w = lw(fb, 0)
    dWi, dWh, db, _ = _lstm_bwd(fG0, fC0, fH0, Xfb, w[0], w[1], dfH0, False)
    grads.update({"fb_model.sequence_model.weight_ih_l0": dWi, "fb_model.sequence_model.weight_hh_l0": dWh,
                  "fb_model.sequence_model.bias_ih_l0": db, "fb_model.sequence_model.bias_hh_l0": db})
    return loss, grads, crm

"""TEST INFRASTRUCTURE ONLY - not shipped, not imported by ``fullsubnet_b200``.

CPU fp32 restatement of the FullSubNet enhancement hot path (SURVEY.md section 8a
rows A1-A12).  Every function cites the reference file:line it follows
(paths relative to the upstream repository root).  The arithmetic is written
out with elementary torch CPU tensor ops (index, matmul, sigmoid, tanh,
rfft/irfft) - no ``torch.stft``, ``torch.istft``, ``nn.LSTM``, ``F.unfold`` -
so that it is an independent statement of what those library calls compute.

Parity pin: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4).  This restatement is pinned against the reference ITSELF,
executed in the build container: ``oracle/make_golden.py`` imports the
unmodified upstream modules from ``/root/reference`` and writes
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function
here against those fixtures.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

# audio_zen/constant.py:9
EPSILON = float(np.finfo(np.float32).eps)


# --------------------------------------------------------------------------- #
# A1 / A9: STFT and iSTFT                                                      #
# --------------------------------------------------------------------------- #
def hann_window(win_length: int, n_fft: Optional[int] = None) -> torch.Tensor:
    """Periodic hann (torch.hann_window default, audio_zen/acoustics/feature.py:38),
    centre-padded with zeros to ``n_fft`` as torch.stft does when win_length < n_fft."""
    n_fft = n_fft or win_length
    n = torch.arange(win_length, dtype=torch.float64)
    w = 0.5 - 0.5 * torch.cos(2.0 * math.pi * n / win_length)
    out = torch.zeros(n_fft, dtype=torch.float64)
    left = (n_fft - win_length) // 2
    out[left:left + win_length] = w
    return out.to(torch.float32)


def _reflect_index(i: torch.Tensor, n: int) -> torch.Tensor:
    """Index map of 'reflect' padding (no edge repeat): -k -> k, n-1+k -> n-1-k."""
    i = i.abs()
    return torch.where(i >= n, 2 * (n - 1) - i, i)


def stft(y: torch.Tensor, n_fft: int, hop_length: int, win_length: int):
    """audio_zen/acoustics/feature.py:9-50 (-> torch.stft, center=True, reflect pad,
    one-sided, un-normalised).  Returns (mag, phase, real, imag), each [B,F,T] or
    [B,C,F,T] for 3-D input."""
    assert y.dim() in (2, 3), "Only support 2D or 3D Input"  # feature.py:25
    batch = y.shape[0]
    L = y.shape[-1]
    three_d = y.dim() == 3
    if three_d:
        y = y.reshape(-1, L)  # feature.py:30-31
    pad = n_fft // 2
    T = 1 + L // hop_length
    idx = torch.arange(T)[:, None] * hop_length + torch.arange(n_fft)[None, :] - pad
    idx = _reflect_index(idx, L)
    frames = y[:, idx] * hann_window(win_length, n_fft)  # [B,T,n_fft]
    spec = torch.fft.rfft(frames, dim=-1).transpose(1, 2)  # [B,F,T]
    if three_d:
        spec = spec.reshape(batch, -1, spec.shape[-2], spec.shape[-1])  # feature.py:43-44
    return torch.abs(spec), torch.angle(spec), spec.real.contiguous(), spec.imag.contiguous()


def istft(features, n_fft: int, hop_length: int, win_length: int,
          length: Optional[int] = None, input_type: str = "complex") -> torch.Tensor:
    """audio_zen/acoustics/feature.py:53-91 (-> torch.istft, center=True)."""
    if input_type == "real_imag":
        assert isinstance(features, (tuple, list))  # feature.py:69
        real, imag = features
    elif input_type == "complex":
        assert torch.is_complex(features), "The input feature is not complex."  # feature.py:73
        real, imag = features.real, features.imag
    elif input_type == "mag_phase":
        assert isinstance(features, (tuple, list))
        mag, phase = features
        real, imag = mag * torch.cos(phase), mag * torch.sin(phase)  # feature.py:78
    else:
        raise NotImplementedError("Only 'real_imag', 'complex', and 'mag_phase' are supported.")
    spec = torch.complex(real.float(), imag.float())  # [B,F,T]
    B, F, T = spec.shape
    w = hann_window(win_length, n_fft)
    frames = torch.fft.irfft(spec.transpose(1, 2), n=n_fft, dim=-1) * w  # [B,T,n_fft]
    full = n_fft + hop_length * (T - 1)
    y = torch.zeros(B, full, dtype=torch.float32)
    env = torch.zeros(full, dtype=torch.float32)
    w2 = w * w
    for t in range(T):  # overlap-add (torch.istft uses col2im/fold)
        y[:, t * hop_length:t * hop_length + n_fft] += frames[:, t]
        env[t * hop_length:t * hop_length + n_fft] += w2
    start = n_fft // 2
    end = full - n_fft // 2 if length is None else start + length
    y, env = y[:, start:end], env[start:end]
    assert float(env.abs().min()) > 1e-11, "window overlap add min"
    y = y / env
    if length is not None and y.shape[-1] < length:
        y = torch.nn.functional.pad(y, (0, length - y.shape[-1]))
    return y


def mag_phase(complex_tensor):
    """audio_zen/acoustics/feature.py:94-96"""
    return torch.abs(complex_tensor), torch.angle(complex_tensor)


# --------------------------------------------------------------------------- #
# A7: drop_band                                                                #
# --------------------------------------------------------------------------- #
def drop_band(x: torch.Tensor, num_groups: int = 2) -> torch.Tensor:
    """audio_zen/acoustics/feature.py:309-345.  [B,C,F,T] -> [B,C,F//G,T]; group g keeps
    clips g::G and frequencies g::G of the first F-(F%G) bins; groups concatenated on
    the batch axis (batch order 0,2,4,...,1,3,5,... for G=2)."""
    B, _, F, _ = x.shape
    assert B > num_groups, (
        f"Batch size = {B}, num_groups = {num_groups}. The batch size should larger than the num_groups.")
    if num_groups <= 1:
        return x
    if F % num_groups != 0:
        x = x[..., : F - (F % num_groups), :]
        F = x.shape[2]
    out = []
    for g in range(num_groups):
        out.append(x[g::num_groups][:, :, g::num_groups, :])
    return torch.cat(out, dim=0)


def drop_band_index_map(B: int, F: int, G: int) -> Tuple[np.ndarray, np.ndarray]:
    """Index form of drop_band: output row (b', f') comes from input (src_b[b'], src_f[b', f'])."""
    Fk = F - (F % G)
    src_b, src_f = [], []
    for g in range(G):
        for b in range(g, B, G):
            src_b.append(b)
            src_f.append(np.arange(g, Fk, G))
    return np.asarray(src_b), np.stack(src_f)


# --------------------------------------------------------------------------- #
# A9 / A10: masks                                                              #
# --------------------------------------------------------------------------- #
def compress_cIRM(mask: torch.Tensor, K: float = 10, C: float = 0.1) -> torch.Tensor:
    """audio_zen/acoustics/mask.py:32-44"""
    mask = -100 * (mask <= -100) + mask * (mask > -100)
    return K * (1 - torch.exp(-C * mask)) / (1 + torch.exp(-C * mask))


def build_complex_ideal_ratio_mask(nr, ni, cr, ci) -> torch.Tensor:
    """audio_zen/acoustics/mask.py:7-29 -> [B,F,T,2] (compressed)."""
    den = nr * nr + ni * ni + EPSILON
    mr = (nr * cr + ni * ci) / den
    mi = (nr * ci - ni * cr) / den
    return compress_cIRM(torch.stack((mr, mi), dim=-1), K=10, C=0.1)


def decompress_cIRM(mask: torch.Tensor, K: float = 10, limit: float = 9.9) -> torch.Tensor:
    """audio_zen/acoustics/mask.py:47-64"""
    mask = limit * (mask >= limit) - limit * (mask <= -limit) + mask * (torch.abs(mask) < limit)
    return -K * torch.log((K - mask) / (K + mask))


# --------------------------------------------------------------------------- #
# A3 / A5: norm and sub-band unfold                                            #
# --------------------------------------------------------------------------- #
def offline_laplace_norm(x: torch.Tensor) -> torch.Tensor:
    """audio_zen/model/base_model.py:203-218: per-clip mean over every non-batch axis."""
    mu = x.mean(dim=list(range(1, x.dim())), keepdim=True)
    return x / (mu + 1e-5)


def cumulative_laplace_norm(x: torch.Tensor) -> torch.Tensor:
    """audio_zen/model/base_model.py:220-251."""
    B, C, F, T = x.shape
    x = x.reshape(B * C, F, T)
    cum = torch.cumsum(x.sum(dim=1), dim=-1)
    cnt = torch.arange(F, F * T + 1, F, dtype=x.dtype).reshape(1, T)
    mean = (cum / cnt).reshape(B * C, 1, T)
    return (x / (mean + EPSILON)).reshape(B, C, F, T)


def freq_unfold(x: torch.Tensor, num_neighbors: int) -> torch.Tensor:
    """audio_zen/model/base_model.py:13-46.  [B,C,F,T] -> [B,F,C,2N+1,T]; unit f holds
    rows reflect(f-N .. f+N) of the input."""
    assert x.dim() == 4, f"The dim of the input is {x.dim()}. It should be four dim."
    B, C, F, T = x.shape
    if num_neighbors <= 0:
        return x.permute(0, 2, 1, 3).reshape(B, F, C, 1, T)
    N = num_neighbors
    rows = torch.arange(F)[:, None] + torch.arange(-N, N + 1)[None, :]  # [F, 2N+1]
    rows = _reflect_index(rows, F)
    out = x[:, :, rows, :]  # [B,C,F,2N+1,T]
    return out.permute(0, 2, 1, 3, 4).contiguous()


def reflect_count(F: int, N: int) -> np.ndarray:
    """c[r] = #{(f,k): reflect(f+k)=r, |k|<=N}: how many times row r of the magnitude
    appears in the unfolded sub-band input (SURVEY 8a row A6 closed form)."""
    c = np.zeros(F, dtype=np.int64)
    for f in range(F):
        for k in range(-N, N + 1):
            r = abs(f + k)
            if r >= F:
                r = 2 * (F - 1) - r
            c[r] += 1
    return c


# --------------------------------------------------------------------------- #
# A4 / A8: SequenceModel (stacked LSTM + Linear + activation)                  #
# --------------------------------------------------------------------------- #
def lstm_stack(x: torch.Tensor, sd: Dict[str, torch.Tensor], prefix: str, num_layers: int = 2,
               operand_round=None) -> torch.Tensor:
    """nn.LSTM(batch_first=True) restated (audio_zen/model/module/sequence_model.py:52-58,117):
    gates (i,f,g,o) = W_ih x_t + b_ih + W_hh h_{t-1} + b_hh; c_t = s(f) c + s(i) tanh(g);
    h_t = s(o) tanh(c_t); zero initial state.  x: [B,T,In] -> [B,T,H].

    ``operand_round`` (test-only) rounds matmul operands (e.g. to fp16) to model the
    tensor-core operand precision of the CUDA fast path."""
    rnd = operand_round or (lambda t: t)
    B, T, _ = x.shape
    inp = x
    for layer in range(num_layers):
        w_ih = rnd(sd[f"{prefix}weight_ih_l{layer}"]).t().contiguous()
        w_hh = rnd(sd[f"{prefix}weight_hh_l{layer}"]).t().contiguous()
        bias = sd[f"{prefix}bias_ih_l{layer}"] + sd[f"{prefix}bias_hh_l{layer}"]
        H = w_hh.shape[0]
        h = torch.zeros(B, H)
        c = torch.zeros(B, H)
        outs = []
        # input projection for every step at once (it does not depend on the recurrence), as ATen's
        # CPU LSTM does; the recurrent half stays a serial loop over t
        xproj = torch.addmm(bias, rnd(inp).reshape(B * T, -1), w_ih).reshape(B, T, 4 * H)
        for t in range(T):
            g = torch.addmm(xproj[:, t], rnd(h), w_hh)
            i, f, gg, o = g.split(H, dim=1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
            h = torch.sigmoid(o) * torch.tanh(c)
            outs.append(h)
        inp = torch.stack(outs, dim=1)
    return inp


def gru_stack(x: torch.Tensor, sd: Dict[str, torch.Tensor], prefix: str, num_layers: int = 2) -> torch.Tensor:
    """nn.GRU(batch_first=True) restated (audio_zen/model/module/sequence_model.py:59-66,117): weights [3H,K] with gate
    order (r,z,n); r = s(W_ir x + b_ir + W_hr h + b_hr), z likewise, n = tanh(W_in x + b_in + r * (W_hn h + b_hn)),
    h' = (1 - z) n + z h; zero initial state.  x: [B,T,In] -> [B,T,H]."""
    B, T, _ = x.shape
    inp = x
    for layer in range(num_layers):
        w_ih = sd[f"{prefix}weight_ih_l{layer}"].t().contiguous()
        w_hh = sd[f"{prefix}weight_hh_l{layer}"].t().contiguous()
        b_ih, b_hh = sd[f"{prefix}bias_ih_l{layer}"], sd[f"{prefix}bias_hh_l{layer}"]
        H = w_hh.shape[0]
        h = torch.zeros(B, H)
        outs = []
        xproj = torch.addmm(b_ih, inp.reshape(B * T, -1), w_ih).reshape(B, T, 3 * H)
        for t in range(T):
            hp = torch.addmm(b_hh, h, w_hh)
            xr, xz, xn = xproj[:, t].split(H, dim=1)
            hr, hz, hn = hp.split(H, dim=1)
            r = torch.sigmoid(xr + hr)
            z = torch.sigmoid(xz + hz)
            n = torch.tanh(xn + r * hn)
            h = (1.0 - z) * n + z * h
            outs.append(h)
        inp = torch.stack(outs, dim=1)
    return inp


def sequence_model(x: torch.Tensor, sd: Dict[str, torch.Tensor], prefix: str,
                   activation: Optional[str], operand_round=None) -> torch.Tensor:
    """audio_zen/model/module/sequence_model.py:106-125.  x: [B,F,T] -> [B,F_out,T].  The cell (LSTM / GRU) is read off
    the weight shapes: [4H,K] vs [3H,K]."""
    assert x.dim() == 3, f"The shape of input is {x.shape}."
    pre = prefix + "sequence_model."
    if sd[pre + "weight_hh_l0"].shape[0] == 3 * sd[pre + "weight_hh_l0"].shape[1]:
        o = gru_stack(x.permute(0, 2, 1), sd, pre)
    else:
        o = lstm_stack(x.permute(0, 2, 1), sd, pre, operand_round=operand_round)
    o = o @ sd[prefix + "fc_output_layer.weight"].t() + sd[prefix + "fc_output_layer.bias"]
    if activation:
        if activation == "ReLU":
            o = torch.relu(o)
        elif activation == "Tanh":
            o = torch.tanh(o)
        elif activation == "ReLU6":
            o = torch.clamp(o, 0, 6)
        else:
            raise NotImplementedError(f"Not implemented activation function {activation}")
    return o.permute(0, 2, 1)


# --------------------------------------------------------------------------- #
# A2-A8, A12: fullsubnet Model.forward                                         #
# --------------------------------------------------------------------------- #
DEFAULT_MODEL_ARGS = dict(  # recipes/dns_interspeech_2020/fullsubnet/inference.toml:33-44
    num_freqs=257, look_ahead=2, sequence_model="LSTM", fb_num_neighbors=0, sb_num_neighbors=15,
    fb_output_activate_function="ReLU", sb_output_activate_function=False,
    fb_model_hidden_size=512, sb_model_hidden_size=384, norm_type="offline_laplace_norm",
    num_groups_in_drop_band=2, weight_init=False,
)


def model_forward(noisy_mag: torch.Tensor, sd: Dict[str, torch.Tensor], args: Optional[dict] = None,
                  operand_round=None, return_intermediates: bool = False):
    """recipes/dns_interspeech_2020/fullsubnet/model.py:72-136.  noisy_mag [B,1,F,T] ->
    cRM [B,2,F',T] (F' = F, or F//G with the drop_band batch permutation when B>1, G>1)."""
    a = dict(DEFAULT_MODEL_ARGS)
    a.update(args or {})
    norm = {"offline_laplace_norm": offline_laplace_norm,
            "cumulative_laplace_norm": cumulative_laplace_norm}.get(a["norm_type"])
    if norm is None:
        raise NotImplementedError("You must set up a type of Norm.")
    assert noisy_mag.dim() == 4  # model.py:84
    la, Nf, Ns = a["look_ahead"], a["fb_num_neighbors"], a["sb_num_neighbors"]
    noisy_mag = torch.nn.functional.pad(noisy_mag, [0, la])  # model.py:85
    B, C, F, T = noisy_mag.shape
    assert C == 1  # model.py:87-89
    fb_input = norm(noisy_mag).reshape(B, C * F, T)  # model.py:92
    fb_output = sequence_model(fb_input, sd, "fb_model.", a["fb_output_activate_function"]).reshape(B, 1, F, T)
    fb_unf = freq_unfold(fb_output, Nf).reshape(B, F, 2 * Nf + 1, T)  # model.py:98-101
    mag_unf = freq_unfold(noisy_mag, Ns).reshape(B, F, 2 * Ns + 1, T)  # model.py:104-107
    sb_input = norm(torch.cat([mag_unf, fb_unf], dim=2))  # model.py:110-111
    if B > 1:  # model.py:114-119 (eval mode too)
        sb_input = drop_band(sb_input.permute(0, 2, 1, 3), num_groups=a["num_groups_in_drop_band"])
        F = sb_input.shape[2]
        sb_input = sb_input.permute(0, 2, 1, 3)
    sb_input = sb_input.reshape(B * F, (2 * Ns + 1) + (2 * Nf + 1), T)  # model.py:121-125
    sb_mask = sequence_model(sb_input, sd, "sb_model.", a["sb_output_activate_function"],
                             operand_round=operand_round)
    sb_mask = sb_mask.reshape(B, F, 2, T).permute(0, 2, 1, 3).contiguous()  # model.py:129-133
    out = sb_mask[:, :, :, la:]  # model.py:135
    if return_intermediates:
        return out, dict(fb_output=fb_output, sb_input=sb_input)
    return out


def enhance(noisy: torch.Tensor, sd: Dict[str, torch.Tensor], args: Optional[dict] = None,
            n_fft: int = 512, hop_length: int = 256, win_length: int = 512,
            batched: bool = True, return_crm: bool = False):
    """recipes/dns_interspeech_2020/inferencer.py:130-145 (Inferencer.full_band_crm_mask)
    for ``noisy`` [B,L].  The reference inferencer is B=1 only
    (audio_zen/inferencer/base_inferencer.py:78,173); batched inference is defined as a loop
    of B=1 calls, i.e. ``num_groups_in_drop_band=1``."""
    a = dict(DEFAULT_MODEL_ARGS)
    a.update(args or {})
    a["num_groups_in_drop_band"] = 1
    outs, crms = [], []
    chunks = [noisy] if batched else [noisy[i:i + 1] for i in range(noisy.shape[0])]
    for y in chunks:
        mag, _, real, imag = stft(y, n_fft, hop_length, win_length)
        crm = model_forward(mag.unsqueeze(1), sd, a)
        m = decompress_cIRM(crm.permute(0, 2, 3, 1))
        er = m[..., 0] * real - m[..., 1] * imag  # inferencer.py:139-140
        ei = m[..., 1] * real + m[..., 0] * imag
        outs.append(istft((er, ei), n_fft, hop_length, win_length, length=y.shape[-1], input_type="real_imag"))
        crms.append(crm)
    wav = torch.cat(outs, 0)
    return (wav, torch.cat(crms, 0)) if return_crm else wav


# --------------------------------------------------------------------------- #
# A10 / A11: training-step pieces that do not need autograd                    #
# --------------------------------------------------------------------------- #
def train_targets_and_loss(noisy: torch.Tensor, clean: torch.Tensor, sd, args=None,
                           n_fft=512, hop_length=256, win_length=512):
    """recipes/dns_interspeech_2020/fullsubnet/trainer.py:46-61 without DDP/AMP:
    cIRM target (drop_band'ed), forward with drop_band, MSE loss."""
    a = dict(DEFAULT_MODEL_ARGS)
    a.update(args or {})
    nm, _, nr, ni = stft(noisy, n_fft, hop_length, win_length)
    _, _, cr, ci = stft(clean, n_fft, hop_length, win_length)
    cirm = build_complex_ideal_ratio_mask(nr, ni, cr, ci)
    cirm = drop_band(cirm.permute(0, 3, 1, 2), a["num_groups_in_drop_band"]).permute(0, 2, 3, 1)
    crm = model_forward(nm.unsqueeze(1), sd, a).permute(0, 2, 3, 1)
    loss = torch.mean((cirm - crm) ** 2)  # audio_zen/loss.py:4
    return cirm, crm, loss


# --------------------------------------------------------------------------- #
# Seeded synthetic weights / inputs                                #
# --------------------------------------------------------------------------- #
def state_dict_shapes(args: Optional[dict] = None) -> Sequence[Tuple[str, Tuple[int, ...]]]:
    """The 20 state_dict entries of fullsubnet/model.py:Model in registration order (A12)."""
    a = dict(DEFAULT_MODEL_ARGS)
    a.update(args or {})
    F, Hf, Hs = a["num_freqs"], a["fb_model_hidden_size"], a["sb_model_hidden_size"]
    sb_in = (2 * a["sb_num_neighbors"] + 1) + (2 * a["fb_num_neighbors"] + 1)
    out = []
    ng = 3 if a.get("sequence_model", "LSTM") == "GRU" else 4  # nn.GRU: [3H,K] (r,z,n); nn.LSTM: [4H,K] (i,f,g,o)
    for pre, In, H, Out in (("fb_model.", F, Hf, F), ("sb_model.", sb_in, Hs, 2)):
        for l in range(2):
            k = In if l == 0 else H
            out += [(f"{pre}sequence_model.weight_ih_l{l}", (ng * H, k)),
                    (f"{pre}sequence_model.weight_hh_l{l}", (ng * H, H)),
                    (f"{pre}sequence_model.bias_ih_l{l}", (ng * H,)),
                    (f"{pre}sequence_model.bias_hh_l{l}", (ng * H,))]
        out += [(f"{pre}fc_output_layer.weight", (Out, H)), (f"{pre}fc_output_layer.bias", (Out,))]
    return out


def make_state_dict(seed: int = 0, args: Optional[dict] = None, sb_fc_gain: float = 1.0
                    ) -> Dict[str, torch.Tensor]:
    """Weight set W-a of SURVEY 8d: uniform(-1/sqrt(H), 1/sqrt(H)) like PyTorch's default LSTM /
    Linear init, but drawn from numpy's PCG64 stream so the values do not depend on the torch
    version.  ``sb_fc_gain`` > 1 gives weight set W-b (cRM spanning the +-9.9 clip region)."""
    rng = np.random.default_rng(seed)
    sd = {}
    for name, shape in state_dict_shapes(args):
        fan = shape[-1] if "fc_output_layer.weight" in name else None
        if "sequence_model" in name:
            H = shape[0] // (3 if (args or {}).get("sequence_model", "LSTM") == "GRU" else 4)
            k = 1.0 / math.sqrt(H)
        else:
            H = fan if fan is not None else sd[name.replace("bias", "weight")].shape[1]
            k = 1.0 / math.sqrt(H)
        w = rng.uniform(-k, k, size=shape).astype(np.float32)
        if name.startswith("sb_model.fc_output_layer"):
            w = w * np.float32(sb_fc_gain)
        sd[name] = torch.from_numpy(w)
    return sd


def make_noisy(B: int, L: int, seed: int = 0, speechlike: bool = False, sr: int = 16000) -> torch.Tensor:
    """Synthetic clips of SURVEY 8d: 0.1*randn, optionally plus 5 harmonics of 120 Hz with 3 Hz vibrato."""
    rng = np.random.default_rng(1000 + seed)
    y = 0.1 * rng.standard_normal((B, L)).astype(np.float32)
    if speechlike:
        t = np.arange(L, dtype=np.float64) / sr
        f0 = 120.0 * (1 + 0.02 * np.sin(2 * np.pi * 3.0 * t))
        ph = 2 * np.pi * np.cumsum(f0) / sr
        s = sum(np.sin(h * ph) / h for h in range(1, 6))
        y = y + (0.1 * s / np.abs(s).max()).astype(np.float32)[None]
    return torch.from_numpy(y.astype(np.float32))

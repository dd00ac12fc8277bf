"""TEST / BENCH INFRASTRUCTURE ONLY - not shipped, not imported by ``fullsubnet_b200``.

The CPU arm of ``bench.py``: the FullSubNet enhancement path written with the SAME third-party PyTorch library
calls the reference makes (``torch.stft`` / ``torch.istft`` - audio_zen/acoustics/feature.py:33-40,84-91;
``nn.LSTM`` + ``nn.Linear`` - audio_zen/model/module/sequence_model.py:52-58,82-84,117-123; ``F.pad`` reflect +
``F.unfold`` - audio_zen/model/base_model.py:35-44), so that it runs at the reference's own CPU speed (ATen's fused
LSTM, MKL FFT).  ``oracle/fullsubnet_oracle.py`` (elementary ops, 1.7x slower) stays the PARITY checker; this file is
only the thing that gets TIMED on the host cores, and ``tests/test_oracle_golden.py`` pins it to the same goldens
(outputs of the unmodified reference) so that the timed code is known to compute the reference's result.
The reference itself is pure Python with uninstalled dependencies (librosa, soundfile) and no setup.py/pyproject,
so it can neither be pip-installed into ``baseline/_ref`` nor travel to the GPU box.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import fullsubnet_oracle as O


class LibcallModel(torch.nn.Module):
    """recipes/dns_interspeech_2020/fullsubnet/model.py:9-136 with torch's own layers (LSTM recipe only)."""

    def __init__(self, sd: Dict[str, torch.Tensor], args: Optional[dict] = None):
        super().__init__()
        a = dict(O.DEFAULT_MODEL_ARGS)
        a.update(args or {})
        self.a = a
        nf, Hf, Hs = a["num_freqs"], a["fb_model_hidden_size"], a["sb_model_hidden_size"]
        sb_in = (2 * a["sb_num_neighbors"] + 1) + (2 * a["fb_num_neighbors"] + 1)
        self.fb_lstm = torch.nn.LSTM(nf, Hf, num_layers=2, batch_first=True)   # sequence_model.py:52-58
        self.fb_fc = torch.nn.Linear(Hf, nf)                                   # sequence_model.py:82-84
        self.sb_lstm = torch.nn.LSTM(sb_in, Hs, num_layers=2, batch_first=True)
        self.sb_fc = torch.nn.Linear(Hs, 2)
        with torch.no_grad():
            for pre, lstm, fc in (("fb_model.", self.fb_lstm, self.fb_fc), ("sb_model.", self.sb_lstm, self.sb_fc)):
                for name, p in lstm.named_parameters():
                    p.copy_(sd[f"{pre}sequence_model.{name}"])
                fc.weight.copy_(sd[f"{pre}fc_output_layer.weight"])
                fc.bias.copy_(sd[f"{pre}fc_output_layer.bias"])
        self.eval()

    @staticmethod
    def _norm(x):  # base_model.py:203-218
        mu = torch.mean(x, dim=list(range(1, x.dim())), keepdim=True)
        return x / (mu + 1e-5)

    @staticmethod
    def _unfold(x, N):  # base_model.py:13-46
        B, C, Fq, T = x.shape
        if N <= 0:
            return x.permute(0, 2, 1, 3).reshape(B, Fq, C, 1, T)
        xp = F.pad(x, [0, 0, N, N], mode="reflect")
        out = F.unfold(xp, (2 * N + 1, T))
        return out.reshape(B, C, 2 * N + 1, T, Fq).permute(0, 4, 1, 2, 3).contiguous()

    def forward(self, noisy_mag):  # model.py:72-136, B = 1 (the reference inferencer's only batch) or G = 1
        a = self.a
        la, Nf, Ns = a["look_ahead"], a["fb_num_neighbors"], a["sb_num_neighbors"]
        x = F.pad(noisy_mag, [0, la])
        B, C, Fq, T = x.shape
        fb_in = self._norm(x).reshape(B, C * Fq, T)
        fb = torch.relu(self.fb_fc(self.fb_lstm(fb_in.permute(0, 2, 1))[0])).permute(0, 2, 1).reshape(B, 1, Fq, T)
        fb_u = self._unfold(fb, Nf).reshape(B, Fq, 2 * Nf + 1, T)
        mag_u = self._unfold(x, Ns).reshape(B, Fq, 2 * Ns + 1, T)
        sb_in = self._norm(torch.cat([mag_u, fb_u], dim=2)).reshape(B * Fq, (2 * Ns + 1) + (2 * Nf + 1), T)
        sb = self.sb_fc(self.sb_lstm(sb_in.permute(0, 2, 1))[0]).permute(0, 2, 1)
        sb = sb.reshape(B, Fq, 2, T).permute(0, 2, 1, 3).contiguous()
        return sb[:, :, :, la:]


def enhance(noisy: torch.Tensor, model: LibcallModel, n_fft=512, hop=256, win=512, return_crm=False):
    """recipes/dns_interspeech_2020/inferencer.py:130-145, looped over clips with batch 1
    (audio_zen/inferencer/base_inferencer.py:78,173)."""
    outs, crms = [], []
    w = torch.hann_window(win)
    with torch.no_grad():
        for i in range(noisy.shape[0]):
            y = noisy[i:i + 1]
            spec = torch.stft(y, n_fft, hop_length=hop, win_length=win, window=w, return_complex=True)
            crm = model(spec.abs().unsqueeze(1))
            m = O.decompress_cIRM(crm.permute(0, 2, 3, 1))
            er = m[..., 0] * spec.real - m[..., 1] * spec.imag
            ei = m[..., 1] * spec.real + m[..., 0] * spec.imag
            outs.append(torch.istft(torch.complex(er, ei), n_fft, hop_length=hop, win_length=win, window=w,
                                    length=y.shape[-1]))
            crms.append(crm)
    wav = torch.cat(outs, 0)
    return (wav, torch.cat(crms, 0)) if return_crm else wav

"""Drop-in for recipes/dns_interspeech_2020/improved_fullsubnet/model.py:452-591 (class Model, BASELINE config 5).

Same constructor kwargs and ``state_dict`` keys (``fb_model.*``, ``sb_model.sb_models.{s}.*``);
``forward(y [B,L] | [B,1,L]) -> [B,1,L]`` (waveform in, enhanced waveform out) is one call into libfsn_b200
(``fsn_improved_forward``: STFT -> |X|^fdrc -> full band -> per-section sub bands -> element-wise mask -> iSTFT)."""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.nn as nn

from .. import _lib
from ..model.base_model import BaseModel
from ..model.module.sequence_model import SequenceModel


class SubbandModel(nn.Module):
    """Parameter container with the reference's layout (model.py:250-318): one 2-layer stack per section."""

    def __init__(self, freq_cutoffs, sb_num_center_freqs, sb_num_neighbor_freqs, fb_num_center_freqs,
                 fb_num_neighbor_freqs, hidden_size, sequence_model, activate_function):
        super().__init__()
        assert len(freq_cutoffs) + 1 == len(sb_num_center_freqs) == len(sb_num_neighbor_freqs) \
            == len(fb_num_center_freqs) == len(fb_num_neighbor_freqs)
        self.sb_models = nn.ModuleList([
            SequenceModel(input_size=(sb_num_center_freqs[s] + sb_num_neighbor_freqs[s] * 2)
                          + (fb_num_center_freqs[s] + fb_num_neighbor_freqs[s] * 2),
                          output_size=sb_num_center_freqs[s] * 2, hidden_size=hidden_size, num_layers=2,
                          bidirectional=False, sequence_model=sequence_model, output_activate_function=activate_function)
            for s in range(len(sb_num_center_freqs))])
        self.freq_cutoffs = list(freq_cutoffs)
        self.sb_num_center_freqs = list(sb_num_center_freqs)
        self.sb_num_neighbor_freqs = list(sb_num_neighbor_freqs)
        self.fb_num_center_freqs = list(fb_num_center_freqs)
        self.fb_num_neighbor_freqs = list(fb_num_neighbor_freqs)


class Model(BaseModel):
    def __init__(self, n_fft=512, hop_length=128, win_length=512, fdrc=0.5, num_freqs=257, freq_cutoffs=[20, 80],
                 sb_num_center_freqs=[1, 4, 8], sb_num_neighbor_freqs=[15, 15, 15], fb_num_center_freqs=[1, 4, 8],
                 fb_num_neighbor_freqs=[15, 15, 15], fb_hidden_size=512, sb_hidden_size=384, sequence_model="LSTM",
                 fb_output_activate_function=False, sb_output_activate_function=False,
                 norm_type="offline_laplace_norm"):
        super().__init__()
        self.n_fft, self.hop_length, self.win_length, self.fdrc = n_fft, hop_length, win_length, fdrc
        self.num_freqs = num_freqs
        self.fb_model = SequenceModel(input_size=num_freqs - 1, output_size=num_freqs - 1, hidden_size=fb_hidden_size,
                                      num_layers=2, bidirectional=False, sequence_model=sequence_model,
                                      output_activate_function=fb_output_activate_function)
        self.sb_model = SubbandModel(freq_cutoffs, sb_num_center_freqs, sb_num_neighbor_freqs, fb_num_center_freqs,
                                     fb_num_neighbor_freqs, sb_hidden_size, sequence_model, sb_output_activate_function)
        if len(sb_num_center_freqs) > _lib.IMP_MAX_SECTIONS:
            raise NotImplementedError(f"libfsn_b200 builds at most {_lib.IMP_MAX_SECTIONS} sub-band sections")
        if norm_type != "offline_laplace_norm":
            # model.py:226-236 also offers cumulative_laplace_norm / offline_gaussian_norm
            raise NotImplementedError("libfsn_b200 builds offline_laplace_norm for improved_fullsubnet")
        self.norm_type = norm_type
        # arithmetic of the sub-band sections (98 % of the FLOPs): "fp32" (FMA kernels), "tf32_tc" (tcgen05 kind::tf32
        # GEMMs, fp32 accumulate; waveform within 1e-4 of the reference) or "auto" (= tf32_tc when sb_hidden % 4 == 0)
        self.precision = os.environ.get("FSN_IMPROVED_PRECISION", "auto")

    def _resolve_precision(self) -> str:
        ok = self.sb_model.sb_models[0].hidden_size % 4 == 0
        if self.precision == "auto":
            return "tf32_tc" if ok else "fp32"
        if self.precision not in ("fp32", "tf32_tc"):
            raise ValueError("precision must be 'fp32', 'tf32_tc' or 'auto'")
        return self.precision

    def _structs(self):
        sb = self.sb_model
        d = _lib.ImprovedDesc(n_fft=self.n_fft, hop_length=self.hop_length, win_length=self.win_length,
                              num_freqs=self.num_freqs, fdrc=float(self.fdrc), num_sections=len(sb.sb_models),
                              fb_hidden=self.fb_model.hidden_size, sb_hidden=sb.sb_models[0].hidden_size,
                              fb_activation=_lib.ACT[self.fb_model.output_activate_function],
                              sb_activation=_lib.ACT[sb.sb_models[0].output_activate_function],
                              precision=_lib.PREC[self._resolve_precision()])
        for s in range(len(sb.sb_models)):
            if s < len(sb.freq_cutoffs):
                d.freq_cutoffs[s] = sb.freq_cutoffs[s]
            d.sb_num_center[s], d.sb_num_neighbor[s] = sb.sb_num_center_freqs[s], sb.sb_num_neighbor_freqs[s]
            d.fb_num_center[s], d.fb_num_neighbor[s] = sb.fb_num_center_freqs[s], sb.fb_num_neighbor_freqs[s]
        w = _lib.ImprovedWeights()
        w.fb = self.fb_model.weight_struct()
        for s, m in enumerate(sb.sb_models):
            w.sb[s] = m.weight_struct()
        return d, w

    def forward(self, y, return_crm: bool = False):
        """y [B,L] or [B,1,L] -> enhanced [B,1,L]  (model.py:541-591).  ``return_crm`` additionally returns the
        [B,2,F,T] mask (Nyquist row zero) - an extension used by the parity tests."""
        ndim = y.dim()
        assert ndim in (2, 3), "Input must be 2D (B, T) or 3D tensor (B, 1, T)"
        if ndim == 3:
            assert y.size(1) == 1, "Input must be 2D (B, T) or 3D tensor (B, 1, T)"
            y = y.squeeze(1)
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("fullsubnet_b200: backward kernels are not built yet; use torch.no_grad()/eval().")
        x = _lib.require_cuda(y, "y")
        B, L = x.shape
        sb = self.sb_model
        bounds = [0] + sb.freq_cutoffs + [self.num_freqs - 1]
        for s in range(len(sb.sb_models)):  # model.py:341-345 (both the noisy and the full-band unfold)
            if (bounds[s + 1] - bounds[s]) % sb.sb_num_center_freqs[s] or \
                    (bounds[s + 1] - bounds[s]) % sb.fb_num_center_freqs[s]:
                raise ValueError(
                    "The number of center frequencies should be divisible by the subband freqency interval. "
                    f"Got {sb.sb_num_center_freqs[s]} and {bounds[s + 1] - bounds[s]}.")
        lib = _lib.load()
        with torch.cuda.device(x.device):
            d, w = self._structs()
            n = lib.fsn_improved_workspace_bytes(C.byref(d), B, L)
            if n == 0:
                _lib.check_workspace(n)
            ws = torch.empty(n, dtype=torch.uint8, device=x.device)
            out = torch.empty(B, 1, L, dtype=torch.float32, device=x.device)
            crm = torch.empty(B, 2, self.num_freqs, 1 + L // self.hop_length, dtype=torch.float32,
                              device=x.device) if return_crm else None
            _lib.check(lib.fsn_improved_forward(C.byref(d), C.byref(w), x.data_ptr(), B, L, out.data_ptr(),
                                                _lib.ptr(crm), ws.data_ptr(), n, _lib.stream_ptr(x.device)))
        return (out, crm) if return_crm else out

"""Drop-in for recipes/dns_interspeech_2020/fullband_baseline/model.py:8-68 (class Model).

Same constructor kwargs and ``state_dict`` keys (``fullband_model.sequence_model.weight_ih_l{0,1,2}`` ...,
``fullband_model.fc_output_layer.*``); ``forward(noisy_mag [B,1,F,T]) -> [B,2,F,T]`` is one call into libfsn_b200
(``fsn_fullband_forward``)."""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib
from ..model.base_model import BaseModel
from ..model.module.sequence_model import SequenceModel


class Model(BaseModel):
    def __init__(self, num_freqs, hidden_size, sequence_model, output_activate_function, look_ahead,
                 norm_type="offline_laplace_norm", weight_init=True):
        super().__init__()
        self.fullband_model = SequenceModel(input_size=num_freqs, output_size=num_freqs * 2, hidden_size=hidden_size,
                                            num_layers=3, bidirectional=False, sequence_model=sequence_model,
                                            output_activate_function=output_activate_function)
        self.num_freqs = num_freqs
        self.look_ahead = look_ahead
        self.norm = self.norm_wrapper(norm_type)
        if weight_init:
            self.apply(self.weight_init)

    def forward(self, noisy_mag):
        """noisy_mag [B,1,F,T] -> [B,2,F,T]  (fullband_baseline/model.py:46-68)."""
        assert noisy_mag.dim() == 4
        batch_size, num_channels, num_freqs, num_frames = noisy_mag.size()
        assert num_channels == 1, f"{self.__class__.__name__} takes the mag feature as inputs."
        assert num_freqs == self.num_freqs, f"num_freqs {num_freqs} != {self.num_freqs}"
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("fullsubnet_b200: fullband_baseline is inference-only; use torch.no_grad().")
        x = _lib.require_cuda(noisy_mag, "noisy_mag")
        seq = self.fullband_model
        lib = _lib.load()
        with torch.cuda.device(x.device):
            d = _lib.FullbandDesc(num_freqs=num_freqs, hidden=seq.hidden_size, num_layers=seq.num_layers,
                                  look_ahead=self.look_ahead, activation=_lib.ACT[seq.output_activate_function],
                                  norm_type=self.norm)
            layers = (_lib.LstmLayer * seq.num_layers)(*(seq.layer_struct(l) for l in range(seq.num_layers)))
            fc_w, fc_b = seq.fc_ptrs()
            n = _lib.check_workspace(lib.fsn_fullband_workspace_bytes(C.byref(d), batch_size, num_frames))
            ws = torch.empty(n, dtype=torch.uint8, device=x.device)
            out = torch.empty(batch_size, 2, num_freqs, num_frames, dtype=torch.float32, device=x.device)
            _lib.check(lib.fsn_fullband_forward(C.byref(d), layers, fc_w, fc_b, x.data_ptr(), batch_size, num_frames,
                                                out.data_ptr(), ws.data_ptr(), n, _lib.stream_ptr(x.device)))
        return out

"""Drop-in for recipes/dns_interspeech_2020/fast_fullsubnet/model.py:11-202 (class Model, BASELINE config 4).

Same constructor kwargs and the same 31 ``state_dict`` entries (incl. the ``mel_scale.fb`` buffer that
torchaudio's MelScale registers); ``forward(mix_mag [B,1,F,T]) -> [B,2,F,T]`` is one call into libfsn_b200
(``fsn_fast_model_forward``, fp32 kernels)."""
from __future__ import annotations

import ctypes as C
import math
import os

import torch
import torch.nn as nn

from .. import _lib
from ..model.base_model import BaseModel
from ..model.module.sequence_model import SequenceModel


def melscale_fbanks(n_freqs: int, n_mels: int, sample_rate: int = 16000, f_min: float = 0.0, f_max: float = 8000.0):
    """HTK mel filterbank [n_freqs, n_mels] = torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk'),
    the buffer behind torchaudio.transforms.MelScale (fast_fullsubnet/model.py:57-63)."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    f_pts = 700.0 * (10 ** (torch.linspace(m_min, m_max, n_mels + 2) / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    return torch.max(torch.zeros(1), torch.min((-1.0 * slopes[:, :-2]) / f_diff[:-1], slopes[:, 2:] / f_diff[1:]))


class _MelScale(nn.Module):
    """Owns the ``fb`` buffer under the reference's key ``mel_scale.fb``."""

    def __init__(self, n_mels, n_stft):
        super().__init__()
        self.register_buffer("fb", melscale_fbanks(n_stft, n_mels))


class Model(BaseModel):
    def __init__(self, look_ahead, shrink_size, sequence_model, num_mels, encoder_input_size, bottleneck_hidden_size,
                 bottleneck_num_layers, noisy_input_num_neighbors, encoder_output_num_neighbors,
                 norm_type="offline_laplace_norm", weight_init=False, precision=None):
        super().__init__()
        assert sequence_model in ("GRU", "LSTM"), f"{self.__class__.__name__} only support GRU and LSTM."
        self.encoder = nn.Sequential(
            SequenceModel(input_size=64, hidden_size=384, output_size=0, num_layers=1, bidirectional=False,
                          sequence_model=sequence_model, output_activate_function=None),
            SequenceModel(input_size=384, hidden_size=257, output_size=64, num_layers=1, bidirectional=False,
                          sequence_model=sequence_model, output_activate_function="ReLU"))
        self.mel_scale = _MelScale(n_mels=num_mels, n_stft=encoder_input_size)
        self.bottleneck = SequenceModel(
            input_size=(noisy_input_num_neighbors * 2 + 1) + (encoder_output_num_neighbors * 2 + 1), output_size=1,
            hidden_size=bottleneck_hidden_size, num_layers=bottleneck_num_layers, bidirectional=False,
            sequence_model=sequence_model, output_activate_function="ReLU")
        self.decoder_lstm = nn.Sequential(
            SequenceModel(input_size=64 + 64, hidden_size=512, output_size=0, num_layers=1, bidirectional=False,
                          sequence_model=sequence_model, output_activate_function=None),
            SequenceModel(input_size=512, hidden_size=512, output_size=257 * 2, num_layers=1, bidirectional=False,
                          sequence_model=sequence_model, output_activate_function=None))
        self.shrink_size = shrink_size
        self.look_ahead = look_ahead
        self.num_mels = num_mels
        self.encoder_input_size = encoder_input_size
        self.noisy_input_num_neighbors = noisy_input_num_neighbors
        self.enc_output_num_neighbors = encoder_output_num_neighbors
        if norm_type != "offline_laplace_norm":
            raise NotImplementedError(f"norm_type {norm_type!r} is not built for fast_fullsubnet (SURVEY 8f)")
        self.norm = self.norm_wrapper(norm_type)
        # arithmetic: 'fp32' (FMA kernels) | 'f16x3_tc' (tensor cores, hi+lo split operands, the fp32 error class) |
        # 'f16_tc' (tensor cores, single pass, ~1e-4) | 'auto' = f16x3_tc when the shape allows, else fp32.  The
        # tensor-core modes run the bottleneck on the tcgen05 pair kernel and the encoder / decoder LSTMs + Linears
        # on the hoisted-GEMM + persistent-recurrence kernels (fsn_lstm_rec_tc.cu)
        self.precision = precision or os.environ.get("FSN_PRECISION", "auto")
        self._packed = None
        self._packed_key = None
        if num_mels != 64 or encoder_input_size != 257:
            raise NotImplementedError("the reference hard-codes 64 mel bins / 257 frequencies in its layer sizes")
        if weight_init:
            self.apply(self.weight_init)

    def _resolve_precision(self) -> str:
        d = self._desc(_lib.PREC["f16_tc"])
        ok = _lib.load().fsn_fast_packed_bytes(C.byref(d)) > 0
        if self.precision == "auto":
            return "f16x3_tc" if ok else "fp32"
        if self.precision not in ("fp32", "f16_tc", "f16x3_tc"):
            raise ValueError("precision must be 'fp32', 'f16x3_tc', 'f16_tc' or 'auto'")
        if self.precision != "fp32" and not ok:
            raise NotImplementedError("the tensor-core precisions need bottleneck_hidden_size = 384, 2 layers and input width <= 32")
        return self.precision

    def _desc(self, prec: int):
        return _lib.FastDesc(num_freqs=self.encoder_input_size, look_ahead=self.look_ahead, shrink_size=self.shrink_size,
                          num_mels=self.num_mels, enc1_hidden=384, enc2_hidden=257,
                          bn_hidden=self.bottleneck.hidden_size, bn_layers=self.bottleneck.num_layers, dec_hidden=512,
                          noisy_num_neighbors=self.noisy_input_num_neighbors,
                          enc_num_neighbors=self.enc_output_num_neighbors, precision=prec)

    def _structs(self, device):
        prec = self._resolve_precision()
        d = self._desc(_lib.PREC[prec])
        w = _lib.FastWeights()
        fb = self.mel_scale.fb
        if not fb.is_cuda:
            raise RuntimeError("fullsubnet_b200: call model.cuda() first - there is no CPU path.")
        w.mel_fb = fb.contiguous().data_ptr()
        w.enc1, w.enc2 = self.encoder[0].layer_struct(0), self.encoder[1].layer_struct(0)
        w.enc_fc_w, w.enc_fc_b = self.encoder[1].fc_ptrs()
        for l in range(self.bottleneck.num_layers):
            w.bn[l] = self.bottleneck.layer_struct(l)
        w.bn_fc_w, w.bn_fc_b = self.bottleneck.fc_ptrs()
        w.dec1, w.dec2 = self.decoder_lstm[0].layer_struct(0), self.decoder_lstm[1].layer_struct(0)
        w.dec_fc_w, w.dec_fc_b = self.decoder_lstm[1].fc_ptrs()
        w.bn_packed = None
        if prec != "fp32":  # tile-ordered fp16 image of the bottleneck weights, rebuilt when a parameter changes
            key = (self.bottleneck.version_key(), str(device), prec)
            if self._packed is None or self._packed_key != key:
                lib = _lib.load()
                buf = torch.empty(lib.fsn_fast_packed_bytes(C.byref(d)), dtype=torch.uint8, device=device)
                _lib.check(lib.fsn_fast_pack_bn_weights(C.byref(d), C.byref(w), buf.data_ptr(), _lib.stream_ptr(device)))
                self._packed, self._packed_key = buf, key
            w.bn_packed = self._packed.data_ptr()
        return d, w

    def forward(self, mix_mag):
        """mix_mag [B,1,F,T] -> [B,2,F,T]  (fast_fullsubnet/model.py:143-202)."""
        assert mix_mag.dim() == 4
        batch_size, num_channels, num_freqs, num_frames = mix_mag.size()
        assert num_channels == 1, f"{self.__class__.__name__} takes a magnitude feature as the input."
        assert num_freqs == self.encoder_input_size
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("fullsubnet_b200: backward kernels are not built yet; use torch.no_grad()/eval().")
        x = _lib.require_cuda(mix_mag, "mix_mag")
        lib = _lib.load()
        with torch.cuda.device(x.device):
            d, w = self._structs(x.device)
            n = lib.fsn_fast_workspace_bytes(C.byref(d), batch_size, num_frames)
            if n == 0:
                _lib.check_workspace(n)
            ws = torch.empty(n, dtype=torch.uint8, device=x.device)
            out = torch.empty(batch_size, 2, num_freqs, num_frames, dtype=torch.float32, device=x.device)
            _lib.check(lib.fsn_fast_model_forward(C.byref(d), C.byref(w), x.data_ptr(), batch_size, num_frames,
                                                  out.data_ptr(), ws.data_ptr(), n, _lib.stream_ptr(x.device)))
        return out

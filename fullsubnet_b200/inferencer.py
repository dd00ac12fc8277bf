"""Drop-in for the hot-path mode of recipes/dns_interspeech_2020/inferencer.py
(``Inferencer.full_band_crm_mask`` :130-145) and the parts of
audio_zen/inferencer/base_inferencer.py it relies on (attribute names ``model``, ``device``,
``torch_stft``, ``torch_istft``; ``_load_model`` :144-161).  Dataset / wav-file handling
(librosa, soundfile) is outside the hot path and not rebuilt here: construct
with a model, or with the reference's (config, checkpoint_path) pair."""
from __future__ import annotations

from functools import partial
from typing import Optional

import numpy as np
import torch

from .acoustics.feature import istft, stft
from .acoustics.mask import decompress_cIRM
from .utils import initialize_module, prepare_device


class Inferencer:
    def __init__(self, config: Optional[dict] = None, checkpoint_path=None, output_dir=None, model=None,
                 device=None):
        self.device = torch.device(device) if device is not None else prepare_device(torch.cuda.device_count())
        acoustics = (config or {}).get("acoustics", {"n_fft": 512, "hop_length": 256, "win_length": 512, "sr": 16000})
        self.acoustic_config = acoustics
        self.n_fft, self.hop_length = acoustics["n_fft"], acoustics["hop_length"]
        self.win_length, self.sr = acoustics["win_length"], acoustics.get("sr", 16000)
        self.torch_stft = partial(stft, n_fft=self.n_fft, hop_length=self.hop_length, win_length=self.win_length)
        self.torch_istft = partial(istft, n_fft=self.n_fft, hop_length=self.hop_length, win_length=self.win_length)
        if model is not None:
            self.model = model.to(self.device).eval()
        else:
            self.model, self.epoch = self._load_model(config["model"], checkpoint_path, self.device)
        self.inference_config = (config or {}).get("inferencer", {"type": "full_band_crm_mask", "args": {}})
        self.config = config

    @staticmethod
    def _load_model(model_config, checkpoint_path, device):
        """base_inferencer.py:144-161 (strict load, DDP 'module.' prefix stripped)."""
        model = initialize_module(model_config["path"], args=model_config["args"], initialize=True)
        ckpt = torch.load(checkpoint_path, map_location="cpu")
        sd = {k.replace("module.", ""): v for k, v in ckpt["model"].items()}
        model.load_state_dict(sd)
        model.to(device)
        model.eval()
        return model, ckpt["epoch"]

    @torch.no_grad()
    def full_band_crm_mask(self, noisy, inference_args=None):
        """inferencer.py:130-145, op by op through the drop-in functions: noisy [1,L] -> np.float32 [L]."""
        noisy_mag, _, noisy_real, noisy_imag = self.torch_stft(noisy)
        noisy_mag = noisy_mag.unsqueeze(1)
        pred_crm = self.model(noisy_mag)
        pred_crm = pred_crm.permute(0, 2, 3, 1)
        pred_crm = decompress_cIRM(pred_crm)
        enhanced_real = pred_crm[..., 0] * noisy_real - pred_crm[..., 1] * noisy_imag
        enhanced_imag = pred_crm[..., 1] * noisy_real + pred_crm[..., 0] * noisy_imag
        enhanced = self.torch_istft((enhanced_real, enhanced_imag), length=noisy.size(-1), input_type="real_imag")
        enhanced = enhanced.detach().squeeze(0).cpu().numpy()
        return enhanced

    @torch.no_grad()
    def enhance_batch(self, noisy: torch.Tensor) -> torch.Tensor:
        """The same path for B independent clips in ONE library call (fsn_enhance): pinned/host or device
        ``noisy`` [B,L] -> device tensor [B,L].  Equivalent to looping full_band_crm_mask over the clips."""
        x = noisy.to(self.device, non_blocking=True)
        if hasattr(self.model, "enhance"):  # fullsubnet: one fused library call
            return self.model.enhance(x, self.n_fft, self.hop_length, self.win_length)
        # other models (fast_fullsubnet): same flow, three library calls (stft -> model -> mask + istft)
        import ctypes as C  # noqa: F401
        from . import _lib
        B, L = x.shape
        F, T = self.n_fft // 2 + 1, 1 + L // self.hop_length
        buf = torch.empty(3, B, F, T, dtype=torch.float32, device=x.device)
        out = torch.empty(B, L, dtype=torch.float32, device=x.device)
        lib = _lib.load()
        with torch.cuda.device(x.device):
            st = _lib.stream_ptr(x.device)
            x = _lib.require_cuda(x, "noisy")
            _lib.check(lib.fsn_stft(x.data_ptr(), B, L, self.n_fft, self.hop_length, self.win_length, buf[0].data_ptr(),
                                    None, buf[1].data_ptr(), buf[2].data_ptr(), None, 0, st))
            crm = self.model(buf[0].unsqueeze(1)).contiguous()
            _lib.check(lib.fsn_istft(buf[1].data_ptr(), buf[2].data_ptr(), 1, crm.data_ptr(), B, T, self.n_fft,
                                     self.hop_length, self.win_length, L, out.data_ptr(), st))
        return out

    @torch.no_grad()
    def enhance_to_pcm(self, noisy: torch.Tensor) -> torch.Tensor:
        """enhance_batch + the int16 scaling of base_inferencer.py:181-182 on the device: noisy [B,L] -> int16 [B,L]
        (what the reference hands to ``sf.write``); only B*L*2 bytes come back to the host."""
        from . import _lib
        enhanced = self.enhance_batch(noisy)
        B, L = enhanced.shape
        pcm = torch.empty(B, L, dtype=torch.int16, device=enhanced.device)
        with torch.cuda.device(enhanced.device):
            _lib.check(_lib.load().fsn_peak_normalize_int16(enhanced.data_ptr(), B, L, 0.8 * float(np.iinfo(np.int16).max),
                                                          pcm.data_ptr(), _lib.stream_ptr(enhanced.device)))
        return pcm

    @staticmethod
    def write_wav(path, pcm, sr: int = 16000) -> None:
        """16-bit mono PCM file with the standard library (the reference uses soundfile, base_inferencer.py:183-187)."""
        import wave
        data = pcm.cpu().numpy() if isinstance(pcm, torch.Tensor) else np.asarray(pcm)
        with wave.open(str(path), "wb") as f:
            f.setnchannels(1)
            f.setsampwidth(2)
            f.setframerate(int(sr))
            f.writeframes(np.ascontiguousarray(data, dtype="<i2").tobytes())

    @torch.no_grad()
    def __call__(self, clips):
        """Host loop of base_inferencer.py:163-195 without the wav I/O: yields (enhanced float32, int16 PCM
        scaled as base_inferencer.py:181-182) per clip."""
        out = []
        for noisy in clips:
            enhanced = getattr(self, self.inference_config["type"])(noisy.to(self.device),
                                                                    self.inference_config.get("args", {}))
            amp = np.iinfo(np.int16).max
            pcm = np.int16(0.8 * amp * enhanced / np.max(np.abs(enhanced)))
            out.append((enhanced, pcm))
        return out

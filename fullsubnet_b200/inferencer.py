"""Drop-in for the hot-path mode of recipes/dns_interspeech_2020/inferencer.py
(``Inferencer.full_band_crm_mask`` :130-145) and the parts of
audio_zen/inferencer/base_inferencer.py it relies on (attribute names ``model``, ``device``,
``torch_stft``, ``torch_istft``; ``_load_model`` :144-161).  Dataset / wav-file handling
(librosa, soundfile) is outside the hot path and not rebuilt here: construct
with a model, or with the reference's (config, checkpoint_path) pair.  The host loop around the path (SURVEY 8f
rank 2) is here in batched form: ``enhance_files`` (wav load -> grouped batches -> fused enhance + int16 -> wav write)."""
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
        if hasattr(self.model, "enhance_pcm") and self.n_fft & (self.n_fft - 1) == 0:
            x = noisy.to(self.device, non_blocking=True)  # fused: peak in the iSTFT epilogue, one scaling pass
            return self.model.enhance_pcm(x, self.n_fft, self.hop_length, self.win_length,
                                          gain=0.8 * float(np.iinfo(np.int16).max))[1]
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

    # ------------------------------------------------------------------ wav files (base_inferencer.py:163-195, dataset_inference.py:39-43)
    @staticmethod
    def load_wav(path, sr: int = 16000) -> np.ndarray:
        """Mono float32 waveform at ``sr`` from a PCM wav file (8/16/24/32-bit; channels averaged like librosa).  The reference calls
        ``librosa.load(path, sr=sr)`` (dataset_inference.py:41): same int -> float scaling (1/32768 for 16-bit); when the
        file's rate differs, a windowed-sinc polyphase resampler stands in for librosa's soxr (not bit-identical to it -
        the hot-path parity contract starts at the 16 kHz waveform)."""
        import wave
        with wave.open(str(path), "rb") as f:
            nch, width, rate, n = f.getnchannels(), f.getsampwidth(), f.getframerate(), f.getnframes()
            raw = f.readframes(n)
        if width == 2:
            y = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
        elif width == 1:
            y = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
        elif width == 4:
            y = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
        elif width == 3:
            b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            y = (np.where(v >= 1 << 23, v - (1 << 24), v)).astype(np.float32) / 8388608.0
        else:
            raise NotImplementedError(f"wav sample width {width}")
        if nch > 1:  # librosa.load(mono=True): mean over the channels
            y = y.reshape(-1, nch).mean(axis=1).astype(np.float32)
        if rate != sr:
            y = Inferencer.resample(y, rate, sr)
        return np.ascontiguousarray(y, dtype=np.float32)

    @staticmethod
    def resample(y: np.ndarray, sr_in: int, sr_out: int, zeros: int = 24) -> np.ndarray:
        """Band-limited (hann-windowed sinc) polyphase resampling, host side."""
        from math import gcd
        g = gcd(int(sr_in), int(sr_out))
        up, down = sr_out // g, sr_in // g
        cutoff = min(1.0, up / down)
        half = int(np.ceil(zeros / cutoff))
        n_out = int(np.ceil(len(y) * up / down))
        t_out = np.arange(n_out, dtype=np.float64) * down / up  # positions in input samples
        base = np.floor(t_out).astype(np.int64)
        k = np.arange(-half + 1, half + 1)
        idx = base[:, None] + k[None, :]
        d = t_out[:, None] - idx
        w = cutoff * np.sinc(cutoff * d) * (0.5 + 0.5 * np.cos(np.pi * np.clip(d / half, -1, 1)))
        ok = (idx >= 0) & (idx < len(y))
        vals = np.where(ok, y[np.clip(idx, 0, len(y) - 1)], 0.0)
        return (vals * w).sum(axis=1).astype(np.float32)

    @torch.no_grad()
    def enhance_files(self, paths, output_dir, batch_size: int = 64, sr=None):
        """Batched form of the host loop of base_inferencer.py:163-195: files are grouped by length (a clip's result
        depends on its own length through the per-clip norms, so clips are never padded), each group goes through ONE
        fused library call per ``batch_size`` clips (pinned staging buffer -> H2D -> fsn_enhance_pcm -> int16 D2H), and
        ``<output_dir>/<stem>.wav`` is written as 16-bit PCM like the reference.  Returns the written paths."""
        from collections import defaultdict
        from pathlib import Path
        sr = int(sr or self.sr)
        out_dir = Path(output_dir)
        out_dir.mkdir(parents=True, exist_ok=True)
        clips = [(Path(p), self.load_wav(p, sr)) for p in paths]
        groups = defaultdict(list)
        for i, (_, y) in enumerate(clips):
            groups[len(y)].append(i)
        written = [None] * len(clips)
        for L, idxs in sorted(groups.items()):
            for s0 in range(0, len(idxs), batch_size):
                chunk = idxs[s0:s0 + batch_size]
                stage = torch.empty(len(chunk), L, dtype=torch.float32).pin_memory()
                for r, i in enumerate(chunk):
                    stage[r] = torch.from_numpy(clips[i][1])
                pcm = self.enhance_to_pcm(stage).cpu().numpy()
                for r, i in enumerate(chunk):
                    dst = out_dir / f"{clips[i][0].stem}.wav"
                    self.write_wav(dst, pcm[r], sr)
                    written[i] = dst
        return written

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

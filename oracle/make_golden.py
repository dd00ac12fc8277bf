"""TEST INFRASTRUCTURE ONLY.  Generates ``tests/golden/*.npz`` by running the UNMODIFIED
upstream code from ``/root/reference`` on CPU (this only works in the build container; the
GPU box has no /root/reference, which is why the outputs are committed as fixtures).

Import recipe: SURVEY.md section 8c - ``librosa`` / ``soundfile`` are not installed and are
stubbed; nothing on the hot path uses them.

Run:  python oracle/make_golden.py
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("FSN_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)


def import_reference():
    sys.path[:0] = [REF, os.path.join(REF, "recipes", "dns_interspeech_2020")]
    for m in ("librosa", "soundfile"):
        sys.modules.setdefault(m, types.ModuleType(m))
    from audio_zen.acoustics import feature, mask  # noqa
    from fullsubnet.model import Model  # noqa
    from inferencer import Inferencer  # noqa
    return feature, mask, Model, Inferencer


def main():
    from functools import partial

    from oracle import fullsubnet_oracle as O

    feature, mask, Model, Inferencer = import_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.set_num_threads(8)

    # ------------------------------------------------------------------ DSP ops
    y = O.make_noisy(3, 3000, seed=1, speechlike=True)
    mag, phase, real, imag = feature.stft(y, 512, 256, 512)
    wav_rt = feature.istft((real, imag), 512, 256, 512, length=y.shape[-1], input_type="real_imag")
    wav_nolen = feature.istft(torch.complex(real, imag), 512, 256, 512)
    y3 = O.make_noisy(4, 1500, seed=2).reshape(2, 2, 1500)
    mag3 = feature.stft(y3, 512, 256, 512)[0]
    # small-FFT variant
    mag_s, _, real_s, imag_s = feature.stft(y, 64, 32, 64)
    wav_s = feature.istft((real_s, imag_s), 64, 32, 64, length=y.shape[-1], input_type="real_imag")
    g = torch.Generator().manual_seed(5)
    m = 12 * torch.randn(2, 9, 7, 2, generator=g)
    m[0, 0, 0, 0], m[0, 0, 1, 0], m[0, 0, 2, 0] = 9.9, -9.9, 9.95
    dec = mask.decompress_cIRM(m)
    big = 60 * torch.randn(2, 9, 7, 2, generator=g) - 40
    comp = mask.compress_cIRM(big)
    yc = O.make_noisy(3, 3000, seed=3)
    _, _, cr, ci = feature.stft(yc, 512, 256, 512)
    cirm = mask.build_complex_ideal_ratio_mask(real, imag, cr, ci)
    xb = torch.randn(5, 2, 9, 4, generator=g)
    db2 = feature.drop_band(xb, 2)
    db3 = feature.drop_band(xb, 3)
    np.savez_compressed(
        os.path.join(out_dir, "dsp.npz"),
        y=y.numpy(), mag=mag.numpy(), phase=phase.numpy(), real=real.numpy(), imag=imag.numpy(),
        wav_rt=wav_rt.numpy(), wav_nolen=wav_nolen.numpy(), y3=y3.numpy(), mag3=mag3.numpy(),
        mag_s=mag_s.numpy(), real_s=real_s.numpy(), imag_s=imag_s.numpy(), wav_s=wav_s.numpy(),
        m=m.numpy(), dec=dec.numpy(), big=big.numpy(), comp=comp.numpy(),
        yc=yc.numpy(), cirm=cirm.numpy(), xb=xb.numpy(), db2=db2.numpy(), db3=db3.numpy(),
    )

    # ------------------------------------------------------------ small model
    small = dict(num_freqs=33, look_ahead=2, sequence_model="LSTM", fb_num_neighbors=0, sb_num_neighbors=3,
                 fb_output_activate_function="ReLU", sb_output_activate_function=False,
                 fb_model_hidden_size=32, sb_model_hidden_size=24, norm_type="offline_laplace_norm",
                 num_groups_in_drop_band=2, weight_init=False)
    sd = O.make_state_dict(seed=7, args=small)
    model = Model(**small).eval()
    model.load_state_dict(sd, strict=True)
    ys = O.make_noisy(3, 1200, seed=4, speechlike=True)
    mag_small = feature.stft(ys, 64, 32, 64)[0]
    with torch.no_grad():
        crm_b1 = model(mag_small[:1].unsqueeze(1))
        crm_g2 = model(mag_small.unsqueeze(1))  # B=3 -> drop_band, G=2
        model.num_groups_in_drop_band = 1
        crm_g1 = model(mag_small.unsqueeze(1))
        fb_out = model.fb_model(model.norm(torch.nn.functional.pad(mag_small.unsqueeze(1), [0, 2])).reshape(3, 33, -1))
    np.savez_compressed(
        os.path.join(out_dir, "model_small.npz"),
        ys=ys.numpy(), mag=mag_small.numpy(), crm_b1=crm_b1.numpy(), crm_g2=crm_g2.numpy(),
        crm_g1=crm_g1.numpy(), fb_out=fb_out.numpy(),
        **{"sd." + k: v.numpy() for k, v in sd.items()},
    )

    # ------------------------------------------------------------- full model
    full = dict(O.DEFAULT_MODEL_ARGS)
    res = {}
    for tag, gain in (("wa", 1.0), ("wb", 220.0)):
        sd = O.make_state_dict(seed=0, args=full, sb_fc_gain=gain)
        model = Model(**full).eval()
        model.load_state_dict(sd, strict=True)
        inf = Inferencer.__new__(Inferencer)  # no dataset / checkpoint (SURVEY 8c recipe)
        inf.model = model
        inf.device = torch.device("cpu")
        inf.torch_stft = partial(feature.stft, n_fft=512, hop_length=256, win_length=512)
        inf.torch_istft = partial(feature.istft, n_fft=512, hop_length=256, win_length=512)
        yf = O.make_noisy(2, 8000, seed=0, speechlike=True)
        with torch.no_grad():
            magf = feature.stft(yf, 512, 256, 512)[0]
            crm = torch.cat([model(magf[i:i + 1].unsqueeze(1)) for i in range(2)], 0)
            wav = np.stack([inf.full_band_crm_mask(yf[i:i + 1], {}) for i in range(2)], 0)
        res[f"{tag}_crm"] = crm.numpy()
        res[f"{tag}_wav"] = wav
        res["y"] = yf.numpy()
        print(tag, "crm range", float(crm.min()), float(crm.max()), "wav max", float(np.abs(wav).max()))
    np.savez_compressed(os.path.join(out_dir, "model_full.npz"), **res)
    # ------------------------------------------------------- fast_fullsubnet (A13)
    from oracle import fast_fullsubnet_oracle as FO
    ti = types.ModuleType("torchinfo"); ti.summary = lambda *a, **k: None
    sys.modules.setdefault("torchinfo", ti)
    from fast_fullsubnet.model import Model as FastModel
    fargs = dict(FO.DEFAULT_FAST_ARGS)
    fsd = FO.make_fast_state_dict(seed=3, args=fargs)
    fm = FastModel(**fargs).eval()
    ref_fb = fm.state_dict()["mel_scale.fb"].clone()
    fm.load_state_dict(fsd, strict=True)
    yq = O.make_noisy(3, 6000, seed=9, speechlike=True)
    magq = feature.stft(yq, 512, 256, 512)[0]
    with torch.no_grad():
        fo_b1 = fm(magq[:1].unsqueeze(1))
        fo_b3 = fm(magq.unsqueeze(1))
    np.savez_compressed(os.path.join(out_dir, "fast_full.npz"), y=yq.numpy(), mag=magq.numpy(),
                        out_b1=fo_b1.numpy(), out_b3=fo_b3.numpy(), mel_fb=ref_fb.numpy())
    print("fast crm range", float(fo_b3.min()), float(fo_b3.max()))
    # ------------------------------------------------------- improved_fullsubnet (A14)
    from oracle import improved_fullsubnet_oracle as IO
    from improved_fullsubnet.model import Model as ImpModel
    res = {}
    for tag, iargs, L in (("k16", dict(IO.DEFAULT_IMPROVED_ARGS), 4000), ("k48", dict(IO.ARGS_48K_1024), 12000)):
        isd = IO.make_improved_state_dict(seed=5, args=iargs)
        im = ImpModel(**iargs).eval()
        assert [k for k, _ in IO.improved_state_dict_shapes(iargs)] == list(im.state_dict().keys())
        im.load_state_dict(isd, strict=True)
        yi = O.make_noisy(2, L, seed=13, speechlike=True)
        with torch.no_grad():
            wi = im(yi)
        res[tag + "_y"] = yi.numpy(); res[tag + "_wav"] = wi.numpy()
        print("improved", tag, tuple(wi.shape), float(wi.abs().max()), sum(v.numel() for v in isd.values()))
    np.savez_compressed(os.path.join(out_dir, "improved.npz"), **res)
    for f in sorted(os.listdir(out_dir)):
        print(f, os.path.getsize(os.path.join(out_dir, f)))


if __name__ == "__main__":
    main()

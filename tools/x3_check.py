"""GPU check of the sub-band precisions: parity vs tests/golden/model_full.npz (W-a, W-b) and stage timing at a
given batch.  usage: python tools/x3_check.py [B] [precisions...]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fullsubnet_oracle as O
from fullsubnet_b200.fullsubnet.model import Model
from fullsubnet_b200 import _lib
import ctypes as C

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
precs = sys.argv[2:] or ["f16_tc", "f16x3_tc", "fp32"]
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "model_full.npz"))
y = torch.from_numpy(g["y"]).to(dev)
for prec in precs:
    for tag, gain in (("wa", 1.0), ("wb", 220.0)):
        m = Model(**O.DEFAULT_MODEL_ARGS, precision=prec)
        m.load_state_dict(O.make_state_dict(seed=0, sb_fc_gain=gain), strict=True)
        m = m.to(dev).eval()
        wav, crm = m.enhance(y, return_crm=True)
        torch.cuda.synchronize()
        rc, rw = g[f"{tag}_crm"], g[f"{tag}_wav"]
        e_crm = float(np.abs(crm.cpu().numpy() - rc).max() / np.abs(rc).max())
        e_wav = float(np.abs(wav.cpu().numpy() - rw).max())
        print(f"[parity] {prec:9s} {tag}: crm max-rel {e_crm:.3e}  wav max-abs {e_wav:.3e}  (|wav|max {np.abs(rw).max():.2f})", flush=True)
lib = _lib.load()
lib.fsn_set_profiling(1)
lib.fsn_last_stage_ms.restype = C.c_float
yb = O.make_noisy(B, 64000, seed=1).to(dev)
for prec in precs:
    if prec == "fp32" and B > 32:
        continue
    m = Model(**O.DEFAULT_MODEL_ARGS, precision=prec)
    m.load_state_dict(O.make_state_dict(seed=0), strict=True)
    m = m.to(dev).eval()
    for i in range(3):
        t0 = time.time(); out = m.enhance(yb); torch.cuda.synchronize(); dt = time.time() - t0
        st = [lib.fsn_last_stage_ms(s) for s in range(4)]
        print(f"[time] {prec:9s} B={B}: wall {dt*1e3:.1f} ms  stages stft {st[0]:.2f} fb {st[1]:.2f} sb {st[2]:.2f} istft {st[3]:.2f}", flush=True)
    assert torch.isfinite(out).all()

"""Time of the sub-band stage alone (fsn_last_stage_ms) for batch sizes around one wave of clusters."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fullsubnet_b200 import _lib
from fullsubnet_b200.fullsubnet.model import Model
from oracle import fullsubnet_oracle as O
dev = torch.device("cuda:0")
lib = _lib.load()
m = Model(**O.DEFAULT_MODEL_ARGS, precision=sys.argv[1] if len(sys.argv) > 1 else "f16_tc")
m.load_state_dict(O.make_state_dict(seed=0), strict=True)
m = m.to(dev).eval()
for B in (1, 5, 10, 18, 36, 72, 256):
    y = O.make_noisy(B, 64000, seed=3).to(dev)
    for _ in range(2):
        m.enhance(y)
    torch.cuda.synchronize()
    lib.fsn_set_profiling(1)
    ts = []
    for _ in range(3):
        m.enhance(y)
        torch.cuda.synchronize()
        ts.append(lib.fsn_last_stage_ms(2))
    lib.fsn_set_profiling(0)
    R = B * 257
    print(f"B={B:4d} rows {R:6d} (pairs {-(-R // 64)}, 6-CTA clusters {-(-R // 128)}): sub-band stage {sorted(ts)[1]:.2f} ms", flush=True)

"""GPU timing sweep of the sub-band stage (uses the library's stage events)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fullsubnet_b200 import _lib
from fullsubnet_b200.fullsubnet.model import Model
from oracle import fullsubnet_oracle as O
dev = torch.device("cuda:0")
lib = _lib.load()
sd = O.make_state_dict(0)
m = Model(**O.DEFAULT_MODEL_ARGS, precision="f16_tc"); m.load_state_dict(sd); m = m.to(dev).eval()
L = int(sys.argv[1]) if len(sys.argv) > 1 else 32000
Tp = 1 + L // 256 + 2
for B in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "2,4,9,18,36,72")]:
    y = O.make_noisy(B, L, seed=1).to(dev)
    for _ in range(2): m.enhance(y)
    torch.cuda.synchronize()
    lib.fsn_set_profiling(1)
    m.enhance(y); torch.cuda.synchronize()
    sb, fb = lib.fsn_last_stage_ms(2), lib.fsn_last_stage_ms(1)
    lib.fsn_set_profiling(0)
    ctas = (B * 257 + 31) // 32
    waves = (ctas + 147) // 148
    print(f"B={B:4d} ctas={ctas:5d} waves={waves:3d} sb={sb:8.3f} ms  fb={fb:7.3f} ms  per-step-per-wave={1e3*sb/(Tp*waves):7.2f} us "
          f"stages={os.environ.get('FSN_TC_STAGES','6')}", flush=True)

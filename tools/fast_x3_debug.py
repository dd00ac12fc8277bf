"""fast_fullsubnet: error of a precision mode against the fp32 kernels at a given length (debug)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fullsubnet_b200.acoustics.feature import stft
from fullsubnet_b200.fast_fullsubnet.model import Model
from oracle import fast_fullsubnet_oracle as FO, fullsubnet_oracle as O
dev = torch.device("cuda:0")
L = int(sys.argv[1]) if len(sys.argv) > 1 else 64000
y = O.make_noisy(2, L, seed=43, speechlike=True).to(dev)
mag = stft(y, 512, 256, 512)[0].unsqueeze(1)
outs = {}
for prec in ("fp32", "f16x3_tc", "f16_tc"):
    m = Model(**FO.DEFAULT_FAST_ARGS, precision=prec)
    m.load_state_dict(FO.make_fast_state_dict(seed=3), strict=True)
    m = m.to(dev).eval()
    with torch.no_grad():
        outs[prec] = m(mag)
ref = outs["fp32"]
for prec in ("f16x3_tc", "f16_tc"):
    d = (outs[prec] - ref)
    print(f"L={L} {prec}: max-rel {float(d.abs().max() / ref.abs().max()):.2e} rel-l2 {float(d.norm() / ref.norm()):.2e}",
          "env", {k: v for k, v in os.environ.items() if k.startswith("FSN_")})

"""4-CTA-cluster sub-band kernel (FSN_TC_CLUSTER4=1, precision f16_tc) against the fp32 kernels and the pair kernel."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fullsubnet_b200.fullsubnet.model import Model
from oracle import fullsubnet_oracle as O
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
L = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
y = O.make_noisy(B, L, seed=3, speechlike=True).to(dev)
sd = O.make_state_dict(seed=0)
outs = {}
for prec in ("fp32", "f16_tc"):
    m = Model(**O.DEFAULT_MODEL_ARGS, precision=prec)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).eval()
    wav, crm = m.enhance(y, return_crm=True)
    torch.cuda.synchronize()
    if prec == "f16_tc" and B >= 64:
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(3):
            m.enhance(y)
        e1.record(); torch.cuda.synchronize()
        print(f"f16_tc B={B}: {e0.elapsed_time(e1) / 3:.2f} ms per call")
    outs[prec] = (wav, crm)
d = outs["f16_tc"][1] - outs["fp32"][1]
ref = outs["fp32"][1]
print(f"B={B} L={L} cluster4={os.environ.get('FSN_TC_CLUSTER4')}: cRM max-rel {float(d.abs().max() / ref.abs().max()):.2e} rel-l2 "
      f"{float(d.norm() / ref.norm()):.2e}; wav max-abs {float((outs['f16_tc'][0] - outs['fp32'][0]).abs().max()):.2e}; finite "
      f"{bool(torch.isfinite(outs['f16_tc'][1]).all())}")

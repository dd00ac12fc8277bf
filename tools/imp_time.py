import torch, time, sys
sys.path.insert(0, '.')
from fullsubnet_b200.improved_fullsubnet.model import Model
from oracle import improved_fullsubnet_oracle as IO
import os
only = sys.argv[1] if len(sys.argv) > 1 else None
for name, args, B, L in (("k16", IO.DEFAULT_IMPROVED_ARGS, 128, 64000), ("k48", IO.ARGS_48K_1024, 128, 96000)):
    if only and name != only:
        continue
    m = Model(**args); m.load_state_dict(IO.make_improved_state_dict(5, args)); m = m.cuda().eval()
    m.precision = os.environ.get("FSN_IMPROVED_PRECISION", "auto")
    y = 0.1 * torch.randn(B, L, device='cuda')
    with torch.no_grad():
        for _ in range(2): m(y)
        torch.cuda.synchronize(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True)
        e0.record()
        for _ in range(3): m(y)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    T = 1 + L // args["hop_length"]
    print(f"improved {name} {m._resolve_precision()}: B={B} L={L} {ms:.1f} ms/step  {B*T/ms*1e3:.0f} frames/s")

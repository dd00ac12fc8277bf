"""fsn_debug_linear_tc against a float64 matmul (debug / accuracy sweep)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fullsubnet_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")


def run(rows, K, N, x3, act=0):
    g = torch.Generator().manual_seed(rows + K + N)
    x = torch.randn(rows, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    ref = x.double() @ W.double().T + b.double()
    if act == 1:
        ref = ref.clamp_min(0)
    Hm = max(8, (N + 3) // 4)
    n = lib.fsn_debug_lstm_tc_workspace_bytes(rows, 1, K, Hm, x3)
    ws = torch.empty(n, dtype=torch.uint8, device=dev)
    out = torch.full((rows, N), float("nan"), device=dev)
    xd, Wd, bd = x.to(dev), W.to(dev), b.to(dev)
    _lib.check(lib.fsn_debug_linear_tc(xd.data_ptr(), rows, K, Wd.data_ptr(), bd.data_ptr(), N, act, x3, out.data_ptr(),
                                       ws.data_ptr(), n, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    e = (out.cpu().double() - ref).abs()
    bad = (e > 1e-2).nonzero()
    print(f"rows={rows} K={K} N={N} x3={x3}: max err {float(e.max()):.2e} rel-l2 {float(e.norm() / ref.norm()):.2e}"
          + (f"  first bad {bad[0].tolist()} of {len(bad)}" if len(bad) else ""), flush=True)


if __name__ == "__main__":
    for x3 in (1, 0):
        for rows in (52, 506, 64768):
            run(rows, 257, 64, x3, act=1)
            run(rows, 512, 514, x3)
            run(rows, 512, 257, x3, act=1)
            run(rows, 257, 2048, x3)

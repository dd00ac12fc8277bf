"""One tensor-core LSTM layer (fsn_debug_lstm_layer_tc) against a float64 nn.LSTM on the CPU (debug / accuracy sweep)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fullsubnet_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")


def run(R, T, K, H, x3, seed=0, reps=2):
    g = torch.Generator().manual_seed(seed)
    k = 1.0 / H ** 0.5
    w_ih = (torch.rand(4 * H, K, generator=g) * 2 - 1) * k
    w_hh = (torch.rand(4 * H, H, generator=g) * 2 - 1) * k
    b_ih = (torch.rand(4 * H, generator=g) * 2 - 1) * k
    b_hh = (torch.rand(4 * H, generator=g) * 2 - 1) * k
    x = torch.randn(R, T, K, generator=g)
    lstm = torch.nn.LSTM(K, H, batch_first=True).double()
    with torch.no_grad():
        lstm.weight_ih_l0.copy_(w_ih); lstm.weight_hh_l0.copy_(w_hh); lstm.bias_ih_l0.copy_(b_ih); lstm.bias_hh_l0.copy_(b_hh)
        ref = lstm(x.double())[0]
    n = lib.fsn_debug_lstm_tc_workspace_bytes(R, T, K, H, x3)
    ws = torch.empty(n, dtype=torch.uint8, device=dev)
    d = [t.to(dev).contiguous() for t in (w_ih, w_hh, b_ih, b_hh, x)]
    outs = []
    for _ in range(reps):
        out = torch.full((R, T, H), float("nan"), device=dev)
        _lib.check(lib.fsn_debug_lstm_layer_tc(*[t.data_ptr() for t in d], R, T, K, H, x3, out.data_ptr(), ws.data_ptr(), n,
                                               torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        outs.append(out.cpu().double())
    e = (outs[0] - ref).abs()
    per_t = e.amax(dim=(0, 2))
    print(f"R={R} T={T} K={K} H={H} x3={x3}: max err {float(e.max()):.2e} (first 8 steps {float(per_t[:8].max()):.1e}, last 8 "
          f"{float(per_t[-8:].max()):.1e}), run-to-run max diff {float((outs[0] - outs[-1]).abs().max()):.1e}", flush=True)


if __name__ == "__main__":
    for x3 in (1, 0):
        run(2, 26, 64, 384, x3)
        run(2, 253, 64, 384, x3)
        run(2, 253, 384, 257, x3)
        run(2, 253, 128, 512, x3)
        run(256, 253, 257, 512, x3)
        run(300, 40, 257, 512, x3)

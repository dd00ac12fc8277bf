"""fast_fullsubnet encoder layers on the tensor-core hook with the REAL inputs/weights vs float64 (debug)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fullsubnet_b200 import _lib
from oracle import fast_fullsubnet_oracle as FO, fullsubnet_oracle as O
lib = _lib.load()
dev = torch.device("cuda:0")
L = int(sys.argv[1]) if len(sys.argv) > 1 else 64000
sd = FO.make_fast_state_dict(seed=3)
y = O.make_noisy(2, L, seed=43, speechlike=True)
mag = O.stft(y, 512, 256, 512)[0].unsqueeze(1)
x = torch.nn.functional.pad(mag, [0, 2])
mel = (x.transpose(-1, -2) @ sd["mel_scale.fb"]).transpose(-1, -2)
enc_in = O.offline_laplace_norm(mel).reshape(2, 64, -1).permute(0, 2, 1).contiguous()  # [B,T,64]
print("enc_in range", float(enc_in.min()), float(enc_in.max()), "T", enc_in.shape[1])


def layer(xin, pre, K, H, x3):
    R, T, _ = xin.shape
    names = [pre + n for n in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")]
    lstm = torch.nn.LSTM(K, H, batch_first=True).double()
    with torch.no_grad():
        for p, n in zip((lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0), names):
            p.copy_(sd[n])
        ref = lstm(xin.double())[0]
    n = lib.fsn_debug_lstm_tc_workspace_bytes(R, T, K, H, x3)
    ws = torch.empty(n, dtype=torch.uint8, device=dev)
    d = [sd[k].to(dev).contiguous() for k in names] + [xin.float().to(dev).contiguous()]
    out = torch.full((R, T, H), float("nan"), device=dev)
    _lib.check(lib.fsn_debug_lstm_layer_tc(*[t.data_ptr() for t in d], R, T, K, H, x3, out.data_ptr(), ws.data_ptr(), n,
                                           torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    e = (out.cpu().double() - ref).abs()
    pt = e.amax(dim=(0, 2))
    print(f"{pre} K={K} H={H} x3={x3}: max err {float(e.max()):.2e} at t={int(pt.argmax())}; per-t max at [0,1,2,10,100,-1]: "
          f"{[f'{float(pt[i]):.1e}' for i in (0, 1, 2, 10, min(100, T - 1), -1)]}  |ref|max {float(ref.abs().max()):.2f}")
    return ref.float()


for x3 in (1, 0):
    h = layer(enc_in, "encoder.0.sequence_model.", 64, 384, x3)
    layer(h, "encoder.1.sequence_model.", 384, 257, x3)

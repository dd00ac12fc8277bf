"""Runs N training steps of BASELINE configs[2] (for ncu launch lists): python tools/train_steps.py [steps] [batch]"""
import sys

import torch

sys.path.insert(0, ".")
from fullsubnet_b200.fullsubnet.model import Model
from fullsubnet_b200.loss import mse_loss
from fullsubnet_b200.optim import FusedClipAdam
from fullsubnet_b200.trainer import Trainer
from oracle import fullsubnet_oracle as O

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda:0")
m = Model(**dict(O.DEFAULT_MODEL_ARGS, weight_init=False))
m.load_state_dict(O.make_state_dict(seed=0))
m = m.to(dev).train()
cfg = {"meta": {"use_amp": False, "save_dir": "/tmp/fsn", "experiment_name": "p"},
       "acoustics": {"n_fft": 512, "hop_length": 256, "win_length": 512},
       "trainer": {"train": {"epochs": 1, "save_checkpoint_interval": 1, "clip_grad_norm_value": 10}}}
tr = Trainer(None, 0, cfg, False, False, m, mse_loss(), FusedClipAdam(m.parameters(), lr=1e-3), None, None)
noisy, clean = O.make_noisy(B, 48000, seed=0).to(dev), (0.5 * O.make_noisy(B, 48000, seed=100)).to(dev)
for i in range(steps):
    loss = tr.train_step(noisy, clean)
torch.cuda.synchronize()
print("loss", float(loss))

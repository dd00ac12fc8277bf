"""Mirror of the optimisation loop of recipes/dns_interspeech_2020/fullsubnet/trainer.py:14-76 on top of
audio_zen/trainer/base_trainer.py:28-218 - the part of the trainer that is on the hot path:
STFT of noisy/clean, cIRM target + drop_band, Model.forward, MSE, backward, gradient mean over ranks, clip, Adam.

Same constructor arguments and config keys as the reference (``meta.use_amp`` is accepted and ignored: the kernels
compute in fp32, which is at least the precision of the reference's fp16 autocast; ``scaler`` is kept in the
checkpoint schema {epoch, best_score, optimizer, scaler, model} of base_trainer.py:208-218 as an empty dict).
Validation, TensorBoard and audio visualisation (trainer.py:78-181) are outside the hot path and not built:
``validation_dataloader`` must be None.

Two gradient paths, both one NCCL all-reduce of gradients per step:
  * the model may be wrapped in DistributedDataParallel exactly like base_trainer.py:32 - the autograd Function
    behind Model.forward delivers the gradients to DDP's hooks;
  * default here: ``model.flat_grad()`` makes every ``p.grad`` a view of one flat buffer, ``dist.all_reduce`` moves
    that buffer once, and FusedClipAdam folds the 1/world mean into its clip coefficient."""
from __future__ import annotations

from functools import partial
from pathlib import Path

import torch

from .acoustics.feature import drop_band, istft, stft
from .acoustics.mask import build_complex_ideal_ratio_mask
from .optim import FusedClipAdam


class Trainer:
    def __init__(self, dist, rank, config, resume, only_validation, model, loss_function, optimizer,
                 train_dataloader, validation_dataloader=None):
        if validation_dataloader is not None or only_validation:
            raise NotImplementedError("fullsubnet_b200.Trainer builds the training step only (validation: SURVEY 8f rank 4)")
        self.dist, self.rank = dist, rank
        self.device = torch.device("cuda", rank)
        self.model = model.cuda(rank)
        self.loss_function = loss_function
        self.optimizer = optimizer
        self.world_size = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
        self.use_amp = config["meta"].get("use_amp", False)
        ac = config["acoustics"]
        self.torch_stft = partial(stft, n_fft=ac["n_fft"], hop_length=ac["hop_length"], win_length=ac["win_length"])
        self.torch_istft = partial(istft, n_fft=ac["n_fft"], hop_length=ac["hop_length"], win_length=ac["win_length"])
        self.train_config = config["trainer"]["train"]
        self.epochs = self.train_config["epochs"]
        self.save_checkpoint_interval = self.train_config["save_checkpoint_interval"]
        self.clip_grad_norm_value = self.train_config["clip_grad_norm_value"]
        assert self.save_checkpoint_interval >= 1, \
            "Check the 'save_checkpoint_interval' parameter in the config. It should be large than one."
        self.start_epoch = 1
        self.best_score = float("-inf")
        self.save_dir = Path(config["meta"]["save_dir"]).expanduser().absolute() / config["meta"]["experiment_name"]
        self.checkpoints_dir = self.save_dir / "checkpoints"
        self.train_dataloader = train_dataloader
        if isinstance(optimizer, FusedClipAdam):
            optimizer.max_norm = self.clip_grad_norm_value
        if resume:
            self._resume_checkpoint()

    # ------------------------------------------------------------------ one optimisation step (trainer.py:41-71)
    def train_step(self, noisy, clean):
        model = self.model
        self.optimizer.zero_grad(set_to_none=False)
        noisy = noisy.to(self.device, non_blocking=True)
        clean = clean.to(self.device, non_blocking=True)
        noisy_mag, _, noisy_real, noisy_imag = self.torch_stft(noisy)
        _, _, clean_real, clean_imag = self.torch_stft(clean)
        cIRM = build_complex_ideal_ratio_mask(noisy_real, noisy_imag, clean_real, clean_imag)  # [B, F, T, 2]
        cIRM = drop_band(cIRM.permute(0, 3, 1, 2), model.num_groups_in_drop_band).permute(0, 2, 3, 1)
        cRM = model(noisy_mag.unsqueeze(1)).permute(0, 2, 3, 1)
        loss = self.loss_function(cIRM, cRM)
        loss.backward()
        scale = 1.0
        if self.world_size > 1:  # DDP's mean all-reduce (base_trainer.py:32) as one collective over the flat buffer
            flat = model.flat_grad()
            self.dist.all_reduce(flat)
            scale = 1.0 / self.world_size
        if isinstance(self.optimizer, FusedClipAdam):
            self.optimizer.step(grad_scale=scale)
        else:
            if scale != 1.0:
                for p in model.parameters():
                    p.grad.mul_(scale)
            torch.nn.utils.clip_grad_norm_(model.parameters(), self.clip_grad_norm_value)
            self.optimizer.step()
        return loss.detach()

    def _train_epoch(self, epoch):
        loss_total = torch.zeros((), device=self.device)
        for noisy, clean in self.train_dataloader:
            loss_total += self.train_step(noisy, clean)
        return float(loss_total) / max(1, len(self.train_dataloader))  # the step loop itself never synchronises

    def train(self):
        for epoch in range(self.start_epoch, self.epochs + 1):
            self.model.train()
            self.last_epoch_loss = self._train_epoch(epoch)
            if self.rank == 0 and self.save_checkpoint_interval != 0 and epoch % self.save_checkpoint_interval == 0:
                self._save_checkpoint(epoch)

    # ------------------------------------------------------------------ checkpoints (base_trainer.py:170-252)
    def _save_checkpoint(self, epoch, is_best_epoch=False):
        state = {"epoch": epoch, "best_score": self.best_score, "optimizer": self.optimizer.state_dict(), "scaler": {},
                 "model": self.model.state_dict()}
        self.checkpoints_dir.mkdir(parents=True, exist_ok=True)
        torch.save(state, (self.checkpoints_dir / "latest_model.tar").as_posix())
        torch.save(state["model"], (self.checkpoints_dir / f"model_{str(epoch).zfill(4)}.pth").as_posix())
        if is_best_epoch:
            torch.save(state, (self.checkpoints_dir / "best_model.tar").as_posix())

    def _resume_checkpoint(self):
        path = self.checkpoints_dir.expanduser().absolute() / "latest_model.tar"
        assert path.exists(), f"{path} does not exist, can not load latest checkpoint."
        ckpt = torch.load(path.as_posix(), map_location="cpu")
        self.start_epoch = ckpt["epoch"] + 1
        self.best_score = ckpt["best_score"]
        self.optimizer.load_state_dict(ckpt["optimizer"])
        self.model.load_state_dict(ckpt["model"])

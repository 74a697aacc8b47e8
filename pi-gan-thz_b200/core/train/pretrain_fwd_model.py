"""Drop-in ``core.train.pretrain_fwd_model`` — same ``pretrain_forward_model(...)`` signature, return value and
output files as the reference (core/train/pretrain_fwd_model.py:24-158), with the loop body (:68-92) replaced by
one call per batch into the sm_100a surrogate-training step (pigan_b200.fwd_trainer.ForwardTrainer).

Kept from the reference: Adam(lr) with torch's default betas, CosineAnnealingLR(T_max=num_epochs,
eta_min=0.01*lr) stepped per epoch (:43-47,135); F in train mode (Dropout 0.2, :53); loss = MSE(spectrum) +
MSE(metrics) (:80-84); clip_grad_norm_(1.0) (:90); epoch averages over len(dataloader) (:128-131);
``forward_model_pretrained.pth`` = bare state_dict and ``fwd_pretrain_loss_history.pt`` = {'train_losses': [...]}
(:146-156).  Changed on purpose: Dropout masks come from a counter-based generator (not torch's RNG stream) and the
losses are read back once per log interval / epoch instead of three ``.item()`` calls per batch.
"""
import argparse
import math
import os
import sys
import time

import torch

project_root = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
if project_root not in sys.path:
    sys.path.append(project_root)
repo_root = os.path.dirname(project_root)
if repo_root not in sys.path:
    sys.path.append(repo_root)

import config.config as cfg
from core.models.forward_model import ForwardModel
from core.utils.data_loader import MetamaterialDataset
from core.utils.set_seed import set_seed

from pigan_b200.fwd_trainer import ForwardTrainer


def cosine_lr(epoch: int, num_epochs: int, lr0: float) -> float:
    """CosineAnnealingLR(T_max=num_epochs, eta_min=0.01*lr0) in closed form (pretrain_fwd_model.py:46)."""
    eta_min = lr0 * 0.01
    return eta_min + (lr0 - eta_min) * (1.0 + math.cos(math.pi * epoch / max(1, num_epochs))) / 2.0


def pretrain_forward_model(forward_model, dataloader, device, num_epochs: int, lr: float, log_interval: int = 10):
    print("\n--- pre-training the forward model (B200-native step) ---")
    device = torch.device(device)
    max_batch = getattr(dataloader, "batch_size", None) or 0
    if not max_batch:
        max_batch = next(iter(dataloader))[0].size(0)
    trainer = ForwardTrainer(forward_model, device, max_batch=max_batch, seed=getattr(cfg, "RANDOM_SEED", 0))
    forward_model.train()
    epoch_losses = []
    rows_seen = 0
    for epoch in range(num_epochs):
        cur_lr = cosine_lr(epoch, num_epochs, lr)
        acc = torch.zeros(3, device=device, dtype=torch.float32)
        n_batches = len(dataloader)
        start = time.time()
        print(f"\nEpoch {epoch + 1}/{num_epochs}")
        for i, (real_spectrum, _, real_params_norm, _, real_metrics_norm) in enumerate(dataloader):
            spec = real_spectrum.to(device, non_blocking=True).float().contiguous()
            pn = real_params_norm.to(device, non_blocking=True).float().contiguous()
            mn = real_metrics_norm.to(device, non_blocking=True).float().contiguous()
            acc += trainer.step(pn, spec, mn, cur_lr, first_row=rows_seen)
            rows_seen += pn.shape[0]
            if (i + 1) % log_interval == 0:
                l, ls, lm = (acc / (i + 1)).tolist()
                print(f"\rProgress: {i + 1}/{n_batches} | Loss:{l:.4f} Spec:{ls:.4f} Metrics:{lm:.4f}", end="",
                      flush=True)
        avg, avg_s, avg_m = (acc / max(1, n_batches)).tolist()
        epoch_losses.append(avg)
        print(f"\rEpoch [{epoch + 1}/{num_epochs}] Summary - loss: {avg:.4f}, spectrum: {avg_s:.4f}, "
              f"metrics: {avg_m:.4f}, lr: {cosine_lr(epoch + 1, num_epochs, lr):.2e}, {time.time() - start:.0f}s")
    os.makedirs(cfg.SAVED_MODELS_DIR, exist_ok=True)
    path = os.path.join(cfg.SAVED_MODELS_DIR, "forward_model_pretrained.pth")
    torch.save(forward_model.state_dict(), path)
    print(f"\npre-trained forward model saved to {path}")
    hist = os.path.join(cfg.SAVED_MODELS_DIR, "fwd_pretrain_loss_history.pt")
    torch.save({"train_losses": epoch_losses}, hist)
    print(f"loss history saved to {hist}")
    return epoch_losses


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="pre-train the forward model")
    parser.add_argument("--epochs", type=int, default=cfg.FWD_PRETRAIN_EPOCHS)
    parser.add_argument("--lr", type=float, default=cfg.FWD_PRETRAIN_LR)
    parser.add_argument("--batch_size", type=int, default=cfg.BATCH_SIZE)
    parser.add_argument("--log_interval", type=int, default=10)
    args = parser.parse_args()
    if not torch.cuda.is_available():
        print("error: the B200 path needs a CUDA device (no CPU fallback)")
        sys.exit(1)
    device = torch.device("cuda")
    set_seed(cfg.RANDOM_SEED)
    cfg.create_directories()
    data_path = cfg.FULL_DATA_PATH
    if not os.path.exists(data_path):
        print(f"error: dataset not found at {data_path}")
        sys.exit(1)
    dataset = MetamaterialDataset(data_path=data_path, num_points_per_sample=cfg.SPECTRUM_DIM)
    dataloader = torch.utils.data.DataLoader(dataset, batch_size=args.batch_size, shuffle=True,
                                             num_workers=cfg.NUM_WORKERS, pin_memory=True)
    forward_model = ForwardModel(cfg.FORWARD_MODEL_INPUT_DIM, cfg.FORWARD_MODEL_OUTPUT_SPEC_DIM,
                                 cfg.FORWARD_MODEL_OUTPUT_METRICS_DIM).to(device)
    pretrain_forward_model(forward_model, dataloader, device, args.epochs, args.lr, args.log_interval)

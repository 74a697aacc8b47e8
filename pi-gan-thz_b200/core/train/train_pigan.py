"""Drop-in ``core.train.train_pigan`` — same ``train_pigan(...)`` signature, loss-history keys, checkpoint and
final-model files as the reference (core/train/train_pigan.py:34-311), with the inner loop (:114-187) replaced by
one call per batch into the fused sm_100a step (pigan_b200.trainer.NativeTrainer).

Kept from the reference: Adam(lr=cfg.LR_G/LR_D, betas=(0.5,0.999)) with CosineAnnealingLR (G) / StepLR (D) stepped
per epoch (:56-62,252-253); G and D in train mode, F in eval mode (:73-75); epoch averages over len(dataloader)
(:236-244); checkpoints every cfg.SAVE_MODEL_INTERVAL epochs (:284-295); final state_dicts + loss history (:299-309).
Changed on purpose: losses stay on the device and are read back once per log interval / epoch instead of 12
``.item()`` synchronisations per step.
"""
import argparse
import os
import sys
import time

import torch
import torch.optim as optim
from torch.optim.lr_scheduler import CosineAnnealingLR, StepLR
from torch.utils.data import DataLoader

project_root = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
if project_root not in sys.path:
    sys.path.append(project_root)
repo_root = os.path.dirname(project_root)
if repo_root not in sys.path:
    sys.path.append(repo_root)

import config.config as cfg
from core.models.discriminator import Discriminator
from core.models.forward_model import ForwardModel
from core.models.generator import Generator
from core.utils.data_loader import MetamaterialDataset
from core.utils.set_seed import set_seed

from pigan_b200.trainer import LOSS_KEYS, NativeTrainer


def _is_rank0() -> bool:
    d = torch.distributed
    return not (d.is_available() and d.is_initialized()) or d.get_rank() == 0


def train_pigan(dataloader, device, generator, discriminator, forward_model, dataset, num_epochs: int,
                log_interval: int = 10):
    print("\n--- starting PI-GAN training (B200-native step) ---")
    device = torch.device(device)
    optimizer_g = optim.Adam(generator.parameters(), lr=cfg.LR_G, betas=(0.5, 0.999))
    optimizer_d = optim.Adam(discriminator.parameters(), lr=cfg.LR_D, betas=(0.5, 0.999))
    scheduler_g = CosineAnnealingLR(optimizer_g, T_max=num_epochs, eta_min=cfg.LR_G * 0.01)
    scheduler_d = StepLR(optimizer_d, step_size=max(1, num_epochs // 4), gamma=0.5)

    generator.train()
    discriminator.train()
    forward_model.eval()

    max_batch = getattr(dataloader, "batch_size", None) or 0
    if not max_batch:
        first = next(iter(dataloader))
        max_batch = first[0].size(0)
    trainer = NativeTrainer(generator, discriminator, forward_model, device, max_batch=max_batch, cfg=cfg,
                            f1_idx=dataset.metric_name_to_idx["f1"], f2_idx=dataset.metric_name_to_idx["f2"])
    # parameters were re-pointed at flat buffers: rebuild the optimisers' parameter lists on the same objects
    optimizer_g.param_groups[0]["params"] = list(generator.parameters())
    optimizer_d.param_groups[0]["params"] = list(discriminator.parameters())

    loss_history = {k: [] for k in LOSS_KEYS}
    for epoch in range(num_epochs):
        total_batches = len(dataloader)
        print(f"\nEpoch {epoch + 1}/{num_epochs}")
        start_time = time.time()
        epoch_sum = torch.zeros(9, device=device, dtype=torch.float64)
        window_sum = torch.zeros(9, device=device, dtype=torch.float64)
        lr_g = optimizer_g.param_groups[0]["lr"]
        lr_d = optimizer_d.param_groups[0]["lr"]
        for i, (real_spectrum, real_params_denorm, _real_params_norm, _real_metrics_denorm,
                real_metrics_norm) in enumerate(dataloader):
            spec = real_spectrum.to(device, torch.float32, non_blocking=True).contiguous()
            pden = real_params_denorm.to(device, torch.float32, non_blocking=True).contiguous()
            mnorm = real_metrics_norm.to(device, torch.float32, non_blocking=True).contiguous()
            losses = trainer.step(spec, pden, mnorm, lr_g, lr_d)
            epoch_sum += losses
            window_sum += losses
            if (i + 1) % log_interval == 0:
                w = (window_sum / log_interval).tolist()  # the only host sync of the window
                progress = (i + 1) / total_batches
                bar = "█" * int(50 * progress) + "-" * (50 - int(50 * progress))
                eta = (time.time() - start_time) / progress * (1 - progress)
                print(f"\rProgress: [{bar}] {i + 1}/{total_batches} | D:{w[0]:.4f} G:{w[1]:.4f} A:{w[2]:.4f} | "
                      f"ETA: {eta:.0f}s", end="", flush=True)
                window_sum.zero_()
        avg = (epoch_sum / max(1, len(dataloader))).tolist()
        print(f"\rProgress: [{'█' * 50}] {total_batches}/{total_batches} | D:{avg[0]:.4f} G:{avg[1]:.4f} | "
              f"Completed in {time.time() - start_time:.0f}s")
        scheduler_g.step()
        scheduler_d.step()
        if (epoch + 1) % cfg.LOG_INTERVAL == 0:
            print(f"\nEpoch [{epoch + 1}/{num_epochs}] Summary:")
            print(f"  D_Loss: {avg[0]:.4f}, G_Loss: {avg[1]:.4f}")
            print(f"  Learning Rates - G: {optimizer_g.param_groups[0]['lr']:.2e}, "
                  f"D: {optimizer_d.param_groups[0]['lr']:.2e}")
            print(f"  G_SubLosses - Adv: {avg[2]:.4f}, Recon_Spec: {avg[3]:.4f}, Recon_Metrics: {avg[4]:.4f}")
            print(f"  Physics_Losses - Maxwell: {avg[5]:.4f}, LC: {avg[6]:.4f}, ParamRange: {avg[7]:.4f}, "
                  f"BNN_KL: {avg[8]:.4f}")
        for k, v in zip(LOSS_KEYS, avg):
            loss_history[k].append(v)

        if (epoch + 1) % cfg.SAVE_MODEL_INTERVAL == 0 and _is_rank0():   # replicas are identical: rank 0 writes
            os.makedirs(cfg.CHECKPOINT_DIR, exist_ok=True)
            trainer.export_optimizer_state(optimizer_g, optimizer_d)
            path = os.path.join(cfg.CHECKPOINT_DIR, f"pigan_epoch_{epoch + 1}.pth")
            torch.save({
                "epoch": epoch + 1,
                "generator_state_dict": generator.state_dict(),
                "discriminator_state_dict": discriminator.state_dict(),
                "forward_model_state_dict": forward_model.state_dict(),
                "optimizer_g_state_dict": optimizer_g.state_dict(),
                "optimizer_d_state_dict": optimizer_d.state_dict(),
            }, path)
            print(f"checkpoint saved to {path}")

    print("--- PI-GAN training finished ---")
    if _is_rank0():
        os.makedirs(cfg.SAVED_MODELS_DIR, exist_ok=True)
        torch.save(generator.state_dict(), os.path.join(cfg.SAVED_MODELS_DIR, "generator_final.pth"))
        torch.save(discriminator.state_dict(), os.path.join(cfg.SAVED_MODELS_DIR, "discriminator_final.pth"))
        torch.save(forward_model.state_dict(), os.path.join(cfg.SAVED_MODELS_DIR, "forward_model_final.pth"))
        torch.save(loss_history, os.path.join(cfg.SAVED_MODELS_DIR, "pigan_loss_history.pt"))
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.barrier()      # nobody returns (and reads the files) before rank 0 has written them
    return loss_history


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="Train the PI-GAN model (B200-native step).")
    parser.add_argument("--epochs", type=int, default=cfg.NUM_EPOCHS)
    parser.add_argument("--batch_size", type=int, default=cfg.BATCH_SIZE)
    parser.add_argument("--lr_g", type=float, default=cfg.LR_G)   # parsed and ignored, as in the reference (:320-323)
    parser.add_argument("--lr_d", type=float, default=cfg.LR_D)
    parser.add_argument("--fwd_model_path", type=str,
                        default=os.path.join(cfg.SAVED_MODELS_DIR, "forward_model_pretrained.pth"))
    parser.add_argument("--log_interval", type=int, default=10)
    args = parser.parse_args()

    cfg.create_directories()
    device = torch.device(cfg.DEVICE)
    set_seed(cfg.RANDOM_SEED)
    if not os.path.exists(cfg.DATASET_PATH):
        print(f"error: dataset not found at {cfg.DATASET_PATH}")
        sys.exit(1)
    dataset = MetamaterialDataset(data_path=cfg.DATASET_PATH, num_points_per_sample=cfg.SPECTRUM_DIM)
    loader = DataLoader(dataset, batch_size=args.batch_size, shuffle=True, num_workers=cfg.NUM_WORKERS,
                        pin_memory=True)
    generator = Generator(cfg.GENERATOR_INPUT_DIM, cfg.GENERATOR_OUTPUT_PARAM_DIM).to(device)
    discriminator = Discriminator(cfg.DISCRIMINATOR_INPUT_SPEC_DIM, cfg.DISCRIMINATOR_INPUT_PARAM_DIM).to(device)
    forward_model = ForwardModel(cfg.FORWARD_MODEL_INPUT_DIM, cfg.FORWARD_MODEL_OUTPUT_SPEC_DIM,
                                 cfg.FORWARD_MODEL_OUTPUT_METRICS_DIM).to(device)
    if not os.path.exists(args.fwd_model_path):
        print(f"error: pretrained forward model not found at {args.fwd_model_path}")
        sys.exit(1)
    forward_model.load_state_dict(torch.load(args.fwd_model_path, map_location=device))
    train_pigan(loader, device, generator, discriminator, forward_model, dataset, args.epochs, args.log_interval)

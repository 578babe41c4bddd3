"""Multi-GPU plumbing of the path: driver frames are independent once the source identity has been
encoded, so the only exchange step is ONE broadcast per identity of the cached source state

    target_latent_volume  1 x 16 x 64 x 64 x 96 fp32 = 25.2 MB
    idt_embed             1 x 512 x 4 x 4 fp32      = 32 KB
    source theta          4 x 4 fp32                = 64 B

from the rank that ran the source pass (torch.distributed / NCCL over NVLink; `gloo` in the CPU tests).  After it every
rank runs its own frames; there is no data-path collective (frames are sharded round-robin, weak scaling).
The reference's own num_gpus > 1 inference branch (notebooks/infer.py:99-136) is not usable as a model for this.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.distributed as dist

from .config import HotPathConfig


def source_state_shapes(cfg: HotPathConfig):
    return {
        "target_latent_volume": (1, cfg.D, cfg.S, cfg.S, cfg.C),
        "idt_embed": (1, cfg.idt_channels, cfg.embed_size, cfg.embed_size),
        "pred_source_theta": (1, 4, 4),
    }


def broadcast_source_state(st, cfg: HotPathConfig, device, src: int = 0, group=None):
    """Rank `src` passes its source state; every rank returns a state usable by Model.driver_pass."""
    rank = dist.get_rank(group)
    shapes = source_state_shapes(cfg)
    # one flat buffer -> one collective
    total = sum(int(torch.tensor(s).prod()) for s in shapes.values())
    flat = torch.empty(total, dtype=torch.float32, device=device)
    if rank == src:
        off = 0
        for k, shp in shapes.items():
            t = getattr(st, k).reshape(-1)
            flat[off:off + t.numel()].copy_(t)
            off += t.numel()
    dist.broadcast(flat, src=src, group=group)
    out = SimpleNamespace()
    off = 0
    for k, shp in shapes.items():
        n = int(torch.tensor(shp).prod())
        setattr(out, k, flat[off:off + n].view(shp).clone().contiguous())
        off += n
    out.source_theta_dev = out.pred_source_theta[0].contiguous()
    return out


def shard_frames(num_frames: int, rank: int, world: int):
    """frame indices rank `rank` processes: i == rank (mod world)"""
    return list(range(rank, num_frames, world))

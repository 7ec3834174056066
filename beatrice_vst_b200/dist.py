"""Multi-GPU plumbing: one process per GPU, streams sharded as independent partitions.

The per-hop path has NO collective (voice streams are independent).
``torch.distributed`` is used once at load: rank 0 reads the five model files and broadcasts
their images (NCCL over NVLink on GPUs, gloo in the CPU tests); every rank then hands the same
bytes to ``BeatriceB200_LoadModelFromMemory``.
"""
from __future__ import annotations

import os

import numpy as np

MODEL_FILES = ["phone_extractor.bin", "pitch_estimator.bin", "waveform_generator.bin",
               "embedding_setter.bin", "speaker_embeddings.bin"]


def shard_streams(n_total: int, world: int, rank: int):
    """Contiguous blocks: stream -> rank = id // ceil(n_total / world).  Returns (first, count)."""
    per = (n_total + world - 1) // world
    first = min(rank * per, n_total)
    return first, max(0, min(per, n_total - first))


def read_model_images(model_dir: str):
    return [np.fromfile(os.path.join(model_dir, f), dtype=np.uint8) for f in MODEL_FILES]


def broadcast_model_images(images, src: int = 0, device=None):
    """Broadcasts the five file images from ``src``; returns them as numpy uint8 arrays."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    device = device or torch.device("cpu")
    sizes = torch.zeros(len(MODEL_FILES), dtype=torch.int64, device=device)
    if rank == src:
        sizes = torch.tensor([im.size for im in images], dtype=torch.int64, device=device)
    dist.broadcast(sizes, src=src)
    total = int(sizes.sum().item())
    blob = torch.empty(total, dtype=torch.uint8, device=device)
    if rank == src:
        blob.copy_(torch.from_numpy(np.concatenate(images)))
    dist.broadcast(blob, src=src)
    host = blob.cpu().numpy()
    out, pos = [], 0
    for n in sizes.tolist():
        out.append(host[pos:pos + n].copy())
        pos += n
    return out

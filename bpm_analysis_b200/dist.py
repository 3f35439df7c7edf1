"""Multi-GPU partitioning: one process per GPU, recordings are the unit of work.

The front end shards naturally over independent recordings (BASELINE configs[2], the C5
parameter sweep, and bench.py's one-recording-per-GPU runs): each rank owns a contiguous
block of the batch and runs the whole hot path on it with no data-path collective.  The
only communication is an optional gather of small per-recording summaries (counts, mean
BPM, ...) to rank 0 for reporting, over ``torch.distributed`` (NCCL on GPUs, gloo in the
CPU tests).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank ``rank``; sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_by_cost(costs: Sequence[float], world: int) -> List[List[int]]:
    """Longest-processing-time assignment of ragged recordings (cost = raw samples) to ranks."""
    order = np.argsort(-np.asarray(costs, dtype=np.float64), kind="stable")
    load = np.zeros(world)
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(load))
        out[r].append(int(i))
        load[r] += costs[i]
    return [sorted(v) for v in out]


def gather_summaries(local: np.ndarray, device: torch.device) -> np.ndarray:
    """All ranks contribute a (n_local, k) float64 table; every rank gets the concatenation in
    rank order.  Row counts may differ per rank (padded all_gather + counts)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    local = np.ascontiguousarray(local, dtype=np.float64)
    if local.ndim != 2:
        raise ValueError("summary table must be 2-D")
    if world == 1:
        return local.copy()
    k = local.shape[1]
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=device)
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(counts, n_local)
    counts = [int(c.item()) for c in counts]
    cap = max(counts) if counts else 0
    buf = torch.zeros((cap, k), dtype=torch.float64, device=device)
    buf[:local.shape[0]] = torch.from_numpy(local).to(device)
    parts = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    return np.concatenate([p[:c].cpu().numpy() for p, c in zip(parts, counts)], axis=0)


def summarize(result_item: Dict[str, object], duration_sec: float) -> np.ndarray:
    """Per-recording summary row: [duration_s, n_troughs, n_peaks, mean envelope, mean floor]."""
    return np.array([duration_sec, len(result_item["troughs"]), len(result_item["peaks"]),
                     float(np.mean(result_item["envelope"])), float(np.mean(result_item["floor"]))])


def bind_to_gpu_numa(device_index: int) -> bool:
    """Pin the calling process to the CPU cores nearest to GPU ``device_index`` (NVML's ideal CPU
    affinity) BEFORE it allocates pinned host buffers: the zero-copy ingest reads those buffers over
    PCIe, and pages that sit on the other socket cross the inter-socket link as well.  Returns
    False (and changes nothing) when NVML or sched_setaffinity is unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:                                   # noqa: BLE001 - an optimisation only
        return False

"""Data-parallel gradient synchronisation of the training step: one process per GPU, weights
replicated, gradients averaged over ranks after loss.backward().

Under Lightning's DDP strategy (the reference's launcher, configs/trainer/default.yaml) torch's DistributedDataParallel
wrapper does this through per-parameter hooks and nothing here is needed: the UNet's autograd node hands ordinary
gradient tensors to autograd.  For a plain loop (bench / tools) `sync_gradients` does the same with the least traffic
the layout allows: the UNet engine writes all 316 parameter gradients into ONE flat fp32 buffer, so they travel as a
single all-reduce (NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests); the remaining parameters (the condition
encoder) are coalesced into one more.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def _flat_groups(params: Iterable[torch.nn.Parameter]) -> List[List[torch.Tensor]]:
    """Group gradients that are consecutive views of one storage (the engine's flat buffer) so each group can be
    all-reduced in place as a single tensor; everything else ends up in single-tensor groups."""
    groups: List[List[torch.Tensor]] = []
    by_storage = {}
    for p in params:
        g = p.grad
        if g is None:
            continue
        key = (g.untyped_storage().data_ptr(), g.dtype, g.device)
        by_storage.setdefault(key, []).append(g)
    for gs in by_storage.values():
        groups.append(gs)
    return groups


def sync_gradients(module: torch.nn.Module, group: Optional[dist.ProcessGroup] = None) -> int:
    """Average .grad over the ranks of `group`.  Returns the number of collectives issued (2 for a cDDPM model)."""
    if not dist.is_available() or not dist.is_initialized():
        return 0
    world = dist.get_world_size(group)
    if world == 1:
        return 0
    calls = 0
    loose: List[torch.Tensor] = []
    for gs in _flat_groups(module.parameters()):
        if len(gs) > 1:
            # views of one buffer: reduce the covering span in place
            st = gs[0].untyped_storage()
            lo = min(g.storage_offset() for g in gs)
            hi = max(g.storage_offset() + g.numel() for g in gs)
            span = torch.empty(0, dtype=gs[0].dtype, device=gs[0].device).set_(st, lo, (hi - lo,))
            dist.all_reduce(span, group=group)
            span.div_(world)
            calls += 1
        else:
            loose.append(gs[0])
    if loose:
        flat = torch._utils._flatten_dense_tensors(loose)
        dist.all_reduce(flat, group=group)
        flat.div_(world)
        for g, f in zip(loose, torch._utils._unflatten_dense_tensors(flat, loose)):
            g.copy_(f)
        calls += 1
    return calls

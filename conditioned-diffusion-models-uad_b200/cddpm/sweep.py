"""Test-sweep driver — the part of the reference's train.py that evaluates a trained model (train.py:182-237:
`trainer.test` over the validation then the test loader of every test set, `utils.summarize`, `preds_dict.pkl`),
without a Lightning Trainer and sharded over ranks (BASELINE configs[3]).

One process per GPU; volumes are dealt round-robin to ranks; every rank runs `on_test_start / test_step / on_test_end` of
the LightningModule drop-in on its volumes.  Collectives: the per-volume result lists are all-gathered once per stage
(restored to the loader's order), and the validation stage's global Dice threshold is computed inside `_test_end` from
all-reduced counts — so a sharded sweep reports what the single-process sweep reports.  (The healthy-set FPR thresholds
of `_test_end` use the rank's own voxels.)  The val -> test hand-off of `threshold['total']` lives in the module, as in
the reference.
"""
from __future__ import annotations

import copy
import os
import pickle
from typing import Any, Dict, Iterable, Mapping, Optional, Tuple

import torch
import torch.distributed as dist


def _world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def summarize(eval_dict: Mapping[str, Any], prefix: str) -> Dict[str, Any]:
    """utils.summarize (utils.py:172-178): prefix the keys, drop list entries."""
    return {prefix + "/" + k: v for k, v in eval_dict.items() if type(v) is not list}


def _to_device(obj, device):
    if torch.is_tensor(obj):
        return obj.to(device, non_blocking=True)
    if isinstance(obj, dict):
        return {k: _to_device(v, device) for k, v in obj.items()}
    return obj


def merge_lists(local: Mapping[str, Any], n_local: int, n_total: int) -> Dict[str, list]:
    """All-gather every list-valued entry of a rank-local eval_dict.  Lists with one entry per volume come back in the
    loader's order (volume i was evaluated by rank i % world); other lists are concatenated rank by rank."""
    rank, world = _world()
    lists = {k: v for k, v in local.items() if type(v) is list}
    if world == 1:
        return dict(lists)
    payload = {k: [x.cpu() if torch.is_tensor(x) else x for x in v] for k, v in lists.items()}
    gathered = [None] * world
    dist.all_gather_object(gathered, (n_local, payload))
    out: Dict[str, list] = {}
    for k in lists:
        per_rank = [g[1].get(k, []) for g in gathered]
        if all(len(per_rank[r]) == gathered[r][0] for r in range(world)):
            merged = [None] * n_total
            for r in range(world):
                for j, item in enumerate(per_rank[r]):
                    merged[r + j * world] = item
            out[k] = merged
        else:
            out[k] = [item for pr in per_rank for item in pr]
    return out


def run_stage(model, batches: Iterable[Mapping[str, Any]], device: Optional[torch.device] = None) -> Dict[str, Any]:
    """`trainer.test(model, dataloader)` for one stage: returns the module's eval_dict after on_test_end."""
    rank, world = _world()
    if device is None:
        device = next(model.parameters()).device
    model.eval()
    model.on_test_start()
    batches = list(batches)
    mine = list(enumerate(batches))[rank::world]
    if hasattr(model, "test_step_reconstruct") and device.type == "cuda" and os.environ.get("CDDPM_SWEEP_PIPELINE", "1") != "0":
        # Software pipeline over volumes: the reconstruction of volume i+1 (pure enqueue, ~26 ms of GPU work for 50
        # slices) goes onto the main stream BEFORE volume i is scored; the scoring tail, whose threshold bisection reads
        # a few counters back per step, runs on a high-priority side stream, so its host round trips no longer idle the
        # GPU.  Results and their order are those of the plain loop (the host RNG draws all happen in the first half).
        with torch.cuda.device(device):
            tail = getattr(model, "_tail_stream", None)
            if tail is None:
                tail = model._tail_stream = torch.cuda.Stream(device=device, priority=-1)
            pending = None
            for idx, batch in mine + [(None, None)]:
                nxt = (idx, model.test_step_reconstruct(_to_device(batch, device))) if batch is not None else None
                if pending is not None:
                    with torch.cuda.stream(tail):
                        model.test_step_finish(pending[1], pending[0])
                    tail.synchronize()  # volume i's tensors may be recycled by the main stream from here on
                pending = nxt
    else:
        for idx, batch in mine:
            model.test_step(_to_device(batch, device), idx)
    if world > 1:
        model.eval_dict.update(merge_lists(model.eval_dict, len(mine), len(batches)))
    model.on_test_end()
    return model.eval_dict


def test_sweep(model, testsets: Mapping[str, Tuple[Iterable, Iterable]], fold: int = 0, log_dir: Optional[str] = None,
               pickle_preds: bool = True):
    """train.py:182-237: for every test set, the validation loader (finds thresholds), then the test loader.
    `testsets` maps the set name (e.g. 'Datamodules_eval.Brats21') to (val_batches, test_batches).
    Returns (preds_dict, log_dict); rank 0 pickles preds_dict to `<log_dir>/<fold+1>_preds_dict.pkl`."""
    rank, _ = _world()
    preds: Dict[str, Dict[str, Any]] = {"val": {}, "test": {}}
    logs: Dict[str, Any] = {}
    for name, (val_batches, test_batches) in testsets.items():
        preds["val"][name] = copy.copy(run_stage(model, val_batches))
        log = summarize(preds["val"][name], "val")
        preds["test"][name] = copy.copy(run_stage(model, test_batches))
        log.update(summarize(preds["test"][name], "test"))
        logs.update(summarize(log, f"{fold + 1}/" + name))
    if pickle_preds and log_dir is not None and rank == 0:
        os.makedirs(log_dir, exist_ok=True)
        with open(os.path.join(log_dir, f"{fold + 1}_preds_dict.pkl"), "wb") as f:
            pickle.dump(preds, f)
    return preds, logs


test_sweep.__test__ = False  # not a pytest test

"""Adam for the training step — DDPM_2D.configure_optimizers (DDPM_2D.py:305-306: `optim.Adam(self.parameters(), lr)`).

Same update rule and state layout as torch.optim.Adam (state[p] = {step, exp_avg, exp_avg_sq}; weight_decay 0, no
amsgrad), executed as ONE kernel launch over all parameter tensors (`cddpm_adam_step`): a device table of
(param, grad, exp_avg, exp_avg_sq) pointers is refreshed per step (gradient tensors are new every backward) and each
block updates one 4096-element chunk with 16-byte accesses.  67.6 M parameters move 1.9 GB per step; torch's fused
multi-tensor Adam needs 2.2 ms for them on B200, this kernel is bandwidth-bound.
"""
from __future__ import annotations

from typing import Iterable

import torch

from ._lib import CddpmError, check, current_stream, lib, ptr

_CHUNK = 4096


class Adam(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._tables = {}

    _RING = 4  # pinned gradient-pointer tables in flight (one CUDA event each)

    def _init_group(self, gi, group):
        ps = [p for p in group["params"] if p.requires_grad]
        for p in ps:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise CddpmError("cddpm.optim.Adam updates contiguous fp32 CUDA parameters (there is no CPU path)")
        dev = ps[0].device
        for p in ps:
            st = self.state[p]
            if "exp_avg" not in st:
                st["step"] = torch.zeros((), dtype=torch.float32)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            elif not st["exp_avg"].is_cuda or not st["exp_avg"].is_contiguous() or \
                    not st["exp_avg_sq"].is_contiguous():  # e.g. a checkpoint loaded with map_location="cpu"
                st["exp_avg"] = st["exp_avg"].to(dev, torch.float32).contiguous()
                st["exp_avg_sq"] = st["exp_avg_sq"].to(dev, torch.float32).contiguous()
        numel = [p.numel() for p in ps]
        block_tensor, block_off = [], []
        for t, n in enumerate(numel):
            for off in range(0, n, _CHUNK):
                block_tensor.append(t)
                block_off.append(off)
        i64 = dict(dtype=torch.int64, device=dev)
        tab = {
            "params": ps,
            "p": torch.tensor([p.data_ptr() for p in ps], **i64),
            "m": torch.tensor([self.state[p]["exp_avg"].data_ptr() for p in ps], **i64),
            "v": torch.tensor([self.state[p]["exp_avg_sq"].data_ptr() for p in ps], **i64),
            "numel": torch.tensor(numel, **i64),
            "block_tensor": torch.tensor(block_tensor, dtype=torch.int32, device=dev),
            "block_off": torch.tensor(block_off, **i64),
            # gradient tensors are new every backward: their pointer table is refreshed per launch through a ring of
            # pinned host buffers + device copies, each guarded by an event recorded after the kernel that reads it, so a
            # host running ahead of the GPU never overwrites a table a pending copy / kernel still needs
            "g_host": [torch.zeros(len(ps), dtype=torch.int64).pin_memory() for _ in range(self._RING)],
            "g": [torch.zeros(len(ps), **i64) for _ in range(self._RING)],
            "g_event": [None] * self._RING,
            "slot": 0,
            "key": self._key(ps),
        }
        self._tables[gi] = tab
        return tab

    def _key(self, ps):
        """Everything the cached device tables point at: parameter storages AND the moment buffers (load_state_dict
        replaces the latter)."""
        key = []
        for p in ps:
            st = self.state.get(p, {})
            m, v = st.get("exp_avg"), st.get("exp_avg_sq")
            key.append((p.data_ptr(), m.data_ptr() if m is not None else 0, v.data_ptr() if v is not None else 0))
        return key

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._tables = {}  # the moment tensors were replaced: every cached pointer table is stale

    def _launch(self, tab, group, ptrs, step):
        slot = tab["slot"]
        tab["slot"] = (slot + 1) % self._RING
        ev = tab["g_event"][slot]
        if ev is not None:
            ev.synchronize()
        tab["g_host"][slot].copy_(torch.tensor(ptrs, dtype=torch.int64))
        tab["g"][slot].copy_(tab["g_host"][slot], non_blocking=True)
        b1, b2 = group["betas"]
        bc1 = 1.0 - b1 ** step
        bc2 = 1.0 - b2 ** step
        check(lib().cddpm_adam_step(ptr(tab["p"]), ptr(tab["g"][slot]), ptr(tab["m"]), ptr(tab["v"]), ptr(tab["numel"]),
                                    ptr(tab["block_tensor"]), ptr(tab["block_off"]), tab["block_tensor"].numel(),
                                    float(group["lr"]), float(b1), float(b2), float(group["eps"]), bc1, bc2,
                                    current_stream()), "cddpm_adam_step")
        if ev is None:
            ev = tab["g_event"][slot] = torch.cuda.Event()
        ev.record()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            tab = self._tables.get(gi)
            if tab is None or len(tab["params"]) != sum(1 for p in group["params"] if p.requires_grad) or \
                    tab["key"] != self._key(tab["params"]):
                tab = self._init_group(gi, group)  # first step, .to()/.cuda() re-seated storages, or load_state_dict
            ps = tab["params"]
            keep = []
            ptrs = []
            for p in ps:
                g = p.grad
                if g is None:
                    ptrs.append(0)
                    continue
                if g.dtype != torch.float32 or not g.is_contiguous():
                    g = g.float().contiguous()
                    keep.append(g)
                ptrs.append(g.data_ptr())
            # like torch.optim.Adam, only parameters that received a gradient advance their step counter; tensors whose
            # counters differ (a parameter that sat out some steps) are updated by separate launches of the same table
            active = [i for i, a in enumerate(ptrs) if a]
            if not active:
                continue
            steps = [self.state[ps[i]]["step"] for i in active]
            torch._foreach_add_(steps, 1.0)
            vals = [float(s) for s in steps]
            if all(v == vals[0] for v in vals):
                self._launch(tab, group, ptrs, vals[0])
            else:
                for sv in sorted(set(vals)):
                    members = {i for i, v in zip(active, vals) if v == sv}
                    self._launch(tab, group, [a if i in members else 0 for i, a in enumerate(ptrs)], sv)
            _bump_versions([ps[i] for i in active])  # in-place update through raw pointers: tell the engines
            if keep:  # converted copies must outlive the kernel that reads them
                torch.cuda.current_stream().synchronize()
            del keep
        return loss


def _bump_versions(params):
    """The engines re-pack a parameter when its (data_ptr, _version) changes; the kernel wrote through raw pointers,
    so advance the version counters the way an in-place torch op would."""
    params = list(params)
    setter = getattr(torch._C._autograd, "_unsafe_set_version_counter", None)
    if setter is not None:
        setter(params, [p._version + 1 for p in params])
    else:  # older torch: a zero add bumps the counters (one multi-tensor launch)
        torch._foreach_add_(params, 0.0)

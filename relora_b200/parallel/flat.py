"""Flat parameter / gradient / optimizer-state storage and the AdamW that runs on it.

Every trainable tensor lives at a fixed offset of ONE contiguous buffer per role
(``params``, ``grads``, ``exp_avg``, ``exp_avg_sq``); ``nn.Parameter.data`` / ``.grad`` are views.
That turns the per-update work of data-parallel training into a handful of whole-buffer kernels:

* gradient all-reduce  = one collective (NCCL baseline) or one peer-memory kernel over ``grads``;
* clipping             = one sum-of-squares over ``grads`` (``clip_grad_norm_`` semantics,
                         reference ``torchrun_main.py:805-808``);
* AdamW                = one fused kernel over the four buffers (``torchrun_main.py:666, 814``);
* ZeRO-1 (``adam_zero``, reference ``:668-675``) = reduce-scatter → update own shard → all-gather,
  on the same buffers, with shard boundaries aligned so TMA / 128-bit accesses stay aligned.

:class:`FlatAdamW` subclasses ``torch.optim.Optimizer`` and exposes ordinary per-parameter state
(views), so ``state_dict()`` has exactly the ``torch.optim.AdamW`` layout the reference writes to
``optimizer.pt`` and LR schedulers attach to it unchanged.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from ..ops import reference as ref

__all__ = ["FlatParamStore", "FlatAdamW"]

_ALIGN = 128  # elements; 256 B for bf16 — keeps every tensor 16-byte (TMA / v4) aligned


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class FlatParamStore:
    """Re-homes ``params`` into one flat buffer and gives each a gradient view in a second one."""

    def __init__(
        self,
        named_params: Sequence[Tuple[str, torch.nn.Parameter]],
        *,
        world_size: int = 1,
        grad_dtype: Optional[torch.dtype] = None,
        allocator: Optional[Callable[[int, torch.dtype, torch.device], torch.Tensor]] = None,
        bind_grads: bool = True,
        grad_allocator: Optional[Callable[[int, torch.dtype, torch.device], torch.Tensor]] = None,
        storage_shapes: Optional[Dict[int, Tuple[int, int]]] = None,
    ):
        if not named_params:
            raise ValueError("no trainable parameters")
        self.names = [n for n, _ in named_params]
        self.param_list = [p for _, p in named_params]
        p0 = self.param_list[0]
        self.dtype, self.device = p0.dtype, p0.device
        for n, p in named_params:
            if p.dtype != self.dtype or p.device != self.device:
                raise ValueError(f"parameter {n} has dtype/device {p.dtype}/{p.device}, expected {self.dtype}/{self.device}")
        # optional padded storage: a 2-D parameter [R, C] may live in a larger [R_s, C_s] block (row stride C_s) so that
        # stacked / TMA-aligned views exist; the padding stays exactly zero (zero gradients, zero moments)
        self.storage: Dict[int, Tuple[int, int]] = dict(storage_shapes or {})
        for p in self.param_list:
            if id(p) in self.storage:
                rs, cs = self.storage[id(p)]
                if p.dim() != 2 or rs < p.shape[0] or cs < p.shape[1]:
                    raise ValueError("storage shape must cover the 2-D parameter")
        self.offsets: List[int] = []
        off = 0
        for p in self.param_list:
            self.offsets.append(off)
            off += _round_up(self._storage_numel(p), _ALIGN)
        self.used = off
        self.numel = _round_up(off, _ALIGN * max(1, world_size))
        alloc = allocator or (lambda n, dt, dev: torch.zeros(n, dtype=dt, device=dev))
        self.params = alloc(self.numel, self.dtype, self.device)
        self.params.zero_()  # alignment gaps and padded blocks must be exactly zero (custom allocators may not clear)
        self.grad_dtype = grad_dtype or self.dtype
        galloc = grad_allocator or (lambda n, dt, dev: torch.zeros(n, dtype=dt, device=dev))
        self.grads = galloc(self.numel, self.grad_dtype, self.device)
        # autograd can only accumulate into .grad of the parameter's own dtype; executors that write
        # gradients themselves (fp32 accumulation for bf16 parameters) keep the views to themselves
        self.bind_grads = bind_grads and self.grad_dtype == self.dtype
        with torch.no_grad():
            self.index: Dict[int, int] = {id(p): i for i, p in enumerate(self.param_list)}
            for p in self.param_list:
                view = self.view_like(self.params, p)
                view.copy_(p.data)
                p.data = view
                if self.bind_grads:
                    p.grad = self.view_like(self.grads, p)

    # ------------------------------------------------------------------ views
    def _storage_numel(self, p: torch.nn.Parameter) -> int:
        st = self.storage.get(id(p))
        return p.numel() if st is None else st[0] * st[1]

    def is_padded(self, p: torch.nn.Parameter) -> bool:
        return id(p) in self.storage

    def segment(self, p: torch.nn.Parameter) -> Tuple[int, int]:
        """(offset, number of *storage* elements) of ``p`` in the flat buffers."""
        i = self.index[id(p)]
        return self.offsets[i], self._storage_numel(self.param_list[i])

    def view_like(self, flat: torch.Tensor, p: torch.nn.Parameter, base: int = 0) -> torch.Tensor:
        """View of ``flat`` shaped like ``p`` (a strided slice of the padded block for padded parameters)."""
        o, n = self.segment(p)
        st = self.storage.get(id(p))
        if st is None:
            return flat[o - base : o - base + n].view(p.shape)
        return flat[o - base : o - base + n].view(st[0], st[1])[: p.shape[0], : p.shape[1]]

    def to_storage(self, p: torch.nn.Parameter, value: torch.Tensor) -> torch.Tensor:
        """``value`` (shaped like ``p``) laid out as the flat storage segment of ``p`` (zero padded)."""
        st = self.storage.get(id(p))
        if st is None:
            return value.reshape(-1)
        full = torch.zeros(st[0], st[1], dtype=value.dtype, device=value.device)
        full[: p.shape[0], : p.shape[1]] = value
        return full.reshape(-1)

    def rebind_grads(self) -> None:
        """Point every ``.grad`` back at its flat view (after ``zero_grad(set_to_none=True)`` etc.)."""
        if not self.bind_grads:
            return
        for p in self.param_list:
            p.grad = self.view_like(self.grads, p)

    def zero_grads(self) -> None:
        self.grads.zero_()

    def shard_bounds(self, rank: int, world_size: int) -> Tuple[int, int]:
        per = self.numel // world_size
        return rank * per, (rank + 1) * per

    def segments_of(self, params: Iterable[torch.nn.Parameter]) -> List[Tuple[int, int]]:
        return [self.segment(p) for p in params if id(p) in self.index]


class FlatAdamW(torch.optim.Optimizer):
    """AdamW over a :class:`FlatParamStore` (optionally owning only one ZeRO-1 shard of the state).

    ``step(grad_scale=…, skip=…)``: ``grad_scale`` (float or 0-dim tensor) multiplies the gradient
    inside the update (1/world averaging and the clip coefficient are folded in here); ``skip``
    (0-dim bool/float tensor or bool) suppresses the update on the device — the NaN guard of the
    reference loop (``torchrun_main.py:813-822``) without a host round trip on the fused path.
    """

    def __init__(
        self,
        store: FlatParamStore,
        *,
        lr: float = 1e-3,
        betas: Tuple[float, float] = (0.9, 0.999),
        eps: float = 1e-8,
        weight_decay: float = 0.0,
        shard: Optional[Tuple[int, int]] = None,
        state_dtype: Optional[torch.dtype] = None,
        native=None,
    ):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False,
                        maximize=False, foreach=None, capturable=False, differentiable=False, fused=None)
        super().__init__(store.param_list, defaults)
        self.store = store
        self.shard = shard or (0, store.numel)
        lo, hi = self.shard
        sd = state_dtype or store.dtype
        self.exp_avg = torch.zeros(hi - lo, dtype=sd, device=store.device)
        self.exp_avg_sq = torch.zeros(hi - lo, dtype=sd, device=store.device)
        self.step_count = 0
        self._step_t = torch.zeros((), dtype=torch.float32, device=store.device)
        self._native = native
        self._full_state: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
        self._bind_state()

    # ------------------------------------------------------------------ state plumbing
    @property
    def is_sharded(self) -> bool:
        return self.shard != (0, self.store.numel)

    def _bind_state(self) -> None:
        """Expose per-parameter views for parameters that lie fully inside the owned range."""
        self.state.clear()
        lo, hi = self.shard
        for p in self.store.param_list:
            o, n = self.store.segment(p)
            if o >= lo and o + n <= hi:
                self.state[p] = {
                    "step": self._step_t,
                    "exp_avg": self.store.view_like(self.exp_avg, p, base=lo),
                    "exp_avg_sq": self.store.view_like(self.exp_avg_sq, p, base=lo),
                }

    def consolidate_state_dict(self, to: int = 0) -> None:
        """Gather the moment shards (ZeRO-1) so that ``state_dict()`` on rank ``to`` is complete."""
        if not self.is_sharded:
            return
        world = dist.get_world_size()
        full_m = torch.empty(self.store.numel, dtype=self.exp_avg.dtype, device=self.store.device)
        full_v = torch.empty_like(full_m)
        dist.all_gather_into_tensor(full_m, self.exp_avg)
        dist.all_gather_into_tensor(full_v, self.exp_avg_sq)
        assert full_m.numel() == world * self.exp_avg.numel()
        self._full_state = (full_m, full_v)

    def state_dict(self):
        if self.is_sharded:
            if self._full_state is None:
                raise RuntimeError("call consolidate_state_dict() before state_dict() on a sharded optimizer")
            full_m, full_v = self._full_state
            saved = dict(self.state)
            self.state.clear()
            for p in self.store.param_list:
                o, n = self.store.segment(p)
                self.state[p] = {"step": self._step_t, "exp_avg": self.store.view_like(full_m, p), "exp_avg_sq": self.store.view_like(full_v, p)}
            out = super().state_dict()
            self.state.clear()
            self.state.update(saved)
            self._full_state = None
            return out
        return super().state_dict()

    def load_state_dict(self, state_dict):
        lo, hi = self.shard
        groups = state_dict["param_groups"]
        ids = [i for g in groups for i in g["params"]]
        if len(ids) != len(self.store.param_list):
            raise ValueError("loaded state dict contains a parameter group that doesn't match the size of optimizer's group")
        step_val = None
        for pid, p in zip(ids, self.store.param_list):
            st = state_dict["state"].get(pid)
            if not st:
                continue
            o, n = self.store.segment(p)
            a, b = max(o, lo), min(o + n, hi)
            if a < b:
                self.exp_avg[a - lo : b - lo].copy_(self.store.to_storage(p, st["exp_avg"])[a - o : b - o])
                self.exp_avg_sq[a - lo : b - lo].copy_(self.store.to_storage(p, st["exp_avg_sq"])[a - o : b - o])
            s = st.get("step")
            if s is not None:
                step_val = float(s.item() if torch.is_tensor(s) else s)
        if step_val is not None:
            self.step_count = int(step_val)
            self._step_t.fill_(step_val)
        for g_new, g_old in zip(groups, self.param_groups):
            for k, v in g_new.items():
                if k != "params":
                    g_old[k] = v
        self._bind_state()

    def zero_grad(self, set_to_none: bool = False):
        self.store.zero_grads()

    def advance_step(self, skip=None) -> None:
        """Count one update.  A NaN-skipped update must not advance Adam's step (upstream simply does not call
        ``optimizer.step()``, torchrun_main.py:813-822): the authoritative counter ``_step_t`` lives on the device, advances by
        ``1 - skip`` without a host round trip, is what the native kernels derive the bias corrections from and what
        ``state_dict()`` saves.  ``step_count`` is the host-side mirror (exact after :meth:`rollback_skipped_step`)."""
        self.step_count += 1
        if skip is None or not torch.is_tensor(skip):
            if not skip:
                self._step_t.add_(1)
        else:
            self._step_t.add_(1.0 - skip.reshape(()).to(self._step_t.dtype).clamp(0, 1))

    def undo_step_if_nonfinite(self, total_norm: torch.Tensor, skip=None) -> None:
        """Device-side: the kernels skipped this update because the gradient norm was non-finite; take the optimistic count back
        (unless the update was already skipped -- and not counted -- for a NaN loss)."""
        bad = (~torch.isfinite(total_norm)).to(self._step_t.dtype).reshape(())
        if skip is not None and torch.is_tensor(skip):
            bad = bad * (1.0 - skip.reshape(()).to(self._step_t.dtype).clamp(0, 1))
        elif skip:
            return
        self._step_t.sub_(bad)

    def rollback_skipped_step(self) -> None:
        """Host mirror of a skipped update (the trainer calls this after its host-side NaN check)."""
        if self.step_count > 0:
            self.step_count -= 1

    # ------------------------------------------------------------------ update
    @torch.no_grad()
    def step(self, closure=None, *, grad_scale=1.0, skip=None, grads=None):
        group = self.param_groups[0]
        lr, (b1, b2), eps, wd = group["lr"], group["betas"], group["eps"], group["weight_decay"]
        lo, hi = self.shard
        p = self.store.params[lo:hi]
        g = (self.store.grads if grads is None else grads)[lo:hi]
        if self._native is not None and p.is_cuda and p.dtype == torch.bfloat16 and self.exp_avg.dtype == torch.bfloat16:
            # the native kernel is the bf16-parameter / bf16-moment AdamW; other dtypes (--dtype float32) take the PyTorch path below
            self.advance_step(skip)
            self._native.adamw_flat(p, g, self.exp_avg, self.exp_avg_sq, float(lr), b1, b2, eps, wd,
                                    self.step_count, grad_scale, skip, self._step_t)
            if torch.is_tensor(grad_scale):
                self.undo_step_if_nonfinite(grad_scale, skip)
            return None
        if skip is not None and bool(skip):
            return None
        gs = float(grad_scale) if not torch.is_tensor(grad_scale) else float(grad_scale.item())
        if not math.isfinite(gs):  # non-finite gradient norm: skip (the native kernels do the same on the device)
            return None
        self.step_count += 1
        self._step_t.add_(1)
        ref.adamw_step(p, g, self.exp_avg, self.exp_avg_sq, step=self.step_count, lr=lr, beta1=b1, beta2=b2,
                       eps=eps, weight_decay=wd, grad_scale=gs)
        return None

    # ------------------------------------------------------------------ ReLoRA optimizer reset
    @torch.no_grad()
    def prune_state(self, params, keys: List[str], kind: str, ratio: float, *, seed: int = 0, reset_index: int = 0) -> float:
        """Prune the moments of ``params`` (random / magnitude, per tensor like the reference
        ``training_utils.py:354-361``); returns the percentage of zeroed entries."""
        lo, hi = self.shard
        bufs = {"exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq}
        n_total = 0
        n_zero = torch.zeros((), dtype=torch.float32, device=self.store.device)
        self._gather_cache = {}
        if kind == "magnitude" and self.is_sharded:
            for key in keys:  # collective: every rank gathers, whether or not one of its tensors straddles a boundary
                self._gathered(key)
        for t_idx, p in enumerate(params):
            if id(p) not in self.store.index:
                continue
            o, n = self.store.segment(p)
            a, b = max(o, lo), min(o + n, hi)
            if a >= b:
                continue
            for k_idx, key in enumerate(keys):
                seg = bufs[key][a - lo : b - lo]
                if kind == "random":
                    s = ref.mix_seed(seed, reset_index, t_idx, k_idx)
                    if self._native is not None and seg.is_cuda and seg.dtype == torch.bfloat16:
                        self._native.random_prune_(seg, ratio, s, a - o)
                    else:
                        keep = ref.random_prune_keep_mask(s, b - a, ratio, device=seg.device, offset=a - o)
                        seg.mul_(keep.to(seg.dtype))
                else:
                    if self.is_sharded and (a != o or b != o + n):
                        # the tensor straddles a shard boundary: take the quantile over the whole tensor
                        full = self.store.view_like(self._gathered(key), p)
                        from ..relora.optim_reset import magnitude_threshold

                        thr = magnitude_threshold(full, ratio)
                        seg.mul_((seg.abs() > thr).to(seg.dtype))
                    elif self.store.is_padded(p):
                        # quantile over the logical tensor only (the zero padding must not shift it)
                        from ..relora.optim_reset import magnitude_pruning_

                        magnitude_pruning_(self.store.view_like(bufs[key], p, base=lo), ratio)
                    elif self._native is not None and seg.is_cuda and seg.dtype == torch.bfloat16:
                        self._native.magnitude_prune_(seg, ratio)
                    else:
                        from ..relora.optim_reset import magnitude_pruning_

                        magnitude_pruning_(seg, ratio)
                n_total += seg.numel()
                n_zero += (seg == 0).sum()
        self._gather_cache = {}
        return float(n_zero.item()) / (1e-7 + n_total) * 100

    def _gathered(self, key: str) -> torch.Tensor:
        """Full (all shards) copy of a moment buffer, gathered once per pruning call."""
        if key not in self._gather_cache:
            local = self.exp_avg if key == "exp_avg" else self.exp_avg_sq
            full = torch.empty(self.store.numel, dtype=local.dtype, device=local.device)
            dist.all_gather_into_tensor(full, local.contiguous())
            self._gather_cache[key] = full
        return self._gather_cache[key]

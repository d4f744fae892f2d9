"""Liveness-based re-use of activation memory inside a forward plan.

A plan is a flat list of kernel launches over buffers allocated once.  Round 1 gave every activation its own buffer
(30 GB at B = 32).  Here a first ("probe") build of the same plan at batch 1 records, for every allocation, the last
step that touches it and the graph branch those steps run on; the real build then hands the memory of a finished buffer to
later allocations OF THE SAME BRANCH (launches of one branch are stream-ordered, so a buffer whose last reader precedes the
new buffer's first writer in that branch is dead; nothing is shared across branches, which run concurrently in the captured
graph).  Buffers the caller can see (outputs, taps), buffers written outside the graph (the padded input frames) and buffers
no step references are never re-used.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Tuple

import torch

ALIGN = 1024


def tensors_of(obj, depth: int = 0) -> Iterable[torch.Tensor]:
    """Every torch.Tensor reachable from a closure's defaults / cells (tuples, lists, dicts, Act-like objects)."""
    if depth > 4 or obj is None:
        return
    if isinstance(obj, torch.Tensor):
        yield obj
    elif isinstance(obj, (tuple, list, set, frozenset)):
        for o in obj:
            yield from tensors_of(o, depth + 1)
    elif isinstance(obj, dict):
        for o in obj.values():
            yield from tensors_of(o, depth + 1)
    elif hasattr(obj, "buf") and isinstance(getattr(obj, "buf", None), torch.Tensor):
        yield obj.buf


def step_tensors(fn) -> List[torch.Tensor]:
    out = list(tensors_of(getattr(fn, "__defaults__", None)))
    for cell in getattr(fn, "__closure__", None) or ():
        try:
            out.extend(tensors_of(cell.cell_contents))
        except ValueError:
            pass
    return out


def compute_liveness(allocs, steps, step_branch, pinned: Iterable[torch.Tensor]) -> Dict[int, Tuple[int, int]]:
    """allocs: [(index, tensor, step_at_allocation, branch)].  Returns {index: (branch, last_step)} for the allocations whose
    memory may be handed on after `last_step`; everything else is absent (never re-used)."""
    spans = sorted(((a[1].data_ptr(), a[1].data_ptr() + a[1].numel() * a[1].element_size(), a[0]) for a in allocs))
    starts = [s[0] for s in spans]
    import bisect

    def owner(t: torch.Tensor) -> Optional[int]:
        if t.numel() == 0:
            return None
        p = t.data_ptr()
        j = bisect.bisect_right(starts, p) - 1
        if j >= 0 and spans[j][0] <= p < spans[j][1]:
            return spans[j][2]
        return None

    last: Dict[int, int] = {}
    branches: Dict[int, set] = {}
    for i, (_name, fn) in enumerate(steps):
        for t in step_tensors(fn):
            k = owner(t)
            if k is not None:
                last[k] = max(last.get(k, -1), i)
                branches.setdefault(k, set()).add(step_branch[i])
    pinned_ids = {owner(t) for t in pinned if isinstance(t, torch.Tensor)}
    out = {}
    for k, _t, _s, br in allocs:
        if k in pinned_ids or k not in last:
            continue
        if branches[k] == {br}:
            out[k] = (br, last[k])
    return out


class Arena:
    """Hands out activation buffers for one plan build, re-using the memory of dead buffers of the same branch."""

    def __init__(self, liveness: Dict[int, Tuple[int, int]], device):
        self.liveness = liveness
        self.device = device
        self.free: Dict[int, List[Tuple[int, torch.Tensor]]] = {}     # branch -> [(nbytes, uint8 storage view)]
        self.live: List[Tuple[int, int, torch.Tensor]] = []            # (last_step, branch, storage) of re-usable live buffers
        self.fresh_bytes = 0
        self.reused_bytes = 0

    def _release(self, step_now: int):
        keep = []
        for last, br, st in self.live:
            if last < step_now:
                self.free.setdefault(br, []).append((st.numel(), st))
            else:
                keep.append((last, br, st))
        self.live = keep

    def alloc(self, k: int, shape, dtype, step_now: int, branch: int) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= s
        es = torch.empty((), dtype=dtype).element_size()
        need = -(-(n * es) // ALIGN) * ALIGN
        self._release(step_now)
        pool = self.free.get(branch, [])
        best = None
        for j, (sz, _st) in enumerate(pool):
            if sz >= need and (best is None or sz < pool[best][0]):
                best = j
        if best is not None:
            sz, st = pool.pop(best)
            if sz - need >= 64 * ALIGN:          # split: the remainder stays available
                pool.append((sz - need, st[need:]))
                st = st[:need]
            self.reused_bytes += need
        else:
            st = torch.empty(need, dtype=torch.uint8, device=self.device)
            self.fresh_bytes += need
        info = self.liveness.get(k)
        if info is not None and info[0] == branch:
            self.live.append((info[1], branch, st))
        return st[: n * es].view(dtype).view(shape)

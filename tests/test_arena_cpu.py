"""Host logic of the activation arena (mspi_b200/arena.py) on CPU tensors: liveness from step closures, re-use only inside
a branch, pinned and cross-branch buffers never recycled, and a recycled plan computing the same values."""
import torch

from mspi_b200.arena import Arena, compute_liveness, step_tensors


class _Plan:
    """Five buffers, two branches: a -> b -> c on branch 0, d on branch 1 reading a, e on branch 0 after c."""

    def __init__(self, n, liveness=None):
        self.steps, self.step_branch, self.allocs = [], [], []
        self.arena = Arena(liveness, "cpu") if liveness is not None else None
        self.branch = 0
        x = self.x = torch.arange(n, dtype=torch.float32)
        a = self.new(n)
        self.add(lambda a=a, x=x: a.copy_(x * 2))
        b = self.new(n)
        self.add(lambda a=a, b=b: b.copy_(a + 1))
        self.branch = 1
        d = self.new(n)
        self.add(lambda a=a, d=d: d.copy_(a * a))
        self.branch = 0
        c = self.new(n)
        self.add(lambda b=b, c=c: c.copy_(b * 3))
        e = self.new(n)
        self.add(lambda c=c, e=e: e.copy_(c - 1))
        f = self.new(n)
        self.add(lambda e=e, f=f: f.copy_(e + 0.5))
        self.out = (f, d)
        self.bufs = dict(a=a, b=b, c=c, d=d, e=e, f=f)

    def new(self, n):
        k = len(self.allocs)
        if self.arena is not None:
            t = self.arena.alloc(k, (n,), torch.float32, len(self.steps), self.branch)
        else:
            t = torch.empty(n)
        self.allocs.append((k, t, len(self.steps), self.branch))
        return t

    def add(self, fn):
        self.steps.append(("s%d" % len(self.steps), fn))
        self.step_branch.append(self.branch)

    def run(self):
        for _n, fn in self.steps:
            fn()
        return [t.clone() for t in self.out]


def test_step_tensors_finds_defaults_and_cells():
    a, b = torch.zeros(4), torch.ones(4)

    def f(x=a):
        return x + b
    ts = step_tensors(f)
    assert any(t is a for t in ts) and any(t is b for t in ts)


def test_liveness_excludes_cross_branch_and_pinned():
    p = _Plan(4096)
    lv = compute_liveness(p.allocs, p.steps, p.step_branch, pinned=p.out)
    # a (index 0) is read on branch 1 as well; d (2) and f (5) are outputs
    assert 0 not in lv and 2 not in lv and 5 not in lv
    assert lv[1] == (0, 3) and lv[3] == (0, 4) and lv[4] == (0, 5)


def test_arena_reuses_dead_buffers_and_keeps_results():
    n = 128 * 1024
    probe = _Plan(16)
    lv = compute_liveness(probe.allocs, probe.steps, probe.step_branch, pinned=probe.out)
    ref = _Plan(n).run()
    plan = _Plan(n, liveness=lv)
    got = plan.run()
    for r, g in zip(ref, got):
        assert torch.equal(r, g)
    # b is dead once c is written: e (allocated after step 3) takes b's memory; f is then given c's
    assert plan.bufs["e"].data_ptr() == plan.bufs["b"].data_ptr()
    assert plan.bufs["f"].data_ptr() == plan.bufs["c"].data_ptr()
    assert plan.bufs["d"].data_ptr() not in {plan.bufs[k].data_ptr() for k in "abcef"}
    assert plan.arena.reused_bytes == 2 * n * 4 and plan.arena.fresh_bytes == 4 * n * 4
    # a second run over the recycled buffers gives the same answer (nothing live was overwritten)
    for r, g in zip(ref, plan.run()):
        assert torch.equal(r, g)


def test_arena_splits_large_free_blocks():
    lv = {0: (0, 0)}
    ar = Arena(lv, "cpu")
    big = ar.alloc(0, (1 << 20,), torch.uint8, 0, 0)
    s1 = ar.alloc(1, (1 << 18,), torch.uint8, 1, 0)
    s2 = ar.alloc(2, (1 << 18,), torch.uint8, 1, 0)
    assert s1.data_ptr() == big.data_ptr() and s2.data_ptr() == big.data_ptr() + (1 << 18)
    other = ar.alloc(3, (1 << 18,), torch.uint8, 1, 1)      # another branch never takes it
    assert not (big.data_ptr() <= other.data_ptr() < big.data_ptr() + (1 << 20))

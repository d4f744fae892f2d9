"""world_size-2 (and 3) gloo tests of the clip-sharding host logic (mspi_b200/distributed.py) on CPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mspi_b200.distributed import forward_sharded, gather_maps, max_shard, shard_bounds


def test_shard_bounds_cover_batch_exactly_once():
    for n in (0, 1, 2, 5, 32, 33):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_bounds(n, r, world)
                assert 0 <= lo <= hi <= n and hi - lo <= max_shard(n, world)
                seen += list(range(lo, hi))
            assert seen == list(range(n))
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_forward(clips, audios):
    """A per-clip function (no cross-sample coupling), like the eval-mode model: map = mean over (c,t) + audio mean."""
    m = clips.mean((1, 2))
    if audios is not None:
        m = m + audios.mean((1, 2, 3)).view(-1, 1, 1)
    return m, m.mean((1, 2)).mean()


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)  # same global batch on every rank
        clips = torch.randn(n, 3, 4, 8, 12, generator=g)
        aud = torch.randn(n, 1, 5, 7, generator=g)
        full, loss = forward_sharded(_fake_forward, clips, aud)
        ref_maps, _ = _fake_forward(clips, aud)
        ref_loss = ref_maps.mean((1, 2)).mean() if n else torch.tensor(0.0)
        ok = torch.allclose(full, ref_maps, atol=1e-6) and abs(float(loss) - float(ref_loss)) < 1e-6
        # the asynchronous form returns the same maps after wait()
        lo_, hi_ = shard_bounds(n, rank, world)
        if hi_ > lo_ or True:
            local = ref_maps[lo_:hi_].contiguous()
            ok = ok and torch.allclose(gather_maps(local, n, async_op=True).wait(), ref_maps, atol=1e-6)
        # gather_maps rejects a shard of the wrong size
        lo, hi = shard_bounds(n, rank, world)
        try:
            gather_maps(torch.zeros(hi - lo + 1, 8, 12), n)
            ok = False
        except ValueError:
            pass
        q.put((rank, bool(ok), tuple(full.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 4), (2, 5), (3, 2), (2, 1)])
def test_forward_sharded_gloo(world, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert all(shape == (n, 8, 12) for _, _, shape in res)


# ------------------------------------------------------------------------------------------ training collective
def _train_worker(rank, world, port, q):
    """Each rank back-propagates the batch-mean loss of ITS clips through a small trainable stack (a slice of the oracle:
    the SimSiam projector head); the summed flat gradients times 1/world must equal the gradient of the mean over ranks of
    the per-rank losses, which is what the reference's DistributedDataParallel computes."""
    from mspi_b200.distributed import allreduce_gradients, flat_layout
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(3)
        w1, w2 = torch.randn(16, 8, generator=g), torch.randn(4, 16, generator=g)
        x = torch.randn(world, 3, 8, generator=g)            # 3 samples per rank

        def local_loss(a, b, xs):
            return torch.tanh(xs @ a.t()).matmul(b.t()).pow(2).mean()

        a, b = w1.clone().requires_grad_(True), w2.clone().requires_grad_(True)
        local_loss(a, b, x[rank]).backward()
        offs, n = flat_layout([a.shape, b.shape])
        flat = torch.zeros(n)
        flat[offs[0]:offs[0] + a.numel()] = a.grad.flatten()
        flat[offs[1]:offs[1] + b.numel()] = b.grad.flatten()
        scale = allreduce_gradients(flat)
        flat *= scale
        a2, b2 = w1.clone().requires_grad_(True), w2.clone().requires_grad_(True)
        sum(local_loss(a2, b2, x[r]) for r in range(world)).div(world).backward()
        ok = torch.allclose(flat[offs[0]:offs[0] + a.numel()].view_as(a2), a2.grad, atol=1e-6) and \
            torch.allclose(flat[offs[1]:offs[1] + b.numel()].view_as(b2), b2.grad, atol=1e-6) and offs[1] % 4 == 0
        q.put((rank, bool(ok), scale))
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_matches_ddp_averaging_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(ok and scale == 0.5 for _, ok, scale in res), res


# ------------------------------------------------------------------------------------------ DDP-style parameter broadcast
def _small_sd(seed):
    g = torch.Generator().manual_seed(seed)
    return {"visnet.a.weight": torch.randn(6, 5, generator=g), "visnet.a.bn.weight": torch.randn(6, generator=g),
            "visnet.a.bn.running_mean": torch.randn(6, generator=g), "visnet.a.bn.running_var": torch.rand(6, generator=g),
            "visnet.a.bn.num_batches_tracked": torch.tensor(seed), "audnet.c.weight": torch.randn(3, 3, generator=g),
            "readout.0.bias": torch.randn(7, generator=g)}


def _bcast_worker(rank, world, port, q):
    """Ranks seeded differently (the decoder / heads are random-init): after broadcast_training_state every rank holds rank
    0's parameters, buffers, moments and step count, and parameters_checksum() agrees."""
    from mspi_b200.distributed import broadcast_training_state, parameters_checksum
    from mspi_b200.train_engine import TrainState
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        st = TrainState(_small_sd(10 + rank), "cpu")
        st.flat_m.fill_(float(rank + 1))
        st.step_count = 3 + rank
        before = parameters_checksum(st)
        broadcast_training_state(st)
        ref = TrainState(_small_sd(10), "cpu")
        ok = (not before) and parameters_checksum(st) and torch.equal(st.flat_p, ref.flat_p) and st.step_count == 3
        ok = ok and torch.equal(st.live["visnet.a.bn.running_mean"], ref.live["visnet.a.bn.running_mean"])
        ok = ok and torch.equal(st.live["audnet.c.weight"], ref.live["audnet.c.weight"]) and float(st.flat_m.max()) == 1.0
        ok = ok and int(st.live["visnet.a.bn.num_batches_tracked"]) == 10
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_broadcast_training_state_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bcast_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res


def test_train_state_optimizer_state_dict_is_torch_adamw_compatible():
    """TrainState <-> torch.optim.AdamW.state_dict(): export after some steps loads into torch's AdamW, and torch's own state
    (after real torch steps) loads back with moments and step count intact (CPU; the kernels are not involved)."""
    from mspi_b200.train_engine import TrainState, trainable_keys
    sd = _small_sd(1)
    st = TrainState(sd, "cpu")
    keys = trainable_keys(sd)
    assert keys == ["visnet.a.weight", "visnet.a.bn.weight", "readout.0.bias"] and st.n_params == 30 + 6 + 7
    assert all(o % 4 == 0 for o in st.offs.values())
    params = [torch.nn.Parameter(sd[k].clone()) for k in keys]
    opt = torch.optim.AdamW(params, lr=1e-3, weight_decay=0)
    for _ in range(3):
        opt.zero_grad()
        sum((p ** 2).sum() for p in params).backward()
        opt.step()
    st.load_optimizer_state_dict(opt.state_dict())
    assert st.step_count == 3
    assert torch.allclose(st.view(st.flat_m, keys[0]), opt.state[params[0]]["exp_avg"])
    assert torch.allclose(st.view(st.flat_v, keys[2]), opt.state[params[2]]["exp_avg_sq"])
    osd = st.optimizer_state_dict(1e-3, (0.9, 0.999), 1e-8, 0.0)
    opt2 = torch.optim.AdamW([torch.nn.Parameter(sd[k].clone()) for k in keys], lr=1e-3, weight_decay=0)
    opt2.load_state_dict(osd)
    assert float(opt2.state[opt2.param_groups[0]["params"][1]]["step"]) == 3.0
    fresh = TrainState(sd, "cpu")
    assert fresh.optimizer_state_dict(1e-3, (0.9, 0.999), 1e-8, 0.0)["state"] == {}
    with pytest.raises(ValueError):
        fresh.load_optimizer_state_dict({"state": {0: opt.state_dict()["state"][0]}})

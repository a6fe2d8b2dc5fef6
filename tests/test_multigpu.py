"""Row sharding + allreduce.

CPU (gloo, world_size 2): the shard -> partial sums -> allreduce -> replicated m x m -> second pass logic, with the
oracle's reduced form standing in for the per-shard kernels, equals the literal single-process oracle.
GPU (needs >= 2 devices): two ranks, one context per GPU, NCCL allreduce inside the library == the 1-GPU result.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import cases


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _allreduce_dict(part):
    out = {}
    for k in sorted(part):
        t = torch.as_tensor(np.atleast_1d(np.asarray(part[k], dtype=np.float64)).copy())
        dist.all_reduce(t)
        out[k] = t.numpy().reshape(np.shape(part[k])) if np.ndim(part[k]) else float(t[0])
    return out


def _cpu_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bench import shard_bounds
    from oracle import reduced_model as red
    c = cases.config3(n=900, m=40)
    cp = c["cov_par"]
    l = cases.lvec(cp)
    lo, hi = shard_bounds(len(c["x"]), world, rank)
    assert (lo, hi) == red.shard_bounds(len(c["x"]), world)[rank]
    x, r = c["x"][lo:hi], (c["y"] - c["mu"])[lo:hi]
    p1 = _allreduce_dict(red.vi_pass1(x, r, c["xu"], cp["sigma"], l))
    mid = red.vi_mid(p1, c["xu"], cp["sigma"], l, cp["tau"], c["delta"])
    p2 = _allreduce_dict(red.vi_pass2(x, r, c["xu"], cp["sigma"], l, cp["tau"], mid))
    grad = red.vi_finish(p2, mid, c["xu"], cp["sigma"], l, cp["tau"])
    # knot gradient: the per-knot column sums travel in the same allreduce; bounds come from the WHOLE data set
    kgrad = red.knot_finish(p2["g_knot"], mid["N"], c["xu"], cp["sigma"], l, red.knot_bounds(c["x"]))
    q.put((rank, mid["obj"], grad, kgrad))
    dist.destroy_process_group()


def test_sharded_reduction_logic_gloo_world2():
    from oracle import reduced_model as red
    from oracle import ref_model as rm
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cpu_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    c = cases.config3(n=900, m=40)
    obj_ref, g_ref = rm.vi_obj_grad(c["cov_par"], "ard", c["xu"], c["x"], c["y"], c["mu"], c["delta"])
    _, _, kg_ref = red.vi_obj_grad(c["x"], c["y"], c["mu"], c["xu"], c["cov_par"]["sigma"], cases.lvec(c["cov_par"]),
                                   c["cov_par"]["tau"], c["delta"], knots=True)
    for rank, obj, grad, kgrad in res:
        assert obj == pytest.approx(obj_ref, rel=1e-10)
        for k in g_ref:
            assert grad[k] == pytest.approx(g_ref[k], rel=1e-8, abs=1e-9)
        np.testing.assert_allclose(kgrad, kg_ref, rtol=1e-9, atol=1e-9 * np.abs(kg_ref).max())
    # every rank holds the same global result
    assert res[0][1] == res[1][1]
    np.testing.assert_array_equal(res[0][3], res[1][3])


def test_shard_bounds_cover_rows():
    from bench import shard_bounds
    for n in (1, 7, 1000, 1_000_000):
        for world in (1, 2, 4, 8):
            b = [shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def _gpu_worker(rank, world, port, q, model):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bench import shard_bounds
    from sparsergps_b200.context import Context
    c = cases.config5(n=30011, m=300)
    cp = c["cov_par"]
    ctx = Context(rank)
    uid = [Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, 0)
    ctx.comm_init(world, rank, uid[0])
    lo, hi = shard_bounds(len(c["x"]), world, rank)
    ctx.set_data(c["x"][lo:hi], c["y"][lo:hi], None)
    from sparsergps_b200.vi_functions import knot_bounds
    obj, grad, kg, _ = ctx.gauss_obj_grad_knots(model, "ard", c["xu"], cp["sigma"], cases.lvec(cp), cp["tau"], c["delta"],
                                                knot_bounds(c["x"]))
    q.put((rank, obj, grad, kg))
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["vi", "fic"])
def test_two_gpus_match_one_gpu(ctx, model):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    c = cases.config5(n=30011, m=300)
    cp = c["cov_par"]
    ctx.set_data(c["x"], c["y"], None)
    from sparsergps_b200.vi_functions import knot_bounds
    obj1, g1, kg1, _ = ctx.gauss_obj_grad_knots(model, "ard", c["xu"], cp["sigma"], cases.lvec(cp), cp["tau"], c["delta"],
                                                knot_bounds(c["x"]))
    mctx = mp.get_context("spawn")
    q = mctx.Queue()
    port = _free_port()
    procs = [mctx.Process(target=_gpu_worker, args=(r, 2, port, q, model)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, obj, grad, kg in res:
        assert obj == pytest.approx(obj1, rel=1e-11)               # summation order differs across shards
        np.testing.assert_allclose(grad, g1, rtol=1e-9, atol=1e-9 * np.abs(g1).max())
        np.testing.assert_allclose(kg, kg1, rtol=1e-9, atol=1e-9 * np.abs(kg1).max())


@pytest.mark.gpu
def test_contexts_on_two_devices_in_one_process(ctx):
    """One process may hold a context per GPU (include/srgp.h): per-device kernel attributes must be set on both."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from sparsergps_b200.context import Context
    c = cases.config5(n=5000, m=200)
    cp = c["cov_par"]
    ctx.set_data(c["x"], c["y"], None)
    ref = ctx.gauss_obj_grad("vi", "ard", c["xu"], cp["sigma"], cases.lvec(cp), cp["tau"], c["delta"])
    other = Context(1)
    try:
        other.set_data(c["x"], c["y"], None)
        got = other.gauss_obj_grad("vi", "ard", c["xu"], cp["sigma"], cases.lvec(cp), cp["tau"], c["delta"])
        fic = other.gauss_obj_grad("fic", "ard", c["xu"], cp["sigma"], cases.lvec(cp), cp["tau"], c["delta"])
    finally:
        other.close()
    assert got[0] == ref[0]
    np.testing.assert_array_equal(got[1], ref[1])
    assert np.isfinite(fic[0])

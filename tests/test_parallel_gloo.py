"""World-size-2 test of the pass-sharding + reduce path on CPU (gloo).  Each rank renders its
pass blocks (here with the oracle standing in for the GPU, since the sharding arithmetic and
the reduce are what is under test), the buffers are SUM-reduced onto rank 0 and must equal a
single-process render of the same pass set."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, js, steps, P, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from jsraytracer_b200.parallel import pass_block, reduce_accum
    from oracle.oracle import OracleScene
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = OracleScene(js)
    acc = np.zeros((orc.height, orc.width, 3), dtype=np.float32)
    mine = []
    for step in range(steps):
        first, n = pass_block(step, rank, world, P)
        orc.render(n, first_pass=first, seed=1, accum=acc, threads=2)
        mine += list(range(first, first + n))
    t = torch.from_numpy(acc)
    reduce_accum(t, dst=0)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        np.savez(out_path, acc=t.numpy(), passes=np.array(sorted(sum(gathered, []))))
    dist.destroy_process_group()


def test_pass_sharding_and_reduce_world2(tmp_path, blobs):
    import torch.multiprocessing as mp
    from oracle.oracle import OracleScene
    js, _ = blobs("BoxBall", width=48, height=32)
    steps, P, world = 2, 2, 2
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(world, _free_port(), js, steps, P, out), nprocs=world, join=True)
    z = np.load(out)
    assert z["passes"].tolist() == list(range(steps * P * world))          # disjoint and complete
    serial, _ = OracleScene(js).render(steps * P * world, seed=1, threads=2)
    assert np.allclose(z["acc"], serial, rtol=1e-5, atol=1e-6)


def _stripe_worker(rank, world, port, js, passes, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from jsraytracer_b200.parallel import column_stripe, gather_columns
    from oracle.oracle import OracleScene
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = OracleScene(js)
    full, _ = orc.render(passes, seed=1, threads=2)
    x_offset, x_delt = column_stripe(rank, world)
    acc = np.zeros_like(full)
    acc[:, x_offset::x_delt, :] = full[:, x_offset::x_delt, :]      # what render(img, .., x_offset, x_delt) fills in
    t = torch.from_numpy(acc)
    gather_columns(t, dst=0)
    if rank == 0:
        np.savez(out_path, acc=t.numpy(), full=full)
    dist.destroy_process_group()


@pytest.mark.parametrize("width", [48, 47])          # 47: the stripes have different widths
def test_column_stripes_and_gather_world2(tmp_path, blobs, width):
    """The reference's own sharding (interleaved columns per worker, src/worker.js:30-32) + one gather."""
    import torch.multiprocessing as mp
    js, _ = blobs("BoxBall", width=width, height=32)
    out = str(tmp_path / "s0.npz")
    mp.spawn(_stripe_worker, args=(2, _free_port(), js, 2, out), nprocs=2, join=True)
    z = np.load(out)
    assert np.array_equal(z["acc"], z["full"])


def test_pass_block_and_shard_helpers():
    from jsraytracer_b200.parallel import pass_block, shard_passes
    seen = []
    for step in range(3):
        for r in range(4):
            f, n = pass_block(step, r, 4, 8)
            seen += list(range(f, f + n))
    assert sorted(seen) == list(range(96))
    allp = sum([[p for p, _ in shard_passes(10, r, 3)] for r in range(3)], [])
    assert sorted(allp) == list(range(10))


def test_strong_share_partitions_a_step():
    from jsraytracer_b200.parallel import strong_share
    for P in (1, 2, 7, 16, 33):
        for world in (1, 2, 3, 4, 8, 16):
            shares = [strong_share(P, r, world) for r in range(world)]
            assert sum(shares) == P and max(shares) - min(shares) <= 1 and shares == sorted(shares, reverse=True)

"""The N > 1 plumbing of bench.py / the corpus sweep on CPU: two processes, gloo, 127.0.0.1.
What is checked is what the multi-GPU path relies on: every image is owned by exactly one rank
(i mod world), the timed region is the MAX over ranks, rows merge back in image order on rank 0."""
import os
import socket

import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from oavif_b200.host import dist
    r, w, _ = dist.init("gloo")
    assert (r, w) == (rank, world)
    mine = dist.shard_indices(11, r, w)
    dist.barrier()
    ms = dist.max_over_ranks(10.0 + 5.0 * r)          # slowest rank defines the step time
    total = dist.sum_over_ranks(float(len(mine)))     # units processed by the whole job
    rows = dist.gather_rows([(i, f"img{i}", r) for i in mine])
    q.put((r, mine, ms, total, rows))
    dist.finalize()


def test_two_rank_sharding_timing_and_merge():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    (r0, m0, ms0, t0, rows0), (r1, m1, ms1, t1, rows1) = res
    assert m0 == [0, 2, 4, 6, 8, 10] and m1 == [1, 3, 5, 7, 9]
    assert sorted(m0 + m1) == list(range(11))                     # a partition: nothing lost, nothing doubled
    assert ms0 == ms1 == 15.0 and t0 == t1 == 11.0
    assert rows1 is None and [r[0] for r in rows0] == list(range(11))
    assert all(r[2] == r[0] % 2 for r in rows0)


def test_single_process_fallbacks():
    from oavif_b200.host import dist
    assert dist.shard_indices(5, 0, 1) == [0, 1, 2, 3, 4]
    assert dist.max_over_ranks(3.5) == 3.5 and dist.sum_over_ranks(2.0) == 2.0
    assert dist.gather_rows([(2, "b"), (1, "a")]) == [(1, "a"), (2, "b")]

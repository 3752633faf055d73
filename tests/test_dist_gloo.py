"""World-size-2 check (gloo, CPU) of the multi-GPU plumbing bench.py uses: pairs sharded r::W with no data-path
collective and ONE all_gather of the packed per-pair metric rows."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_pairs, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from geotransformer_b200.distributed import gather_metric_rows, max_over_ranks, pair_ids     # what bench.py calls
    from geotransformer_b200.synth import make_pair
    rows = []
    for pid in pair_ids(n_pairs // world, rank, world):
        pair = make_pair('modelnet717', pid)
        rows.append([float(pair['ref_points'].sum()), float(pair['transform'][0, 3]), 0.0, float(pid), 0.0, 0.0, 0.0, 1.0])
    rows = torch.tensor(rows, dtype=torch.float32)
    allrows = gather_metric_rows(rows, world)
    tmax = max_over_ranks(float(rank + 1), torch.device('cpu'), world)    # the max-over-ranks timing reduction
    if rank == 0:
        q.put((allrows.numpy(), tmax))
    dist.barrier()
    dist.destroy_process_group()


def test_pair_sharding_and_metric_all_gather():
    world, n_pairs = 2, 6
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pairs, q)) for r in range(world)]
    for p in procs:
        p.start()
    rows, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 2.0
    ids = sorted(rows[:, 3].astype(int).tolist())
    assert ids == list(range(n_pairs)), 'every pair processed exactly once across ranks'
    from geotransformer_b200.synth import make_pair
    for r in rows:
        pair = make_pair('modelnet717', int(r[3]))
        assert abs(r[0] - float(pair['ref_points'].sum())) < 1e-2   # results do not depend on the world size

"""CPU, world_size 2, gloo: the multi-process logic of the data-parallel path (flat-bucket gradient all-reduce with
the 1/world scale folded into the update, parameter broadcast, case sharding + metric gather)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import chap_losses as L


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from chap_b200 import parallel
    r, _, w = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(100 + rank)                      # replicas start different, then adopt rank 0's weights
    flat_p = torch.randn(1003)
    parallel.broadcast_parameters(flat_p, 0)
    g = torch.Generator().manual_seed(7 + rank)        # every replica sees its own batch -> its own gradient
    flat_g = torch.randn(1003, generator=g)
    local_g = flat_g.clone()
    parallel.make_grad_hook(world)(flat_g)             # SUM over ranks
    buf = [None]
    params = [flat_p.clone()]
    L.sgd_momentum_step(params, [flat_g / world], buf, lr=0.01)      # == grad_scale 1/world in the fused kernel
    cases = parallel.shard_round_robin(7, rank, world)
    rows = parallel.gather_rows(np.array([[i, i * 0.5] for i in cases]), world)
    out[rank] = dict(p0=flat_p.clone(), p1=params[0], local_g=local_g, cases=cases, rows=rows)
    dist.barrier()
    dist.destroy_process_group()


def test_flat_bucket_allreduce_and_case_sharding_world2():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = {k: out[k] for k in range(world)}
    assert torch.equal(res[0]["p0"], res[1]["p0"])                       # broadcast
    assert torch.equal(res[0]["p1"], res[1]["p1"])                       # identical update on both replicas
    mean_g = (res[0]["local_g"] + res[1]["local_g"]) / 2
    want = [res[0]["p0"].clone()]
    L.sgd_momentum_step(want, [mean_g], [None], lr=0.01)
    assert torch.allclose(res[0]["p1"], want[0], atol=1e-7)              # == single-process step on the mean gradient
    assert res[0]["cases"] == [0, 2, 4, 6] and res[1]["cases"] == [1, 3, 5]
    gathered = np.concatenate(res[0]["rows"])
    assert sorted(gathered[:, 0].tolist()) == list(range(7))


def test_single_process_helpers():
    from chap_b200 import parallel
    assert parallel.make_grad_hook(1) is None
    assert parallel.shard_round_robin(5, 0, 1) == [0, 1, 2, 3, 4]
    rows = np.zeros((2, 4))
    assert parallel.gather_rows(rows, 1) is rows

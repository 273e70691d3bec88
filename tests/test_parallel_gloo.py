"""CPU, world_size 2 over gloo: the host-side data-parallel logic (sharding, the single gradient
all-reduce, solver-parameter broadcast for smoothing, count reduction)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["CUDA_VISIBLE_DEVICES"] = ""          # this is the CPU / gloo test, also when run on a GPU box
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import metasolver_b200  # noqa: F401
    from metasolver_b200 import parallel
    from metasolver_b200.sopa.src.solvers.utils import create_solver, noise_params
    r, w, dev = parallel.init_distributed("gloo")
    assert (r, w) == (rank, world) and dev.type == "cpu"
    out = {}
    # 1. one all-reduce over the flat gradient buffer == mean of per-rank grads
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.Linear(4, 3))
    x = torch.full((2, 5), float(rank + 1))
    model(x).sum().backward()
    local = [p.grad.clone() for p in model.parameters()]
    red = parallel.GradAllReducer(model.parameters())
    red()
    out["grads"] = [p.grad.tolist() for p in model.parameters()]
    out["local"] = [g.tolist() for g in local]
    out["nbytes"] = red.nbytes
    # 2. solver smoothing: ranks draw different u, rank 0's value wins everywhere
    s = create_solver("rk2", "u", 8, -1, 0.5, -1, torch.float32, "cpu")
    s.freeze_params()
    torch.manual_seed(100 + rank)
    s.u, s.v = noise_params(s.u0, s.v0, std=0.0125, bernoulli_p=1.0, noise_type="normal")
    s.build_ButcherTableau()
    out["u_before"] = float(s.u)
    parallel.sync_solver_params([s])
    out["u_after"] = float(s.u)
    out["tab"] = s.host_tableau()
    # 3. sharding and count reduction
    lo, hi = parallel.shard_range(11, rank, world)
    out["shard"] = (lo, hi)
    out["count"] = parallel.allreduce_sum_int(hi - lo, dev)
    out["counts"] = parallel.allreduce_sum_counts(torch.tensor([hi - lo, rank + 1, 0]))     # whole sweep vector, one all-reduce
    # the peer-memory exchange needs CUDA devices: refused loudly here, and the reducers keep the collective path
    try:
        parallel.PeerExchange(8)
        out["peer_refused"] = False
    except RuntimeError as exc:
        out["peer_refused"] = "CUDA" in str(exc)
    out["peer_or_none"] = parallel.peer_exchange_or_none(8) is None
    red_p = parallel.GradAllReducer(model.parameters(), peer=True)
    out["reducer_peer_is_none"] = red_p.peer is None
    # 4. a sharded synthetic evaluation set (scripts/eval_pgd_sweep.py): every rank generates only its slice
    from metasolver_b200 import detrand
    out["shard_data"] = detrand.uniform((hi - lo, 3, 4, 4), 9100, 0.0, 1.0, offset=lo * 48).tolist()
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_host_logic_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    a, b = res[0], res[1]
    for g0, g1, l0, l1 in zip(a["grads"], b["grads"], a["local"], b["local"]):
        assert g0 == g1
        assert torch.allclose(torch.tensor(g0), (torch.tensor(l0) + torch.tensor(l1)) / 2, rtol=0, atol=1e-6)
    assert a["nbytes"] == 4 * (5 * 4 + 4 + 4 * 3 + 3)
    import numpy as np
    from metasolver_b200 import detrand
    full = detrand.uniform((11, 3, 4, 4), 9100, 0.0, 1.0)
    assert np.array_equal(np.concatenate([np.asarray(a["shard_data"], dtype=np.float32),
                                          np.asarray(b["shard_data"], dtype=np.float32)]), full)
    assert a["u_before"] != b["u_before"]
    assert a["u_after"] == b["u_after"] == pytest.approx(a["u_before"], abs=1e-7)
    assert a["tab"] == b["tab"]
    assert a["shard"] == (0, 6) and b["shard"] == (6, 11)
    assert a["count"] == b["count"] == 11
    assert a["counts"] == b["counts"] == [11, 3, 0]
    for o in (a, b):
        assert o["peer_refused"] is True and o["peer_or_none"] and o["reducer_peer_is_none"]

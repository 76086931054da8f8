"""world_size-2 tests on CPU (gloo): the host-side logic of the multi-GPU path -- agreeing on one ncclUniqueId,
and the row dealing that makes shards equal-sized and rating-balanced."""
import os
import socket
import subprocess
import sys
import textwrap
from pathlib import Path

import numpy as np

from pio_b200 import sharding, synth

ROOT = Path(__file__).resolve().parent.parent


def test_row_dealing_properties():
    nu, ni, nnz = 5000, 700, 200000
    u, i, r = synth.synth_ratings(nu, ni, nnz, seed=3)
    for W in (1, 2, 4, 8):
        for n, idx in ((nu, u), (ni, i)):
            deg = np.bincount(idx, minlength=n)
            perm, inv = sharding.assign_internal(deg, W)
            R = sharding.rows_per_rank(n, W)
            assert np.array_equal(inv[perm], np.arange(n))                       # bijection onto used ids
            own = sharding.owner_rank(perm, n, W)
            cnt = np.bincount(own, minlength=W)
            assert cnt.max() <= R and cnt.sum() == n                             # equal (padded) shards
            load = np.bincount(own, weights=deg, minlength=W)
            assert load.max() <= load.mean() * 1.05 + deg.max()                  # ratings balanced
            for rk in range(W):                                                   # a rank's rows: degree-descending
                rows = inv[rk * R:(rk + 1) * R]
                d = deg[rows[rows >= 0]]
                assert (np.diff(d) <= 0).all()


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import numpy as np, torch, torch.distributed as dist
    import pio_b200
    from pio_b200 import sharding, synth, workflow
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    sc = workflow.WorkflowContext()
    assert (sc.world_rank, sc.world_size) == (rank, world)
    nid = sc.new_nccl_id()                       # rank 0 creates, everyone receives the same 128 bytes
    got = [None] * world
    dist.all_gather_object(got, nid)
    assert len(nid) == 128 and all(g == nid for g in got), "ranks disagree on the ncclUniqueId"
    # every rank derives the same dealing from the same ratings and owns a disjoint, covering set of rows
    u, i, r = synth.synth_ratings(3000, 400, 50000, seed=3)
    deg = np.bincount(u, minlength=3000)
    perm, inv = sharding.assign_internal(deg, world)
    mine = np.flatnonzero(sharding.owner_rank(perm, 3000, world) == rank)
    sizes = [None] * world
    dist.all_gather_object(sizes, (int(mine.size), int(deg[mine].sum()), mine.tolist()))
    assert sum(s[0] for s in sizes) == 3000 and sum(s[1] for s in sizes) == 50000
    allrows = sorted(x for s in sizes for x in s[2])
    assert allrows == list(range(3000))
    # what a multi-GPU training job must agree on and write exactly once: instance id, seed, registry, models
    seeds = [None] * world
    dist.all_gather_object(seeds, sc.agree_seed(None))
    assert len(set(seeds)) == 1, "ranks drew different initial-factor seeds"
    assert sc.agree_seed(7) == 7
    os.environ["PIO_MODELDATA_DIR"] = {md!r}
    inst_id = sc.broadcast_object("inst-" + os.urandom(4).hex())
    inst = workflow.EngineInstance(id=inst_id, status="INIT", startTime="t", endTime="t", engineId="e", engineVersion="1",
                                   engineVariant="default", engineFactory="f", batch="", env={{}}, sparkConf={{}},
                                   variantJson={{}})
    if rank == 0:
        workflow.EngineInstances.insert(inst)
    dist.barrier()
    got = workflow.EngineInstances.get(inst_id)
    assert got is not None and got.id == inst_id
    assert len(workflow.EngineInstances._load()) == 1, "the registry must hold ONE record for the job"
    from pio_b200 import controller
    class _Algo(controller.P2LAlgorithm):
        pass
    eng = controller.Engine.__new__(controller.Engine)
    out = eng.makeSerializableModels(sc, inst_id, [None], [_Algo()], [{{"w": rank}}])
    assert (out == [{{"w": 0}}]) if rank == 0 else (out == [controller.Unit]), "only rank 0 persists models"
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def test_two_process_gloo_agreement(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=str(ROOT), md=str(tmp_path / "modeldata")))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("ok") == 2

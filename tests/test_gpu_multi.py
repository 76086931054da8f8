"""Multi-GPU parity (-m gpu, needs >= 2 B200s; skipped otherwise): row-sharded training with the NCCL factor
all-gather gives bit-identical factors to the single-GPU run (every row's arithmetic is independent of the
sharding), and therefore the same parity with the oracle."""
import os
import socket
import subprocess
import sys
import textwrap
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import numpy as np, torch, torch.distributed as dist
    import pio_b200
    from pio_b200 import native, synth
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    def fresh_id():      # one ncclUniqueId per communicator (= per handle)
        t = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            t = torch.tensor(list(native.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, 0)
        return bytes(t.cpu().tolist())
    # (rank, implicit, users, items, ratings, PIO_ALS_TC): rank 64 runs the mma.sync kernel, plus the tcgen05 kernel on
    # sides with long rows unless PIO_ALS_TC=0; the last two cases have heavy item rows (parts + finish kernel)
    for (rk, implicit, nu, ni, nnz, tc) in ((64, True, 6000, 900, 150000, ""), (10, False, 3000, 500, 40000, ""),
                                            (64, True, 20000, 30, 300000, "0"), (64, True, 20000, 30, 300000, "")):
        os.environ.pop("PIO_ALS_TC", None)
        if tc:
            os.environ["PIO_ALS_TC"] = tc
        u, i, r = synth.synth_ratings(nu, ni, nnz, seed=3, implicit=implicit)
        u0 = synth.synth_init_factors(nu, rk, 5, 0)
        m = native.NativeALS(rk, nu, ni, lam=0.01, implicit=implicit, device=rank, world_size=world, world_rank=rank,
                             nccl_id=fresh_id())
        m.set_ratings(u, i, r, dedup=1 if implicit else 0)
        m.set_init(u0)
        m.run(3)
        uf, itf, uh, ih = m.get_factors()
        if rank == 0:
            s = native.NativeALS(rk, nu, ni, lam=0.01, implicit=implicit, device=0)
            s.set_ratings(u, i, r, dedup=1 if implicit else 0)
            s.set_init(u0)
            s.run(3)
            suf, sitf, suh, sih = s.get_factors()
            du, di = float(np.abs(uf - suf).max()), float(np.abs(itf - sitf).max())
            assert np.array_equal(uf, suf) and np.array_equal(itf, sitf), (
                "sharded != single GPU: rank %d implicit %s: max |du| %.3g (%d rows differ), max |di| %.3g (%d rows differ)"
                % (rk, implicit, du, int((uf != suf).any(1).sum()), di, int((itf != sitf).any(1).sum())))
            assert np.array_equal(uh, suh) and np.array_equal(ih, sih)
            assert m.stats()["last_comm_ms"] > 0
        dist.barrier()
        m.close()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def test_sharded_equals_single_gpu(native, tmp_path):
    n = native.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=str(ROOT)))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-6000:]
    assert out.stdout.count("ok") == world

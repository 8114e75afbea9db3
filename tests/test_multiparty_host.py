"""CPU tier, world_size 2 over gloo: the host-side multi-party logic (rendezvous, king's scatter of witness
shares, byte broadcast, max-reduction of timings).  The GPU data path of the same run is
tests/mp_groth16_check.py under torchrun on >= 2 GPUs."""
import os
import subprocess
import sys
import textwrap
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent

WORKER = textwrap.dedent("""
    import sys
    sys.path.insert(0, %r)
    import numpy as np
    import czk_b200
    from czk_b200 import launch
    from oracle import binding as o
    rank, world = launch.init_control_plane()
    assert world == 2
    k = 257
    secret = o.random_fr_mont(11, k)
    mine = launch.king_share_scatter(secret if rank == 0 else None, k, seed=5)
    ref = czk_b200.king_share_batch(secret, world, 5)
    assert (mine == ref[rank]).all()
    # shares reconstruct the secret (additive sharing, add.rs:105-117)
    import torch, torch.distributed as dist
    both = [torch.zeros((k, 4), dtype=torch.int64) for _ in range(world)]
    dist.all_gather(both, torch.from_numpy(mine.view(np.int64).copy()))
    total = both[0].numpy().view(np.uint64)
    total = o.fr_add(total, both[1].numpy().view(np.uint64))
    assert (total == secret).all()
    msg = launch.broadcast_bytes_from_king(bytes(range(128)) if rank == 0 else None, 128)
    assert msg == bytes(range(128))
    assert launch.max_over_ranks(float(rank + 1)) == 2.0
    launch.barrier()
    print("worker", rank, "ok")
""") % str(ROOT)


def test_two_party_control_plane_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611", WORLD_SIZE="2")
    procs = []
    for rank in range(2):
        e = dict(env, RANK=str(rank), LOCAL_RANK=str(rank))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {rank} failed:\n{out}"
        assert f"worker {rank} ok" in out


def test_king_share_batch_shape_and_sum(czk, oracle):
    x = oracle.random_fr_mont(1, 100)
    for n in (1, 2, 3, 8):
        sh = czk.king_share_batch(x, n, seed=2)
        assert sh.shape == (n, 100, 4)
        tot = sh[0]
        for p in range(1, n):
            tot = oracle.fr_add(tot, sh[p])
        assert (tot == x).all()
    assert (czk.king_share_batch(x, 1, 0)[0] == x).all()


def test_squaring_chain_matches_oracle(czk, oracle):
    s = oracle.random_fr_mont(3, 1)[0]
    assert (czk.squaring_chain(s, 300) == oracle.squaring_chain(s, 300)).all()

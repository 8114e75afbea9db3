"""The real N-party path - one rank per GPU, NCCL slice exchange / all-gather / king gather - under pytest: spawns
`torchrun tests/mp_groth16_check.py` for additive, SPDZ and GSZ and keeps the raw per-rank output under gpurun_out/.
Skipped on a box with one GPU (there the N-party ARITHMETIC is covered by the single-GPU simulation in
test_gpu_shares.py / test_gpu_gsz.py; this file covers the transport)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _gpus():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _run(nproc, scheme, extra=(), port=29641, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(ROOT / "tests" / "mp_groth16_check.py"), "--scheme", scheme, *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=str(ROOT), timeout=900, env=dict(os.environ, **(env or {})))
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    tag = ("_" + "_".join(f"{k}={v}" for k, v in (env or {}).items()) if env else "") + "".join(e.replace("--", "_") for e in extra)
    (out / f"mp_pytest_{scheme}_{nproc}{tag}.log").write_text(" ".join(cmd) + "\n" + r.stdout + "\n--- stderr ---\n" + r.stderr[-20000:])
    return r


@pytest.mark.parametrize("scheme", ["additive", "spdz", "gsz"])
@pytest.mark.parametrize("nproc", [2, 3, 4, 8])
def test_torchrun_multi_party_parity(scheme, nproc):
    if _gpus() < nproc:
        pytest.skip(f"needs {nproc} GPUs, this box has {_gpus()}")
    r = _run(nproc, scheme, port=29641 + nproc)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("parity ok") == nproc, r.stdout[-3000:]  # (ranks share stdout: their lines may run together)


def test_torchrun_both_share_transports():
    """The opens move over NVLink peer memory by default (CUDA IPC mappings, reduce kernel loading from / storing to the
    peers) and over grouped ncclSend / ncclRecv + all-gather with CZK_SHARE_TRANSPORT=nccl: the same parity run on both."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    r = _run(2, "spdz", port=29681, env={"CZK_SHARE_TRANSPORT": "nccl"})
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("parity ok") == 2 and "opens over nccl" in r.stdout, r.stdout[-3000:]
    r = _run(2, "spdz", port=29682)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("parity ok") == 2, r.stdout[-3000:]


def test_torchrun_corrupted_share_trips_the_mac_check():
    """Negative case over the real transport: one rank corrupts a MAC share, EVERY rank must get CZK_ERR_PROTOCOL."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    r = _run(2, "spdz", extra=["--corrupt-mac"], port=29671)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("corrupted MAC detected") == 2, r.stdout[-3000:]

"""bench.py's reference arm runs anywhere (no GPU): check the one-JSON-line contract the driver parses, and that only rank 0
of a torchrun launch does the work."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                           "--cpu-sample-log-n", "8"], capture_output=True, text=True, env=env, cwd=str(ROOT), timeout=300)


def test_reference_arm_prints_one_json_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "ms" and d["higher_is_better"] is False
    # a reduced sample (this test runs 2^8 to stay fast) must be labelled as such and never scaled into `value`:
    # the default (--cpu-sample-log-n 20) measures the BASELINE config itself
    assert d["metric"] == "groth16_proof_ms_2^8_r1cs_bls12_377"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] > 0 and "2^8" in cb["sample"]
    assert cb["same_config"] is False and cb["extrapolated"] is True and cb["log_n"] == 8
    assert abs(cb["extrapolated_2^20_ms"] - cb["value"] * 4096) < 1e-6 * cb["value"] * 4096
    assert d["e2e"] == {"value": d["value"], "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("groth16 spdz 2^8") and d["config"]["same_config"] is False


def test_reference_arm_defaults_to_the_baseline_config():
    sys.path.insert(0, str(ROOT))
    import bench

    ap_default = [a for a in open(ROOT / "bench.py").read().splitlines() if "add_argument(\"--cpu-sample-log-n\"" in a]
    assert ap_default and "default=LOG_N" in ap_default[0] and bench.LOG_N == 20


def test_reference_arm_other_ranks_exit_quietly():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_czk_arm_refuses_to_run_without_a_gpu():
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the czk arm would run the whole benchmark")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1"], capture_output=True, text=True, cwd=str(ROOT), timeout=300)
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr


def test_plonk_reference_arm_both_domain_shapes():
    """--workload plonk: the wiring argument on the host CPU over the reference's 3 * 2^k-point wire domain (default) and
    over a power-of-two domain; one JSON line each, labelled with the domain it ran."""
    for extra, needle in ((["--plonk-domain", "mixed"], "3*2^4"), (["--plonk-domain", "radix2"], "over a 2^4 domain")):
        r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "plonk", "--log-n", "4", "--steps", "1",
                            "--warmup", "0"] + extra, capture_output=True, text=True, cwd=str(ROOT), timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        lines = [l for l in r.stdout.splitlines() if l.strip()]
        assert len(lines) == 1
        d = json.loads(lines[0])
        assert d["impl"] == "reference" and d["metric"] == "plonk_wiring_proof_ms_2^4_bls12_377" and d["value"] > 0
        assert needle in d["config"]["workload"] and needle.replace("over a ", "") in d["cpu_baseline"]["sample"]
        assert d["e2e"]["h2d_bytes_per_step"] == 0

#!/usr/bin/env python3
"""bench.py - Groth16 proof time at 2^20 R1CS constraints on BLS12-377, one MPC party per GPU.

    python bench.py --gpus 1 --steps K --warmup W
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...       # the CPU path (oracle port of the reference) on the host cores

A "step" is one proof: create_random_proof + reveal (mpc-snarks/src/proof.rs:130-139) of the repeated-squaring
circuit with 2^20 squarings (D = 2^21) under SPDZ shares: 14 NTTs of 2^21, the Beaver product with its opens,
MSMs of 2^21-1, 2^20, 2^20+1 (x2) G1 terms and 2^20+1 G2 terms, and the O(1) share/group tail.  Synthetic data: a
seeded witness chain shared additively by the king, and a REAL CRS generated on the device from seeded toxic waste
(czk_groth16_setup) - one more proof is produced after the timed region and verified with the pairing check, so the
measured path is known to produce valid proofs (--key synthetic: shape-only bases, no verification).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

LOG_N = 20
METRIC = "groth16_proof_ms_2^20_r1cs_bls12_377"
# BASELINE.md section 1 (reference's published figures, GCP n2-standard-2, 1 physical core per party)
PUBLISHED_MS = {1: 127400.0, 2: 320400.0, 3: 323300.0}
# The dominant "kernel" is the bucket accumulation of one G1 MSM: 5 launches of k_bat_round<Fq> (halving rounds of the
# batched-affine tree) + k_bat_finish<Fq>; it is timed as one unit by CUDA events on the launching stream.
ACC_KERNEL = "k_bat_round<Fq> x5 + k_bat_finish<Fq> (bucket accumulation of one G1 MSM)"
# ncu --set full capture of those launches (profiles/r1_summary.md): dram__bytes_read + write of the 2^21-1 term MSM's
# round 0 (14.44 GB) and round 1 (4.57 GB) as captured, later rounds halving (sum 23.5 GB); the 2^20 term MSMs move half
# of it -> mean over the step's 4 G1 MSMs
NCU_TRAFFIC_BYTES_2_21 = 23.5e9
NCU_TRAFFIC_BYTES_PER_LAUNCH = NCU_TRAFFIC_BYTES_2_21 * (1 + 3 * 0.5) / 4


def ref_msm_adds(n: int) -> int:
    """SURVEY.md 8(d) normaliser: N*W + 2(2^c - 1)W + 253 with the reference's c, W."""
    def ark_log2(x):
        return 0 if x == 0 else (x.bit_length() - 1 if x & (x - 1) == 0 else x.bit_length())
    c = 3 if n < 32 else ark_log2(n) * 69 // 100 + 2
    w = (253 + c - 1) // c
    return n * w + 2 * ((1 << c) - 1) * w + 253


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 9:
                self.rows.append(parts)

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        pw = [float(r[3]) for r in self.rows if r[3].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(self.rows), "reasons": sorted(reasons)}


def cpu_reference_leg(steps: int, warmup: int, sample_log_n: int, scheme_name: str = "spdz", budget_s: float = 330.0):
    """Time the oracle (a C port of the reference's prover: `kind: port`) on the host cores, all threads.  With
    sample_log_n == LOG_N (the default) every step is one whole proof of the BASELINE config - a measurement, not an
    extrapolation.  A smaller sample (--cpu-sample-log-n) is reported as what it is: `value` stays the SAMPLE's ms and
    the line says `extrapolated_2^20_ms` separately.  `budget_s` bounds the timed steps (at least one always runs)."""
    from oracle import binding as o

    o.build()
    threads = o.cpu_threads()
    n_sq = 1 << sample_log_n
    # synthetic key of the right shapes: points from an arithmetic progression (no setup cost; the CPU prover's cost does
    # not depend on the key being a valid CRS)
    g1, g2 = o.generators()
    ks = o.random_fr_mont(1, 2)
    D = o.groth16_domain_size(n_sq)

    def pts(G, g, n):
        return G.gen_progression(g, ks[0], ks[1], n, threads=threads)

    t_key = time.perf_counter()
    pk = dict(n_sq=n_sq, D=D, a_query=pts(o.G1, g1, n_sq + 2), a_inf=None, b_g1_query=pts(o.G1, g1, n_sq + 2), b1_inf=None,
              b_g2_query=pts(o.G2, g2, n_sq + 2), b2_inf=None, h_query=pts(o.G1, g1, D - 1), h_inf=None,
              l_query=pts(o.G1, g1, n_sq), l_inf=None, vk_g1=pts(o.G1, g1, 3), vk_g2=pts(o.G2, g2, 3))
    t_key = time.perf_counter() - t_key
    chain = o.squaring_chain(o.random_fr_mont(2, 1)[0], n_sq)
    r, s = o.random_fr_mont(3, 1), o.random_fr_mont(4, 1)
    scheme = o.SCHEME_SPDZ if scheme_name == "spdz" else o.SCHEME_PLAIN
    times = []
    t_start = time.perf_counter()
    for i in range(warmup + steps):
        t = time.perf_counter()
        res = o.groth16_prove(scheme, n_sq, [chain], r, s, pk, threads=threads, want_h=False)
        dt = time.perf_counter() - t
        assert res["ok"]
        if i >= warmup:
            times.append(dt)
            if time.perf_counter() - t_start + dt > budget_s:  # the next step would overrun the budget
                break
    ms_sample = 1e3 * sum(times) / len(times)
    out = dict(value=ms_sample, unit="ms", cores=threads, kind="port", log_n=sample_log_n, steps_run=len(times),
               same_config=(sample_log_n == LOG_N), key_s=t_key,
               sample=f"{len(times)} whole Groth16 {scheme_name} proof(s) of one party at 2^{sample_log_n} constraints on {threads} host "
                      f"threads (window-parallel MSM, chunk-parallel NTT: the reference's dormant Rayon split), {ms_sample:.1f} ms each, "
                      + ("the BASELINE config itself - measured, not extrapolated" if sample_log_n == LOG_N
                         else f"a reduced sample: NOT the 2^{LOG_N} config"))
    if sample_log_n != LOG_N:
        out["extrapolated_2^20_ms"] = ms_sample * float(1 << (LOG_N - sample_log_n))
        out["extrapolated"] = True
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # a CPU loop needs no warm-up beyond one pass (page faults, thread pool); the timed steps honour --steps within the budget
    cb = cpu_reference_leg(max(1, args.steps), max(0, min(args.warmup, 1)), args.cpu_sample_log_n)
    line = {"impl": "reference", "metric": METRIC if cb["same_config"] else f"groth16_proof_ms_2^{args.cpu_sample_log_n}_r1cs_bls12_377",
            "value": cb["value"], "unit": "ms", "n_gpus": args.gpus, "steps": cb["steps_run"],
            "warmup": max(0, min(args.warmup, 1)), "ms_per_step": cb["value"], "higher_is_better": False, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64 limbs (Montgomery), integer", "data": "synthetic",
            "config": {"workload": f"groth16 spdz 2^{args.cpu_sample_log_n} constraints (D=2^{args.cpu_sample_log_n + 1}) BLS12-377, one party's work on the "
                                   f"host CPU (oracle port of the reference prover, {cb['cores']} threads)",
                       "parties": args.gpus, "same_config": cb["same_config"]},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="czk", choices=["czk", "reference"])
    ap.add_argument("--log-n", type=int, default=LOG_N, help="log2 constraints (default 20 = the BASELINE config)")
    ap.add_argument("--scheme", default="spdz", choices=["spdz", "additive", "plain", "gsz"])
    ap.add_argument("--cpu-sample-log-n", type=int, default=LOG_N,
                    help="constraints of the CPU arm (default: the BASELINE config itself, 2^20)")
    ap.add_argument("--key", default="real", choices=["real", "synthetic"],
                    help="real: CRS generated on the device from seeded toxic waste, the timed proof is verified with the pairing "
                         "check after the timed region; synthetic: device-generated bases of the same shapes (no verification)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch

    import czk_b200
    from czk_b200 import launch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the czk arm has no CPU fallback (use --impl reference for the CPU path)")
    warmup = max(args.warmup, 3)
    # NCCL prints its version banner on stdout when NCCL_DEBUG is set in the environment; stdout must carry exactly
    # one JSON line, so fd 1 points at stderr while the communicator comes up
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        party = launch.Party()
        party.ctx.batch_open(czk_b200.SCHEME_ADDITIVE, party.ctx.vec(4))  # first collective: forces the lazy NCCL init now
        party.ctx.sync()
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    ctx, rank, world = party.ctx, party.rank, party.world
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    scheme = {"spdz": czk_b200.SCHEME_SPDZ, "additive": czk_b200.SCHEME_ADDITIVE, "plain": czk_b200.SCHEME_PLAIN,
              "gsz": czk_b200.SCHEME_GSZ}[args.scheme]
    n_sq = 1 << args.log_n

    # ---- setup (untimed, like the reference: CRS + king_share_batch happen before start_timer!, proof.rs:113-129)
    imad_peak, _ = ctx.microbench(0, 8, 256, 2000)  # measured IMAD.WIDE.U32 issue rate: the integer roofline denominator
    key_kind = args.key
    pk = None
    if key_kind == "real":
        try:
            rng = np.random.Generator(np.random.PCG64(0x377))
            toxic = rng.integers(0, 1 << 64, size=(7, 4), dtype=np.uint64)
            toxic[:, 3] &= np.uint64((1 << 60) - 1)  # < 2^252 < r: valid Montgomery limbs, same on every rank
            pk = czk_b200.groth16_setup(ctx, n_sq, toxic)
        except Exception as exc:  # fall back to the shape-only key rather than lose the measurement
            print(f"bench.py: real key generation failed ({exc}); using a synthetic key", file=sys.stderr)
            key_kind = "synthetic"
    if pk is None:
        pk = czk_b200.ProvingKey.synthetic(ctx, n_sq, seed=0x377)
    D = pk.domain_size
    start = np.array([0x1234567, 0x89abcdef, 0x55aa55aa, 0x0123], np.uint64)
    if args.scheme == "gsz":
        # king_share_batch under GSZ hands the plaintext to every party (gsz20/mod.rs:202-212): each rank derives it
        mine = czk_b200.squaring_chain(start, n_sq)
    else:
        chain = czk_b200.squaring_chain(start, n_sq) if rank == 0 else None
        mine = launch.king_share_scatter(chain, n_sq + 1, seed=0x5eed)
    pinned = torch.from_numpy(mine.view(np.int64).copy()).pin_memory()
    mine_pinned = pinned.numpy().view(np.uint64)
    chain_dev = ctx.vec_from(mine)
    rho = np.array([0x0123456789abcdef, 0x0fedcba987654321, 0x1111222233334444, 0x0000000055556666], np.uint64)
    sig = np.array([0x0aaaaaaabbbbbbbb, 0x0ccccccccddddddd, 0x0eeeeeeeffffffff, 0x0000000012345678], np.uint64)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{ctx.device}")  # > 126 MB L2

    def step(resident: bool):
        flush.zero_()  # L2 flush between iterations (inputs per proof are also > L2: 64 MB vectors x 6, 100-200 MB bases)
        torch.cuda.synchronize()
        t = time.perf_counter()
        res = czk_b200.groth16_prove(ctx, scheme, pk, chain_dev if resident else mine_pinned, rho, sig)
        ctx.sync()
        return (time.perf_counter() - t) * 1e3, res

    for _ in range(warmup):
        step(True)
    step(False)

    def timed(resident: bool, k: int):
        launch.barrier()
        torch.cuda.synchronize()
        ctx.msm_stats(1, reset=True)
        ctx.msm_stats(2, reset=True)
        l0 = ctx.launches
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        times, phases = [], []
        for _ in range(k):
            ms, res = step(resident)
            times.append(ms)
            phases.append(res["phases_ms"])
        launch.barrier()
        torch.cuda.synchronize()
        total = launch.max_over_ranks(sum(times))
        return total / k, times, phases, ctx.launches - l0, ctx.msm_stats(1), ctx.msm_stats(2)

    sampler = ClockSampler(ctx.device)
    sampler.start()
    ms_res, times_res, phases, launches, st1, st2 = timed(True, args.steps)
    ms_e2e, times_e2e, phases_e2e, _, _, _ = timed(False, args.steps)
    clocks = sampler.stop()

    # acceptance check outside the timed region (mpc-snarks/src/proof.rs:141 verify_proof): one more proof, verified
    verified = None
    if key_kind == "real":
        _, res = step(True)
        if rank == 0:
            start_el = czk_b200.squaring_chain(start, n_sq)[n_sq:n_sq + 1]
            verified = bool(czk_b200.groth16_verify(czk_b200.pk_verifying_key(pk), start_el, res["proof"], res["proof_inf"]))
    if rank != 0:
        party.close()
        return 0

    avg = lambda key: sum(p[key] for p in phases) / len(phases)
    n_h = D - 1
    msm_h_ms = avg("msm_h")
    # dominant kernel: k_msm_accumulate<Fq> (bucket accumulation), CUDA events on the launching stream
    acc_ms = st1["accumulate_ms"] / max(st1["launches"], 1)
    terms_per_launch = st1["terms"] / max(st1["launches"], 1)
    peaks = {}
    pf = ROOT / "MEASURED_PEAKS.json"
    if pf.exists():
        peaks = json.loads(pf.read_text())
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    alg_bytes = 128.0 * terms_per_launch  # 32 B scalar + 96 B affine base per G1 term
    achieved_gbs = alg_bytes / (acc_ms * 1e-3) / 1e9 if acc_ms else 0.0
    entries_per_launch = st1["entries"] / max(st1["launches"], 1)
    # N*W bucket additions; the batched-affine tree spends 6 products per addition + ~0.5 for the block-wide inversion
    # trees (the XYZZ walk: 8M + 2S = 10); one product = 276 wide multiply-adds (2*12^2 - 12: p0 = 1 rows need none)
    wide_mads = entries_per_launch * 6.5 * 276
    int_rate = wide_mads / (acc_ms * 1e-3) if acc_ms else 0.0
    line = {
        "metric": METRIC, "value": ms_res, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": ms_res, "higher_is_better": False, "scaling": "weak",
        "vs_baseline": (ms_res / PUBLISHED_MS[world]) if (args.log_n == LOG_N and world in PUBLISHED_MS) else None,
        "dtype": "u32 limbs (Montgomery Fr/Fq), integer", "data": "synthetic",
        "config": {"workload": f"groth16 {args.scheme} 2^{args.log_n} constraints (D=2^{D.bit_length() - 1}) BLS12-377, one party per GPU",
                   "parties": world, "l2": "256 MiB flush write between iterations",
                   "bases": ("real CRS generated on the device from seeded toxic waste (czk_groth16_setup); the proof is verified after the timed region"
                             if key_kind == "real" else "device-generated synthetic CRS of the reference's shapes"),
                   "proof_verified": verified,
                   "published_reference_ms": PUBLISHED_MS.get(world)},
        "clocks": clocks,
        "e2e": {"value": ms_e2e, "unit": "ms", "h2d_bytes_per_step": int((n_sq + 1) * 32), "d2h_bytes_per_step": int(2 * 48 * 8 + 6 + 5 * 16 * 192)},
        "gpu_launches": int(launches),
        "roofline": {"kernel": ACC_KERNEL, "bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved_gbs / hbm_peak, "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH,
                     "traffic_source": "profiles/r1_summary.md: ncu --set full dram__bytes_read+write summed over the rounds of one 2^21-1 term "
                                       "accumulation, halved for the 2^20 term MSMs, mean over the step's 4 G1 MSMs; round 0 gathers each base "
                                       "once per window from the precomputed table and every round streams points + prefix products, so "
                                       "traffic >> the 128 B/term algorithmic figure",
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                     "note": "bucket accumulation is bound by integer-instruction dispatch (IMAD.WIDE pipe), not HBM: see roofline_int",
                     "launch_ms": acc_ms, "launches_per_step": st1["launches"] / args.steps, "share_of_step": st1["accumulate_ms"] / args.steps / ms_res},
        "roofline_int": {"kernel": ACC_KERNEL, "bound": "imad", "achieved": int_rate, "peak": imad_peak, "unit": "IMAD.WIDE/s",
                         "frac": int_rate / imad_peak if imad_peak else None,
                         "peak_source": "czk_microbench kind 0 (independent mad.wide.u32 chains), measured at bench start"},
        "g1_msm_adds_per_s": ref_msm_adds(n_h) / (msm_h_ms * 1e-3),
        "g1_msm": {"n": n_h, "ms": msm_h_ms, "adds_ref": ref_msm_adds(n_h)},
        "g2_msm_ms": avg("msm_b_g2"), "msm_device_ms_per_step": {"g1": st1["msm_ms"] / args.steps, "g2": st2["msm_ms"] / args.steps},
        "phases_ms": {k: avg(k) for k in phases[0]},
        "times_ms": {"resident": times_res, "e2e": times_e2e},
    }
    if not args.no_cpu_baseline and world == 1:  # the CPU leg runs on rank 0 of the 1-GPU run only
        line["cpu_baseline"] = cpu_reference_leg(1, 0, args.cpu_sample_log_n)  # one whole 2^20 proof: ~20 s on 16 cores
    print(json.dumps(line))
    party.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())

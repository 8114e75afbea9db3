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
# The dominant kernel is the bucket accumulation of a G1 MSM: the halving rounds k_bat_round<Fq> (batched affine additions)
# + k_bat_finish<Fq>.  It is timed ALONE (one h-query-shaped MSM of 2^21 - 1 terms per launch, L2 flushed between launches)
# by CUDA events on the launching stream inside the library (czk_msm_stats): inside a proof the MSMs overlap each other
# and the witness map, so their in-proof event times are not kernel times.
ACC_KERNEL = "k_bat_round<Fq> x rounds + k_bat_finish<Fq>: bucket accumulation of the 2^21-1-term h-query MSM, timed alone"
# ncu --set full capture of exactly those launches (profiles/r2_ncu_accumulate_g1.csv, tools/r2_ncu_acc.sh): sum over ALL the
# rounds and the finish kernel of dram__bytes_read.sum + dram__bytes_write.sum for one 2^21-1-term accumulation
NCU_TRAFFIC_BYTES_PER_LAUNCH = None  # filled from profiles/r2_traffic.json when it exists (written by tools/make_profiles.py)


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


PLONK_METRIC = "plonk_wiring_proof_ms_2^{}_bls12_377"


def run_plonk_reference(args):
    """BASELINE config 3's data path on the host CPU: the oracle's restatement of Prover::prove_wiring, one party's work."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    from oracle import binding as o

    o.build()
    threads = o.cpu_threads()
    mixed = args.plonk_domain == "mixed"
    D = (3 if mixed else 1) << args.log_n
    dom = f"3*2^{args.log_n} (mixed-radix wire domain of 2^{args.log_n} gates)" if mixed else f"2^{args.log_n}"
    g1, _ = o.generators()
    ks = o.random_fr_mont(3, 2)
    powers = o.G1.gen_progression(g1, ks[0], ks[1], D, threads=threads)
    p, w = o.random_fr_mont(11, D), o.random_fr_mont(12, D)
    times = []
    for i in range(max(0, min(args.warmup, 1)) + max(1, args.steps)):
        t = time.perf_counter()
        res = o.plonk_prove_wiring(o.SCHEME_SPDZ, p[None], w, powers, seed=i, threads=threads)
        dt = (time.perf_counter() - t) * 1e3
        assert res["status"] == 1
        if i >= max(0, min(args.warmup, 1)):
            times.append(dt)
    ms = sum(times) / len(times)
    cb = dict(value=ms, unit="ms", cores=threads, kind="port", log_n=args.log_n, steps_run=len(times),
              sample=f"{len(times)} whole wiring proof(s) of one party over a {dom} domain on {threads} host threads, {ms:.1f} ms each")
    print(json.dumps({"impl": "reference", "metric": PLONK_METRIC.format(args.log_n), "value": ms, "unit": "ms", "n_gpus": args.gpus,
                      "steps": len(times), "warmup": max(0, min(args.warmup, 1)), "ms_per_step": ms, "higher_is_better": False, "scaling": "weak",
                      "vs_baseline": None, "dtype": "u64 limbs (Montgomery), integer", "data": "synthetic",
                      "config": {"workload": f"plonk wiring argument spdz over a {dom} domain BLS12-377, one party's work on the host CPU (oracle port)",
                                 "parties": args.gpus},
                      "cpu_baseline": cb, "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
    return 0


def run_plonk(args):
    """BASELINE config 3 (Plonk SPDZ, 2^18, one party per GPU): the wiring argument - KZG10 commitment / opening MSMs, the
    quotient transforms, batch division, prefix products and Beaver products on shares (mpc-plonk/src/lib.rs:110-258)."""
    import numpy as np
    import torch

    import czk_b200
    from czk_b200 import launch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the czk arm has no CPU fallback (use --impl reference for the CPU path)")
    warmup = max(args.warmup, 3)
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        party = launch.Party()
        party.ctx.batch_open(czk_b200.SCHEME_ADDITIVE, party.ctx.vec(4))
        party.ctx.sync()
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    ctx, rank, world = party.ctx, party.rank, party.world
    scheme = {"spdz": czk_b200.SCHEME_SPDZ, "additive": czk_b200.SCHEME_ADDITIVE, "plain": czk_b200.SCHEME_PLAIN}[args.scheme]
    spdz = args.scheme == "spdz"
    mixed = args.plonk_domain == "mixed"
    D = (3 if mixed else 1) << args.log_n
    dom = f"3*2^{args.log_n} (mixed-radix wire domain of 2^{args.log_n} gates)" if mixed else f"2^{args.log_n}"
    powers = ctx.bases_synthetic(1, 0x377, D, 0).precompute(0)  # committer key: device-generated points of the right shape
    rng = np.random.Generator(np.random.PCG64(0x18))

    def rand_fr(n):
        a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
        a[:, 3] &= np.uint64((1 << 60) - 1)
        return a

    p_plain, w_host = rand_fr(D), rand_fr(D)  # same on every rank (same seed)
    mine = launch.king_share_scatter(p_plain if rank == 0 else None, D, seed=0x5eed) if world > 1 else p_plain
    p_pin = torch.from_numpy(mine.view(np.int64).copy()).pin_memory().numpy().view(np.uint64)
    w_pin = torch.from_numpy(w_host.view(np.int64).copy()).pin_memory().numpy().view(np.uint64)
    p_dev, w_dev = ctx.vec_from(mine), ctx.vec_from(w_host)
    m_dev = ctx.vec_from(mine) if spdz else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{ctx.device}")

    def step(resident, seed):
        flush.zero_()
        torch.cuda.synchronize()
        t = time.perf_counter()
        if resident:
            res = czk_b200.plonk_prove_wiring(ctx, scheme, powers, args.log_n, p_dev, m_dev, w_dev, seed=seed, mixed=mixed)
        else:  # through host buffers: upload this party's shares and the public polynomial, read the proof back
            pv, wv = ctx.vec_from(p_pin), ctx.vec_from(w_pin)
            res = czk_b200.plonk_prove_wiring(ctx, scheme, powers, args.log_n, pv, ctx.vec_from(p_pin) if spdz else None, wv, seed=seed,
                                              mixed=mixed)
        ctx.sync()
        return (time.perf_counter() - t) * 1e3, res

    for i in range(warmup):
        step(True, i)
    step(False, 99)

    def timed(resident):
        launch.barrier()
        torch.cuda.synchronize()
        l0 = ctx.launches
        times, phases = [], []
        for i in range(args.steps):
            ms, res = step(resident, 1000 + i)
            times.append(ms)
            phases.append(res["phases_ms"])
        launch.barrier()
        torch.cuda.synchronize()
        return launch.max_over_ranks(sum(times)) / args.steps, times, phases, ctx.launches - l0

    sampler = ClockSampler(ctx.device)
    sampler.start()
    ms_res, t_res, phases, launches = timed(True)
    ms_e2e, t_e2e, _, _ = timed(False)
    clocks = sampler.stop()
    if rank != 0:
        party.close()
        return 0
    avg = lambda k: sum(p[k] for p in phases) / len(phases)
    line = {"metric": PLONK_METRIC.format(args.log_n), "value": ms_res, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_res, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 limbs (Montgomery Fr/Fq), integer", "data": "synthetic",
            "config": {"workload": f"plonk wiring argument {args.scheme} over a {dom} domain BLS12-377, one party per GPU "
                                   "(4 KZG10 commitments, 9 openings, 16 transforms per component, batch division, prefix products, 2 Beaver products)",
                       "parties": world, "l2": "256 MiB flush write between iterations", "share_transport": ctx.share_transport,
                       "transcript": "stand-in (SplitMix64 over absorbed limbs); the reference's Blake2s/ChaCha transcript plugs in through czk_plonk_transcript"},
            "clocks": clocks,
            "e2e": {"value": ms_e2e, "unit": "ms", "h2d_bytes_per_step": int((3 if spdz else 2) * D * 32), "d2h_bytes_per_step": int(2 * 1200)},
            "gpu_launches": int(launches), "phases_ms": {k: avg(k) for k in phases[0]}, "times_ms": {"resident": t_res, "e2e": t_e2e}}
    print(json.dumps(line))
    party.close()
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
    ap.add_argument("--plonk-domain", default="mixed", choices=["mixed", "radix2"],
                    help="plonk workload: wire domain of 3 * 2^log_n points (the reference's MixedRadixEvaluationDomain for 2^log_n gates) or 2^log_n")
    ap.add_argument("--workload", default="groth16", choices=["groth16", "plonk"],
                    help="groth16: the BASELINE metric's proof (default). plonk: BASELINE config 3's wiring argument (use --log-n 18)")
    args = ap.parse_args()
    if args.workload == "plonk":
        return run_plonk_reference(args) if args.impl == "reference" else run_plonk(args)
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
    t_key = time.perf_counter()
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
    ctx.sync()
    key_prepare_ms = (time.perf_counter() - t_key) * 1e3  # once per key, outside the timed region (like the reference's CRS)
    key_bytes = {"points": 0, "table": 0}
    for i in range(5):
        b = pk.query(i).device_bytes()
        key_bytes["points"] += b["points"]
        key_bytes["table"] += b["table"]
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

    # ---- the dominant kernel on its own: standalone MSMs of the h-query (2^21 - 1 terms) and l-query (2^20 terms) shapes
    def standalone_msm(which, n_terms, reps=5):
        rng = np.random.Generator(np.random.PCG64(0x5ca1a + which))
        sc = rng.integers(0, 1 << 64, size=(n_terms, 4), dtype=np.uint64)
        sc[:, 3] &= np.uint64((1 << 60) - 1)  # < 2^252 < r: valid Montgomery limbs
        dsc = ctx.vec_from(sc)
        q = pk.query(which)
        for _ in range(2):
            ctx.msm_bases(q, dsc)
        ctx.msm_stats(1, reset=True)
        wall = []
        for _ in range(reps):
            flush.zero_()
            torch.cuda.synchronize()
            t = time.perf_counter()
            ctx.msm_bases(q, dsc)
            wall.append((time.perf_counter() - t) * 1e3)
        st = ctx.msm_stats(1)
        return {"n": n_terms, "accumulate_ms": st["accumulate_ms"] / reps, "device_ms": st["msm_ms"] / reps, "wall_ms": sum(wall) / reps,
                "entries": st["entries"] / reps}

    alone_h = standalone_msm(3, D - 1)
    alone_l = standalone_msm(4, n_sq)

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
    acc_ms = alone_h["accumulate_ms"]
    peaks = {}
    pf = ROOT / "MEASURED_PEAKS.json"
    if pf.exists():
        peaks = json.loads(pf.read_text())
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    alg_bytes = 128.0 * n_h  # 32 B scalar + 96 B affine base per G1 term
    achieved_gbs = alg_bytes / (acc_ms * 1e-3) / 1e9 if acc_ms else 0.0
    # N*W bucket additions; the batched-affine tree spends 6 products per addition + ~0.5 for the block-wide inversion
    # trees (the XYZZ walk: 8M + 2S = 10); one product = 276 wide multiply-adds (2*12^2 - 12: p0 = 1 rows need none)
    wide_mads = alone_h["entries"] * 6.5 * 276
    int_rate = wide_mads / (acc_ms * 1e-3) if acc_ms else 0.0
    traffic, traffic_src = None, "no committed ncu capture found (profiles/r2_traffic.json)"
    tf = ROOT / "profiles" / "r2_traffic.json"
    if tf.exists():
        tj = json.loads(tf.read_text())
        traffic, traffic_src = tj.get("accumulate_g1_2_21_bytes"), tj.get("source")
    # the G1 accumulations' share of the step, from kernel-alone times: one h-shaped + three l-shaped MSMs per proof
    share = (alone_h["accumulate_ms"] + 3 * alone_l["accumulate_ms"]) / ms_res
    line = {
        "metric": METRIC, "value": ms_res, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": ms_res, "higher_is_better": False, "scaling": "weak",
        "vs_baseline": (ms_res / PUBLISHED_MS[world]) if (args.log_n == LOG_N and world in PUBLISHED_MS) else None,
        "dtype": "u32 limbs (Montgomery Fr/Fq), integer", "data": "synthetic",
        "config": {"workload": f"groth16 {args.scheme} 2^{args.log_n} constraints (D=2^{D.bit_length() - 1}) BLS12-377, one party per GPU",
                   "parties": world, "l2": "256 MiB flush write between iterations",
                   "share_transport": ctx.share_transport,
                   "bases": ("real CRS generated on the device from seeded toxic waste (czk_groth16_setup); the proof is verified after the timed region"
                             if key_kind == "real" else "device-generated synthetic CRS of the reference's shapes"),
                   "proof_verified": verified,
                   "published_reference_ms": PUBLISHED_MS.get(world)},
        "clocks": clocks,
        "e2e": {"value": ms_e2e, "unit": "ms", "h2d_bytes_per_step": int((n_sq + 1) * 32), "d2h_bytes_per_step": int(2 * 48 * 8 + 6 + 5 * 17 * 192)},
        "gpu_launches": int(launches),
        # per-key, outside the timed region ("CRS upload reported separately"): generating / uploading the key and
        # building the merged-window tables 2^(cw) P_i the MSMs gather from
        "key_prepare_ms": key_prepare_ms, "key_device_bytes": key_bytes,
        "roofline": {"kernel": ACC_KERNEL, "bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved_gbs / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst figure: the kernel is timed alone)" if peaks else "fallback 6650 GB/s",
                     "note": "bucket accumulation is bound by integer-instruction dispatch (IMAD.WIDE pipe), not HBM: see roofline_int",
                     "launch_ms": acc_ms, "launches_per_step": 4, "share_of_step": share,
                     "share_note": "kernel-alone accumulation times of the step's four G1 MSMs (1 x 2^21-1 + 3 x 2^20 terms) / ms_per_step; "
                                   "inside the step they overlap the G2 MSM and the witness map"},
        "roofline_int": {"kernel": ACC_KERNEL, "bound": "imad", "achieved": int_rate, "peak": imad_peak, "unit": "IMAD.WIDE/s",
                         "frac": int_rate / imad_peak if imad_peak else None,
                         "peak_source": "czk_microbench kind 0 (independent mad.wide.u32 chains), measured at bench start"},
        "g1_msm_adds_per_s": ref_msm_adds(n_h) / (alone_h["wall_ms"] * 1e-3),
        "g1_msm": {"n": n_h, "ms": alone_h["wall_ms"], "device_ms": alone_h["device_ms"], "accumulate_ms": acc_ms, "adds_ref": ref_msm_adds(n_h),
                   "how": "standalone czk_msm_bases on the resident h-query (host call to result), L2 flushed between calls"},
        "g1_msm_2_20": alone_l,
        "phases_ms": {k: avg(k) for k in phases[0]},
        "phases_note": "witness_map and msm_* are device times of jobs that OVERLAP (context stream + two MSM lanes): they do not add up",
        "times_ms": {"resident": times_res, "e2e": times_e2e},
    }
    if not args.no_cpu_baseline and world == 1:  # the CPU leg runs on rank 0 of the 1-GPU run only
        line["cpu_baseline"] = cpu_reference_leg(1, 0, args.cpu_sample_log_n)  # one whole 2^20 proof: ~20 s on 16 cores
    print(json.dumps(line))
    party.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())

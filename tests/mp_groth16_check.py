#!/usr/bin/env python3
"""Multi-party parity run: `torchrun --nproc-per-node N tests/mp_groth16_check.py [--scheme spdz] [--log-n 10]`.
Each rank is one party on its own GPU (NCCL opens through libczk_b200); rank p compares ITS h share, proof share
and the revealed proof with the oracle's in-process simulation of all N parties on the same shares.
Exit code 0 on parity.  (Test infrastructure: uses oracle/.)"""
import argparse
import random
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np

import czk_b200
from czk_b200 import launch
from oracle import binding as o
from oracle import pymodel as m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scheme", default="spdz", choices=["additive", "spdz", "gsz"])
    ap.add_argument("--log-n", type=int, default=10)
    ap.add_argument("--squarings", type=int, default=0, help="exact number of squarings (overrides --log-n)")
    ap.add_argument("--corrupt-mac", action="store_true", help="negative case: rank 1 corrupts one MAC share, every rank must see the check fail")
    args = ap.parse_args()
    party = launch.Party()
    ctx, rank, world = party.ctx, party.rank, party.world
    scheme = czk_b200.SCHEME_SPDZ if args.scheme == "spdz" else czk_b200.SCHEME_ADDITIVE
    n_sq = args.squarings or (1 << args.log_n)
    if args.scheme == "gsz":
        return main_gsz(party, n_sq)
    if args.corrupt_mac:
        return main_corrupt(party)
    rnd = random.Random(1234 + n_sq)
    toxic = [rnd.randrange(1, m.R_MOD) for _ in range(7)]
    pk = o.groth16_setup(n_sq, o.fr_from_ints(toxic), threads=max(1, o.cpu_threads() // world))
    chain = o.squaring_chain(o.fr_from_ints([rnd.randrange(m.R_MOD)])[0], n_sq)
    # every rank derives the same sharing (the king's scatter is exercised separately below)
    shares = czk_b200.king_share_batch(chain, world, seed=99)
    mine = launch.king_share_scatter(chain if rank == 0 else None, n_sq + 1, seed=99)
    assert (mine == shares[rank]).all(), "king scatter delivered the wrong slice"
    rho, sigma = 1234567, 7654321  # MpcField::rand: same seeded value at every party
    r_sh = o.fr_from_ints([rho] * world)
    s_sh = o.fr_from_ints([sigma] * world)
    exp = o.groth16_prove(scheme, n_sq, list(shares), r_sh, s_sh, pk, threads=max(1, o.cpu_threads() // world))
    assert exp["ok"]
    dpk = czk_b200.ProvingKey.upload(ctx, pk)
    h = czk_b200.groth16_witness_map(ctx, scheme, n_sq, mine)
    assert (h == exp["h"][rank]).all(), f"rank {rank}: h share differs"
    got = czk_b200.groth16_prove(ctx, scheme, dpk, mine, r_sh[rank], s_sh[rank])
    assert (got["proof"] == exp["proof"]).all() and (got["proof_inf"] == exp["proof_inf"]).all(), f"rank {rank}: revealed proof differs"
    assert (got["proof_sh"] == exp["proof_sh"][rank]).all(), f"rank {rank}: proof share differs"
    assert czk_b200.groth16_verify(pk, chain[n_sq:n_sq + 1], got["proof"], got["proof_inf"]), f"rank {rank}: revealed proof does not verify"
    st = ctx.net_stats()
    assert st["broadcasts"] > 0 and st["bytes_sent"] > 0
    # batch_open parity on a random shared vector
    x = o.random_fr_mont(5, 1000)
    xs = czk_b200.king_share_batch(x, world, seed=3)
    opened = ctx.batch_open(scheme, ctx.vec_from(xs[rank]), ctx.vec_from(xs[rank]) if scheme == czk_b200.SCHEME_SPDZ else None)
    assert (opened.numpy() == x).all()
    # any circuit: a random R1CS through czk_groth16_prove_r1cs, every rank holding its own shares of the assignment
    cs, z = o.random_r1cs(seed=77, n_inst=3, n_free=9, n_cons=200, modulus=m.R_MOD)
    pk2 = o.groth16_setup_r1cs(cs, o.fr_from_ints(toxic), threads=max(1, o.cpu_threads() // world))
    oscheme2 = o.SCHEME_SPDZ if scheme == czk_b200.SCHEME_SPDZ else o.SCHEME_ADDITIVE
    full = o.r1cs_full_shares(z, world, seed=8, scheme=oscheme2)
    exp2 = o.groth16_prove_r1cs(oscheme2, cs, full, r_sh, s_sh, pk2)
    assert exp2["ok"]
    dcs, dpk2 = czk_b200.R1cs(ctx, cs), czk_b200.groth16_pk_upload_r1cs(ctx, pk2)
    got2 = czk_b200.groth16_prove_r1cs(ctx, scheme, dpk2, dcs, full[rank], r_sh[rank], s_sh[rank])
    assert (got2["proof"] == exp2["proof"]).all(), f"rank {rank}: generic-circuit proof differs"
    assert (got2["proof_sh"] == exp2["proof_sh"][rank]).all(), f"rank {rank}: generic-circuit proof share differs"
    # Plonk leaves on real shares across the ranks: batch_inv, batch_div, partial_products
    spdz = scheme == czk_b200.SCHEME_SPDZ
    k = 777
    xv, yv = o.random_fr_mont(31, k), o.random_fr_mont(32, k)
    xsh, ysh = czk_b200.king_share_batch(xv, world, seed=41), czk_b200.king_share_batch(yv, world, seed=42)
    oscheme = o.SCHEME_SPDZ if spdz else o.SCHEME_ADDITIVE
    for op, name in ((o.SHARE_BATCH_INV, "batch_inv"), (o.SHARE_PARTIAL_PRODUCTS, "partial_products"), (o.SHARE_BATCH_DIV, "batch_div")):
        xs, xm = ctx.vec_from(xsh[rank]), (ctx.vec_from(xsh[rank]) if spdz else None)
        ys, ym = ctx.vec_from(ysh[rank]), (ctx.vec_from(ysh[rank]) if spdz else None)
        if op == o.SHARE_BATCH_INV:
            ctx.share_batch_inv(scheme, xs, xm)
        elif op == o.SHARE_PARTIAL_PRODUCTS:
            ctx.share_partial_products(scheme, xs, xm)
        else:
            ctx.share_batch_div(scheme, xs, xm, ys, ym)
        stt, exp_sh, exp_mac = o.share_op(op, oscheme, xsh, xsh if spdz else None, ysh, ysh if spdz else None)
        assert stt == 1 and (xs.numpy() == exp_sh[rank]).all(), f"rank {rank}: {name} share differs"
        if spdz:
            assert (xm.numpy() == exp_mac[rank]).all(), f"rank {rank}: {name} MAC share differs"
    # Plonk wiring argument across the ranks: revealed proof and THIS rank's opening-proof shares vs the oracle's simulation
    sys.path.insert(0, str(ROOT / "tests"))
    from helpers import kzg_powers, plonk_wiring_instance

    # twice: a power-of-two domain, and the reference's own wire-domain shape (3 * 2^k points, mixed radix)
    for log_d, mixed in ((8, False), (6, True)):
        D = (3 if mixed else 1) << log_d
        powers = kzg_powers(D, 0xfeed + mixed)
        pp, ww = plonk_wiring_instance(None, seed=4 + mixed, size=D)
        psh = czk_b200.king_share_batch(pp, world, seed=17)
        expp = o.plonk_prove_wiring(oscheme, psh, ww, powers, seed=5, threads=max(1, o.cpu_threads() // world))
        assert expp["status"] == 1
        kb = ctx.bases_upload(1, powers)
        gotp = czk_b200.plonk_prove_wiring(ctx, scheme, kb, log_d, ctx.vec_from(psh[rank]), ctx.vec_from(psh[rank]) if spdz else None,
                                           ctx.vec_from(ww), seed=5, mixed=mixed)
        for key in expp["proof"]:
            assert (gotp["proof"][key] == expp["proof"][key]).all(), f"rank {rank}: plonk wiring proof differs at {key} (mixed={mixed})"
        assert (gotp["proof_share"]["open_pf_xy"] == expp["share_pf_xy"][rank]).all(), f"rank {rank}: plonk opening-proof shares differ"
        assert (gotp["proof_share"]["open_pf_inf"] == expp["share_pf_inf"][rank]).all()
        kb.free()
    launch.barrier()
    print(f"[rank {rank}/{world}] groth16 {args.scheme} n={n_sq}: parity ok; net {st}; opens over {ctx.share_transport}; link bytes {ctx.net_link_bytes()}", flush=True)
    party.close()


def main_corrupt(party):
    """SPDZ open and product with one corrupted MAC share at rank 1: the slice owner's flag is all-gathered, so the
    call must fail with CZK_ERR_PROTOCOL on EVERY rank (spdz.rs:182 assert!s at every party), and the next call must work."""
    ctx, rank, world = party.ctx, party.rank, party.world
    k = 1000
    x, y = o.random_fr_mont(5, k), o.random_fr_mont(6, k)
    xs, ys = czk_b200.king_share_batch(x, world, seed=3), czk_b200.king_share_batch(y, world, seed=4)
    mac = xs[rank].copy()
    if rank == 1:
        mac[k - 1] = o.random_fr_mont(7, 1)[0]
    for what in ("open", "mul"):
        try:
            if what == "open":
                ctx.batch_open(czk_b200.SCHEME_SPDZ, ctx.vec_from(xs[rank]), ctx.vec_from(mac))
            else:
                ctx.batch_mul(czk_b200.SCHEME_SPDZ, ctx.vec_from(xs[rank]), ctx.vec_from(mac), ctx.vec_from(ys[rank]), ctx.vec_from(ys[rank]))
            raise AssertionError(f"rank {rank}: {what} accepted a corrupted MAC share")
        except czk_b200.CzkError as e:
            assert e.code == 5, e
    opened = ctx.batch_open(czk_b200.SCHEME_SPDZ, ctx.vec_from(xs[rank]), ctx.vec_from(xs[rank]))
    assert (opened.numpy() == x).all()
    launch.barrier()
    print(f"[rank {rank}/{world}] corrupted MAC detected by open and mul; clean open ok", flush=True)
    party.close()


def main_gsz(party, n_sq):
    """GSZ20: every party holds the plaintext under the reference's stubs; real Shamir shares are exercised on
    czk_gsz_open / czk_gsz_king_compute (open, degree check, failing degree check)."""
    ctx, rank, world = party.ctx, party.rank, party.world
    rnd = random.Random(4321 + n_sq)
    toxic = [rnd.randrange(1, m.R_MOD) for _ in range(7)]
    threads = max(1, o.cpu_threads() // world)
    pk = o.groth16_setup(n_sq, o.fr_from_ints(toxic), threads=threads)
    chain = o.squaring_chain(o.fr_from_ints([rnd.randrange(m.R_MOD)])[0], n_sq)
    r, s = o.fr_from_ints([rnd.randrange(m.R_MOD)]), o.fr_from_ints([rnd.randrange(m.R_MOD)])
    exp = o.groth16_prove_gsz(world, n_sq, chain, r[0], s[0], pk, threads=threads)
    assert exp["ok"]
    dpk = czk_b200.ProvingKey.upload(ctx, pk)
    before = ctx.gsz_stats()
    got = czk_b200.groth16_prove(ctx, czk_b200.SCHEME_GSZ, dpk, chain, r[0], s[0])
    assert (got["proof"] == exp["proof"]).all() and (got["proof_inf"] == exp["proof_inf"]).all(), f"rank {rank}: proof differs"
    assert (got["field_check"] == exp["field_check"]).all(), f"rank {rank}: field product check values differ"
    assert (got["group_check_x"] == exp["group_check_x"]).all() and (got["group_check_yz"] == exp["group_check_yz"]).all()
    assert got["king_computes"] - before["king_computes"] == exp["king_computes"] and got["opens"] - before["opens"] == exp["opens"]
    # real Shamir shares: k random degree-t polynomials, party j holds p_i(w^j)
    t = (world - 1) // 2
    k = 257
    coeffs = [o.random_fr_mont(900 + i, t + 1) for i in range(k)]
    shares = np.stack([o.gsz_share(world, c) for c in coeffs], axis=1)  # (world, k, 4)
    secrets = np.stack([c[0] for c in coeffs])
    opened = ctx.gsz_open(ctx.vec_from(shares[rank]), t)
    assert (opened.numpy() == secrets).all(), f"rank {rank}: Shamir open differs"
    v = ctx.vec_from(shares[rank])
    ctx.gsz_king_compute(v, t)
    assert (v.numpy() == secrets).all()
    if t >= 1:
        # a product of two degree-t sharings has degree 2t: opens at 2t, must FAIL the degree-t check
        prod = o.fr_mul(shares[rank], shares[rank])
        assert (ctx.gsz_open(ctx.vec_from(prod), 2 * t).numpy() == o.fr_mul(secrets, secrets)).all()
        try:
            ctx.gsz_open(ctx.vec_from(prod), t)
            raise AssertionError("degree check did not fire")
        except czk_b200.CzkError as e:
            assert e.code == 5
    st = ctx.net_stats()
    assert world == 1 or (st["to_king"] > 0 and st["from_king"] > 0)
    launch.barrier()
    print(f"[rank {rank}/{world}] groth16 gsz n={n_sq} t={t}: parity ok; {ctx.gsz_stats()} net {st}", flush=True)
    party.close()


if __name__ == "__main__":
    main()

"""GPU parity for share opening and Beaver multiplication with one party (n_parties = 1): the
collective degenerates to a copy, the arithmetic (sum, sigma, MAC check, Beaver finish) is the same.
mpc-algebra/src/share/{add.rs:121-125, spdz.rs:166-185, field.rs:97-127}."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_single_party_open_and_beaver(ctx, czk, oracle):
    ctx.net_init(0, 1, None)
    n = 3000
    x = oracle.random_fr_mont(41, n)
    y = oracle.random_fr_mont(42, n)
    for scheme in (czk.SCHEME_PLAIN, czk.SCHEME_ADDITIVE, czk.SCHEME_SPDZ):
        xs, ys = ctx.vec_from(x), ctx.vec_from(y)
        xm, ym = (ctx.vec_from(x), ctx.vec_from(y)) if scheme == czk.SCHEME_SPDZ else (None, None)
        opened = ctx.batch_open(scheme, xs, xm)
        assert (opened.numpy() == x).all()
        ctx.batch_mul(scheme, xs, xm, ys, ym)
        exp = oracle.fr_mul(x, y)
        assert (xs.numpy() == exp).all(), scheme
        if xm is not None:
            assert (xm.numpy() == exp).all()


def test_spdz_mac_check_detects_tampering(ctx, czk, oracle):
    ctx.net_init(0, 1, None)
    n = 100
    x = oracle.random_fr_mont(43, n)
    bad = x.copy()
    bad[17] = oracle.random_fr_mont(44, 1)[0]
    with pytest.raises(czk.CzkError) as e:
        ctx.batch_open(czk.SCHEME_SPDZ, ctx.vec_from(x), ctx.vec_from(bad))
    assert e.value.code == 5  # CZK_ERR_PROTOCOL


# ---------------------------------------------------------------------------------------------------------------------
# N-party arithmetic on ONE GPU.  czk_diag_sim_* runs the kernels of czk_batch_open / czk_beaver_batch_mul (pack, slice
# reduce, sigma + Beaver finish, zero check) for every simulated party with that party's constants; only the NCCL
# transport is replaced by direct addressing.  Expected values: the oracle's in-process N-party simulation
# (oracle/czk_oracle_plonk.inc orc_share_op: share/field.rs:97-127 on additive / SPDZ shares).
def _party_vecs(ctx, arr):
    return [ctx.vec_from(arr[q]) for q in range(arr.shape[0])]


@pytest.mark.parametrize("parties,k", [(2, 1 << 10), (3, 1000), (4, 1 << 10), (8, 1 << 10), (5, 77), (2, 1 << 21), (8, 1 << 21)])
@pytest.mark.parametrize("scheme_name", ["additive", "spdz"])
def test_n_party_open_and_beaver_simulated(ctx, czk, oracle, parties, k, scheme_name):
    if k >= (1 << 21) and scheme_name == "additive":
        pytest.skip("the 2^21 case runs under SPDZ (a superset of the additive kernels)")
    spdz = scheme_name == "spdz"
    scheme = czk.SCHEME_SPDZ if spdz else czk.SCHEME_ADDITIVE
    oscheme = oracle.SCHEME_SPDZ if spdz else oracle.SCHEME_ADDITIVE
    x, y = oracle.random_fr_mont(1000 + parties, k), oracle.random_fr_mont(2000 + parties, k)
    xsh, ysh = czk.king_share_batch(x, parties, seed=11), czk.king_share_batch(y, parties, seed=12)
    # batch_open: every party ends with the plain values, no MAC flag
    opened, flags = ctx.sim_batch_open(scheme, _party_vecs(ctx, xsh), _party_vecs(ctx, xsh) if spdz else None)
    assert flags == [0] * parties
    for q in range(parties):
        assert (opened[q].numpy() == x).all(), f"party {q}: opened values differ"
    # batch_mul: every party's share (and MAC share) bit-identical to the oracle's N-party simulation
    st, exp_sh, exp_mac = oracle.share_op(oracle.SHARE_BATCH_MUL, oscheme, xsh, xsh if spdz else None, ysh, ysh if spdz else None,
                                          threads=oracle.cpu_threads())
    assert st == 1
    xs, ys = _party_vecs(ctx, xsh), _party_vecs(ctx, ysh)
    xm, ym = (_party_vecs(ctx, xsh), _party_vecs(ctx, ysh)) if spdz else (None, None)
    flags = ctx.sim_batch_mul(scheme, xs, xm, ys, ym)
    assert flags == [0] * parties
    total = np.zeros_like(x)
    for q in range(parties):
        got = xs[q].numpy()
        assert (got == exp_sh[q]).all(), f"party {q}: product share differs"
        if spdz:
            assert (xm[q].numpy() == exp_mac[q]).all(), f"party {q}: product MAC share differs"
        total = oracle.fr_add(total, got)
    assert (total == oracle.fr_mul(x, y)).all(), "the product shares do not reconstruct x * y"


@pytest.mark.parametrize("parties", [2, 3, 8])
def test_n_party_spdz_mac_check_fires_in_the_right_slice(ctx, czk, oracle, parties):
    k = 1003
    x, y = oracle.random_fr_mont(31, k), oracle.random_fr_mont(32, k)
    xsh, ysh = czk.king_share_batch(x, parties, seed=5), czk.king_share_batch(y, parties, seed=6)
    # open: corrupt one MAC share of party 1 at element `bad`: only the party that checks that slice may raise its flag
    m = (k + parties - 1) // parties
    for bad in (0, k // 2, k - 1):
        mac = xsh.copy()
        mac[1, bad] = oracle.random_fr_mont(77, 1)[0]
        _, flags = ctx.sim_batch_open(czk.SCHEME_SPDZ, _party_vecs(ctx, xsh), _party_vecs(ctx, mac))
        exp = [1 if q == bad // m else 0 for q in range(parties)]
        assert flags == exp, (bad, flags)
    # product: 2k opened elements (s + x | o + y); corrupt y's MAC: element k + bad of the concatenated open
    m2 = (2 * k + parties - 1) // parties
    bad = 17
    ymac = ysh.copy()
    ymac[0, bad] = oracle.random_fr_mont(78, 1)[0]
    flags = ctx.sim_batch_mul(czk.SCHEME_SPDZ, _party_vecs(ctx, xsh), _party_vecs(ctx, xsh), _party_vecs(ctx, ysh), _party_vecs(ctx, ymac))
    assert flags == [1 if q == (k + bad) // m2 else 0 for q in range(parties)], flags


def test_net_stats_follow_mpc_net_accounting(ctx, czk, oracle):
    """One party: no peers, so bytes stay 0, but the broadcast COUNT is mpc-net's (an SPDZ open = broadcast +
    atomic_broadcast = 3; mpc-net/src/multi.rs:145-174, mpc-algebra/src/channel.rs:50-75)."""
    ctx.net_init(0, 1, None)
    ctx.net_reset_stats()
    x = oracle.random_fr_mont(9, 64)
    ctx.batch_open(czk.SCHEME_SPDZ, ctx.vec_from(x), ctx.vec_from(x))
    assert ctx.net_stats()["broadcasts"] == 3
    ctx.batch_open(czk.SCHEME_ADDITIVE, ctx.vec_from(x))
    assert ctx.net_stats()["broadcasts"] == 4
    ctx.batch_mul(czk.SCHEME_SPDZ, ctx.vec_from(x), ctx.vec_from(x), ctx.vec_from(x), ctx.vec_from(x))
    st = ctx.net_stats()
    assert st["broadcasts"] == 10 and st["bytes_sent"] == 0
    assert ctx.net_link_bytes() == {"sent": 0, "received": 0}

"""GPU parity for the Plonk / KZG10 leaves: prefix products, batch inversion, division by (X - z),
the share protocols batch_inv / batch_div / partial_products and KZG10 open, against the oracle's restatement of
mpc-algebra/src/share/field.rs:135-182, poly/src/polynomial/univariate/mod.rs:133-174 and poly-commit/src/kzg10/mod.rs.
One party here; tests/mp_groth16_check.py --scheme {additive,spdz} exercises the same protocols across ranks."""
import numpy as np
import pytest

from helpers import jac_to_affine_ints, make_points

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 2, 7, 2048, 2049, 5000, (1 << 16) + 3, 1 << 21])  # tile edges of the scan; C3 / C2 sizes
def test_prefix_products_match_serial_loop(ctx, oracle, n):
    x = oracle.random_fr_mont(40 + n % 97, n)
    v = ctx.vec_from(x)
    ctx.prefix_products(v)
    st, exp, _ = oracle.share_op(oracle.SHARE_PARTIAL_PRODUCTS, oracle.SCHEME_PLAIN, x[None, :, :])
    assert st == 1 and (v.numpy() == exp[0]).all()


@pytest.mark.parametrize("n", [1, 15, 16, 17, 1000, 1 << 18])
def test_batch_inverse_matches_oracle_and_rejects_zero(ctx, czk, oracle, n):
    x = oracle.random_fr_mont(60 + n % 89, n)
    v = ctx.vec_from(x)
    ctx.batch_inverse(v)
    st, exp, _ = oracle.share_op(oracle.SHARE_BATCH_INV, oracle.SCHEME_PLAIN, x[None, :, :])
    assert st == 1 and (v.numpy() == exp[0]).all()
    x[n // 2] = 0
    with pytest.raises(czk.CzkError) as e:
        ctx.batch_inverse(ctx.vec_from(x))
    assert e.value.code == 5


@pytest.mark.parametrize("n", [1, 2, 3, 1024, 2048, 4097, 1 << 18])
def test_poly_div_linear_matches_long_division(ctx, oracle, pymodel, n):
    p = oracle.random_fr_mont(80 + n % 83, n)
    for z in (oracle.random_fr_mont(81, 1)[0], oracle.fr_from_ints([0])[0], oracle.fr_from_ints([1])[0],
              oracle.fr_from_ints([pymodel.R_MOD - 1])[0]):
        q, rem = ctx.poly_div_linear(ctx.vec_from(p), z)
        eq, erem = oracle.poly_div_linear(p, z)
        assert (rem == erem).all() and (rem == oracle.poly_eval(p, z)).all()
        if n > 1:
            assert (q.numpy(n=n - 1) == eq).all()


@pytest.mark.parametrize("scheme_name", ["plain", "additive", "spdz"])
def test_share_protocols_single_party(ctx, czk, oracle, scheme_name):
    ctx.net_init(0, 1, None)
    scheme = {"plain": czk.SCHEME_PLAIN, "additive": czk.SCHEME_ADDITIVE, "spdz": czk.SCHEME_SPDZ}[scheme_name]
    oscheme = {"plain": oracle.SCHEME_PLAIN, "additive": oracle.SCHEME_ADDITIVE, "spdz": oracle.SCHEME_SPDZ}[scheme_name]
    spdz = scheme == czk.SCHEME_SPDZ
    n = 3000
    x, y = oracle.random_fr_mont(91, n), oracle.random_fr_mont(92, n)
    for op in (oracle.SHARE_BATCH_INV, oracle.SHARE_PARTIAL_PRODUCTS, oracle.SHARE_BATCH_DIV):
        xs, xm = ctx.vec_from(x), (ctx.vec_from(x) if spdz else None)
        ys, ym = ctx.vec_from(y), (ctx.vec_from(y) if spdz else None)
        if op == oracle.SHARE_BATCH_INV:
            ctx.share_batch_inv(scheme, xs, xm)
        elif op == oracle.SHARE_PARTIAL_PRODUCTS:
            ctx.share_partial_products(scheme, xs, xm)
        else:
            ctx.share_batch_div(scheme, xs, xm, ys, ym)
        st, exp, expm = oracle.share_op(op, oscheme, x[None], x[None] if spdz else None, y[None], y[None] if spdz else None)
        assert st == 1
        assert (xs.numpy() == exp[0]).all(), (scheme_name, op)
        if spdz:
            assert (xm.numpy() == expm[0]).all(), (scheme_name, op, "mac")


def test_kzg_commit_and_open_match_oracle(ctx, oracle, pymodel):
    """KZG10 commit = MSM over powers_of_g; open = evaluation + MSM of the witness polynomial (kzg10/mod.rs:141-262).
    The identity the verifier checks, in the exponent: commit - eval * G == (tau - z) * w for powers tau^i G."""
    G = oracle.G1
    n = 1 << 10
    tau = 0x1234567
    g1, _ = oracle.generators()
    # powers_of_g = tau^i * G: a geometric progression is not what gen_progression builds, so use scalar_mul per power
    pts = []
    acc = 1
    for i in range(n):
        pts.append(acc)
        acc = acc * tau % pymodel.R_MOD
    powers = np.stack([G.scalar_mul(g1, oracle.fr_from_ints([k])[0])[0] for k in pts])
    bases = ctx.bases_upload(1, powers, None)
    p = oracle.random_fr_mont(101, n)
    z = oracle.random_fr_mont(102, 1)[0]
    dp = ctx.vec_from(p)
    commit = jac_to_affine_ints(G, ctx.msm_bases(bases, dp))
    exp_commit = G.msm(powers, None, p, threads=4)
    assert commit == (None if exp_commit[1] else G.affine_to_ints(exp_commit[0])[0])
    w, ev = ctx.kzg_open(bases, dp, z)
    ew, ewinf, eev = oracle.kzg_open(powers, None, p, z, threads=4)
    assert (ev == eev).all()
    assert jac_to_affine_ints(G, w) == (None if ewinf else G.affine_to_ints(ew[None, :])[0])
    # commit - eval*G == (tau - z) * w
    evi, zi = oracle.fr_to_ints(ev[None, :])[0], oracle.fr_to_ints(z[None, :])[0]
    gen = G.affine_to_ints(g1[None, :])[0]
    lhs = pymodel.g1_add(commit, pymodel.g1_neg(pymodel.g1_mul(gen, evi)))
    rhs = pymodel.g1_mul(jac_to_affine_ints(G, w), (tau - zi) % pymodel.R_MOD)
    assert lhs == rhs

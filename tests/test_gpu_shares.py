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

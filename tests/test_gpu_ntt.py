"""GPU parity: device NTT (czk_ntt_fr / czk_ntt_vec) vs the CPU oracle, bit-exact.

Mirrors the reference's own NTT tests (algebra/poly/src/domain/radix2/mod.rs:320-360, :381-491):
FFT == polynomial evaluation at w^i / g w^i, iFFT o FFT == id, and all four transform flavours.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("log_d", [0, 1, 2, 3, 5, 8, 10, 11, 12, 13, 16, 18])
def test_ntt_matches_oracle_all_flavours(ctx, oracle, log_d):
    n = 1 << log_d
    v = oracle.random_fr_mont(0x47 + log_d, n)
    for inverse in (False, True):
        for coset in (False, True):
            got = ctx._ntt_host(v, inverse, coset)
            exp = oracle.ntt(v, inverse, coset, threads=8) if n > 1 else v.copy()
            assert got.shape == exp.shape
            assert (got == exp).all(), f"log_d={log_d} inverse={inverse} coset={coset}: first diff at {np.argwhere(got != exp)[:3]}"


def test_ntt_zero_pads_short_input(ctx, oracle):
    # radix2/mod.rs:100-101: coeffs.resize(size, zero)
    v = oracle.random_fr_mont(9, 1000)
    padded = np.concatenate([v, np.zeros((24, 4), np.uint64)])
    assert (ctx.fft(v) == oracle.ntt(padded)).all()


def test_fft_is_polynomial_evaluation(ctx, oracle, pymodel):
    # radix2/mod.rs:320-360 test_fft_correctness, on BLS12-377 Fr
    log_d = 6
    n = 1 << log_d
    coeffs = oracle.random_fr_mont(77, n)
    d = oracle.domain_params(n)
    evals = ctx.fft(coeffs)
    cevals = ctx.coset_fft(coeffs)
    g = oracle.fr_from_ints([pymodel.FR_GENERATOR])[0]
    for i in (0, 1, 2, 17, n - 1):
        wi = oracle.fr_pow_u64(d["group_gen"], i)
        assert (evals[i] == oracle.poly_eval(coeffs, wi)).all()
        assert (cevals[i] == oracle.poly_eval(coeffs, oracle.fr_mul(g[None, :], wi[None, :])[0])).all()


@pytest.mark.parametrize("log_d", [20, 21, 22])
def test_ntt_large_roundtrip_and_spot_parity(ctx, oracle, log_d):
    """BASELINE sizes: bit-exact vs the oracle (multi-threaded) and the size-independent identities."""
    n = 1 << log_d
    v = oracle.random_fr_mont(0x47 + log_d, n)
    dv = ctx.vec_from(v)
    ctx.ntt_in_place(dv, log_d, inverse=False, coset=True)
    fwd = dv.numpy()
    assert (fwd == oracle.ntt(v, False, True, threads=oracle.cpu_threads())).all()
    ctx.ntt_in_place(dv, log_d, inverse=True, coset=True)
    assert (dv.numpy() == v).all()
    # linearity: NTT(a + b) == NTT(a) + NTT(b)
    w = oracle.random_fr_mont(0x1234 + log_d, n)
    da, db = ctx.vec_from(v), ctx.vec_from(w)
    ds = ctx.vec_from(oracle.fr_add(v, w))
    for x in (da, db, ds):
        ctx.ntt_in_place(x, log_d)
    ctx.vec_add(da, db)
    assert (da.numpy() == ds.numpy()).all()


def test_ntt_2_24_properties(ctx, oracle, pymodel):
    """The top of BASELINE config 5's sweep (2^24): round trips of all flavours, linearity, and FFT(p)[i] == p(w^i)
    at a few indices by Horner (the reference's own identity, radix2/mod.rs:320-360) - size-independent checks."""
    log_d = 24
    n = 1 << log_d
    v = oracle.random_fr_mont(0x24, n)
    dv = ctx.vec_from(v)
    ctx.ntt_in_place(dv, log_d)
    fwd = dv.numpy()
    gen = oracle.fr_to_ints(oracle.domain_params(n)["group_gen"][None, :])[0]
    for i in (0, 1, 0x9a5b3c % n, n - 1):
        x = oracle.fr_from_ints([pow(gen, i, pymodel.R_MOD)])[0]
        assert (fwd[i] == oracle.poly_eval(v, x)).all(), i
    ctx.ntt_in_place(dv, log_d, inverse=True)
    assert (dv.numpy() == v).all()
    ctx.ntt_in_place(dv, log_d, inverse=False, coset=True)
    ctx.ntt_in_place(dv, log_d, inverse=True, coset=True)
    assert (dv.numpy() == v).all()
    w = oracle.random_fr_mont(0x25, n)
    da, ds = ctx.vec_from(w), ctx.vec_from(oracle.fr_add(v, w))
    ctx.ntt_in_place(da, log_d)
    ctx.ntt_in_place(ds, log_d)
    ctx.vec_add(da, ctx.vec_from(fwd))
    assert (da.numpy() == ds.numpy()).all()


def test_domain_params_match_oracle(czk, oracle):
    for log_d in (0, 1, 4, 11, 21, 24, 30):
        a = czk.domain_params(log_d)
        b = oracle.domain_params(1 << log_d)
        for k in ("group_gen", "group_gen_inv", "size_inv", "generator_inv"):
            assert (a[k] == b[k]).all(), (log_d, k)


def test_pointwise_helpers(ctx, oracle, pymodel):
    n = 5000
    a = oracle.random_fr_mont(1, n)
    b = oracle.random_fr_mont(2, n)
    for name, ref in (("vec_add", oracle.fr_add), ("vec_sub", oracle.fr_sub), ("vec_mul", oracle.fr_mul)):
        da, db = ctx.vec_from(a), ctx.vec_from(b)
        getattr(ctx, name)(da, db)
        assert (da.numpy() == ref(a, b)).all(), name
    c = oracle.random_fr_mont(3, 1)[0]
    da = ctx.vec_from(a)
    ctx.vec_scale(da, c)
    assert (da.numpy() == oracle.fr_mul(a, np.tile(c, (n, 1)))).all()
    # distribute_powers (domain/mod.rs:93-104): a[i] *= c * g^i
    g = oracle.fr_from_ints([pymodel.FR_GENERATOR])[0]
    da = ctx.vec_from(a)
    ctx.distribute_powers(da, g, c)
    ai = oracle.fr_to_ints(a)
    ci = oracle.fr_to_ints(c[None, :])[0]
    exp = [x * ci * pow(pymodel.FR_GENERATOR, i, pymodel.R_MOD) % pymodel.R_MOD for i, x in enumerate(ai)]
    assert oracle.fr_to_ints(da.numpy()) == exp
    # divide_by_vanishing_poly_on_coset_in_place
    log_d = 12
    x = oracle.random_fr_mont(4, 1 << log_d)
    dx = ctx.vec_from(x)
    ctx.divide_by_vanishing_on_coset(dx, log_d)
    assert (dx.numpy() == oracle.divide_by_vanishing_on_coset(x)).all()


@pytest.mark.parametrize("log_d", [1, 2, 3, 4, 9, 11, 12, 14, 17, 20, 21])
def test_ntt_batch_and_fused_pair_match_oracle(ctx, czk, oracle, log_d):
    """czk_ntt_vec_batch: several vectors per grid, all five ops.  CZK_NTT_IFFT_COSET_FFT (the witness map's transform
    pair, r1cs_to_qap.rs:85-90, run as inverse DIF -> forward DIT with the scaling fused) must equal ifft followed by
    coset_fft of the oracle bit for bit."""
    n = 1 << log_d
    count = 3 if log_d <= 17 else 2
    vs = [oracle.random_fr_mont(0x900 + 7 * log_d + i, n) for i in range(count)]
    th = oracle.cpu_threads()
    for op in (czk.NTT_FFT, czk.NTT_IFFT, czk.NTT_COSET_FFT, czk.NTT_COSET_IFFT, czk.NTT_IFFT_COSET_FFT):
        if log_d >= 20 and op in (czk.NTT_FFT, czk.NTT_IFFT):
            continue  # the large sizes keep to the ops the witness map uses (the others are covered above at every size)
        dv = [ctx.vec_from(v) for v in vs]
        ctx.ntt_batch(dv, log_d, op)
        for v, d in zip(vs, dv):
            if op == czk.NTT_IFFT_COSET_FFT:
                exp = oracle.ntt(oracle.ntt(v, True, False, threads=th), False, True, threads=th)
            else:
                exp = oracle.ntt(v, bool(op & 1), bool(op & 2), threads=th)
            assert (d.numpy() == exp).all(), (log_d, op)


def test_ntt_batch_rejects_bad_arguments(ctx, czk, oracle):
    v = ctx.vec_from(oracle.random_fr_mont(1, 16))
    with pytest.raises(czk.CzkError):
        ctx.ntt_batch([v, v], 4, czk.NTT_FFT)  # the same vector twice
    with pytest.raises(czk.CzkError):
        ctx.ntt_batch([v], 5, czk.NTT_FFT)  # shorter than the domain
    with pytest.raises(czk.CzkError):
        ctx.ntt_batch([v], 4, 7)  # no such transform
    ctx.ntt_batch([], 4, czk.NTT_FFT)  # nothing to do


@pytest.mark.parametrize("log_m", [0, 1, 2, 3, 5, 9, 12, 15])
def test_mixed_radix_ntt_matches_oracle(ctx, czk, oracle, log_m):
    """czk_ntt_mixed_vec_batch over 3 * 2^log_m points (the Plonk prover's wire domain) against the oracle's restatement of
    MixedRadixEvaluationDomain (algebra/poly/src/domain/mixed_radix.rs): all ops, several vectors per call."""
    n = 3 << log_m
    count = 3 if log_m <= 9 else 2
    vs = [oracle.random_fr_mont(0xa00 + 5 * log_m + i, n) for i in range(count)]
    for op in (czk.NTT_FFT, czk.NTT_IFFT, czk.NTT_COSET_FFT, czk.NTT_COSET_IFFT, czk.NTT_IFFT_COSET_FFT):
        dv = [ctx.vec_from(v) for v in vs]
        ctx.ntt_mixed_batch(dv, log_m, op)
        for v, d in zip(vs, dv):
            if op == czk.NTT_IFFT_COSET_FFT:
                exp = oracle.ntt_mixed(oracle.ntt_mixed(v, True, False), False, True)
            else:
                exp = oracle.ntt_mixed(v, bool(op & 1), bool(op & 2))
            assert (d.numpy() == exp).all(), (log_m, op)


def test_mixed_radix_ntt_2_20_roundtrip_and_evaluation(ctx, czk, oracle, pymodel):
    """3 * 2^20 points: iFFT(FFT(x)) = x, coset round trip, and X[j] = p(w^j) at a few j by Horner (size-independent checks)."""
    log_m = 20
    n = 3 << log_m
    x = oracle.random_fr_mont(0xbeef, n)
    d = ctx.vec_from(x)
    ctx.ntt_mixed_batch([d], log_m, czk.NTT_FFT)
    y = d.numpy()
    w = czk.mixed_domain_params(log_m)["group_gen"]
    for j in (0, 1, 3, (1 << log_m) + 5, n - 1):
        assert (oracle.poly_eval(x, oracle.fr_pow_u64(w, j)) == y[j]).all(), j
    ctx.ntt_mixed_batch([d], log_m, czk.NTT_IFFT)
    assert (d.numpy() == x).all()
    ctx.ntt_mixed_batch([d], log_m, czk.NTT_COSET_FFT)
    ctx.ntt_mixed_batch([d], log_m, czk.NTT_COSET_IFFT)
    assert (d.numpy() == x).all()


def test_mixed_domain_params_match_oracle(czk, oracle):
    for log_m in (0, 1, 7, 20):
        got = czk.mixed_domain_params(log_m)
        exp = oracle.mixed_domain_params(3 << log_m)
        for k, e in zip(("group_gen", "group_gen_inv", "size_inv", "generator_inv"), exp):
            assert (got[k] == e).all(), (log_m, k)


def test_mixed_radix_ntt_rejects_bad_arguments(ctx, czk, oracle):
    v = ctx.vec_from(oracle.random_fr_mont(1, 24))
    with pytest.raises(czk.CzkError):
        ctx.ntt_mixed_batch([v, v], 3, czk.NTT_FFT)
    with pytest.raises(czk.CzkError):
        ctx.ntt_mixed_batch([v], 4, czk.NTT_FFT)  # 48 points needed
    with pytest.raises(czk.CzkError):
        ctx.ntt_mixed_batch([v], 3, 9)
    ctx.ntt_mixed_batch([], 3, czk.NTT_FFT)


def test_mixed_radix_ntt_host_entry(ctx, oracle):
    x = oracle.random_fr_mont(0xc0de, 3 << 7)
    for inverse in (False, True):
        for coset in (False, True):
            assert (ctx.ntt_mixed(x, inverse, coset) == oracle.ntt_mixed(x, inverse, coset)).all(), (inverse, coset)

"""The oracle's MixedRadixEvaluationDomain restatement (oracle/czk_oracle_mixed.inc) pinned against the transform's
definition and against the radix-2 restatement: no GPU."""
import numpy as np
import pytest


def _rnd(rng, n):
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    return a


@pytest.mark.parametrize("n", [1, 2, 8, 64])
def test_power_of_two_sizes_agree_with_the_radix2_domain(oracle, n):
    x = _rnd(np.random.default_rng(n), n)
    for inverse in (False, True):
        for coset in (False, True):
            assert (oracle.ntt_mixed(x, inverse, coset) == oracle.ntt(x, inverse, coset)).all(), (inverse, coset)


@pytest.mark.parametrize("n", [3, 6, 12, 96])
def test_mixed_sizes_are_the_dft_under_get_root_of_unity(oracle, pymodel, n):
    """out[j] = sum_i in[i] w^(ij) with w = get_root_of_unity(n) (an element of order exactly n), checked with Python integers
    (the reference's own test evaluates the polynomial at every domain element: mixed_radix.rs tests)."""
    R = pymodel.R_MOD
    x = _rnd(np.random.default_rng(100 + n), n)
    xi = oracle.fr_to_ints(x)
    gen, gen_inv, size_inv, generator_inv = [oracle.fr_to_ints([v])[0] for v in oracle.mixed_domain_params(n)]
    assert pow(gen, n, R) == 1 and all(pow(gen, n // q, R) != 1 for q in (2, 3) if n % q == 0)
    assert gen * gen_inv % R == 1 and size_inv * n % R == 1 and generator_inv * 22 % R == 1
    want = [sum(xi[i] * pow(gen, i * j, R) for i in range(n)) % R for j in range(n)]
    assert oracle.fr_to_ints(oracle.ntt_mixed(x)) == want
    want_coset = [sum(xi[i] * pow(22 * pow(gen, j, R), i, R) for i in range(n)) % R for j in range(n)]
    assert oracle.fr_to_ints(oracle.ntt_mixed(x, coset=True)) == want_coset
    assert (oracle.ntt_mixed(oracle.ntt_mixed(x), inverse=True) == x).all()
    assert (oracle.ntt_mixed(oracle.ntt_mixed(x, coset=True), inverse=True, coset=True) == x).all()


def test_mixed_generator_cubed_is_the_radix2_generator(oracle, pymodel):
    """What the device decomposition (three radix-2 transforms + one combining pass) relies on."""
    R = pymodel.R_MOD
    for k in range(0, 12):
        w = oracle.fr_to_ints([oracle.mixed_domain_params(3 << k)[0]])[0]
        w2 = oracle.fr_to_ints([oracle.domain_params(1 << k)["group_gen"]])[0]
        assert pow(w, 3, R) == w2


def test_sizes_that_are_not_domains_are_refused(oracle):
    for n in (5, 9, 18, 7 * 8):
        with pytest.raises(AssertionError):
            oracle.ntt_mixed(np.zeros((n, 4), np.uint64))


@pytest.mark.parametrize("k", [0, 1, 4])
@pytest.mark.parametrize("inverse", [False, True])
def test_device_decomposition_matches_the_reference_algorithm(oracle, pymodel, k, inverse):
    """The device path (csrc/ntt.cu, k_mr_split / k_mr_combine) is de-interleave -> three radix-2 transforms -> one combining
    pass with zeta = w^M; the same steps in Python integers over the oracle's radix-2 transform give the mixed-radix result."""
    R = pymodel.R_MOD
    M, N = 1 << k, 3 << k
    x = _rnd(np.random.default_rng(7 * k + inverse), N)
    gen, gen_inv, _, _ = [oracle.fr_to_ints([v])[0] for v in oracle.mixed_domain_params(N)]
    w = gen_inv if inverse else gen
    zeta = pow(w, M, R)
    c = pow(3, R - 2, R) if inverse else 1  # the radix-2 inverse already scaled by M^-1
    ys = [oracle.fr_to_ints(oracle.ntt(np.ascontiguousarray(x[r::3]), inverse=inverse)) for r in range(3)]
    out = [0] * N
    for j in range(M):
        t1, t2 = ys[1][j] * pow(w, j, R) % R, ys[2][j] * pow(w, 2 * j, R) % R
        for s in range(3):
            out[j + s * M] = c * (ys[0][j] + pow(zeta, s, R) * t1 + pow(zeta, 2 * s, R) * t2) % R
    assert out == oracle.fr_to_ints(oracle.ntt_mixed(x, inverse=inverse))

"""CPU tier: the product's ark-serialize wire format (czk_fr/g1/g2/proof _serialize / _deserialize, host code in
libczk_b200.so) against the oracle's Python big-int restatement (oracle/pymodel.py) of algebra/serialize,
ff/src/fields/macros.rs:1-87, quadratic_extension.rs:600-647 and short_weierstrass_jacobian.rs:792-895.
The reference holds no byte vectors for BLS12-377 points, so the pins are: the independent model, round
trips, the documented sizes (Fr 32, G1 48 / 96, G2 96 / 192, proof 192 bytes) and the rejections the reference makes."""
import random

import numpy as np
import pytest


def _pts(oracle, pymodel, g, k, seed):
    G = oracle.G1 if g == 1 else oracle.G2
    gen = pymodel.G1_GEN if g == 1 else pymodel.G2_GEN
    mul = pymodel.g1_mul if g == 1 else pymodel.g2_mul
    rnd = random.Random(seed)
    ints = [mul(gen, rnd.randrange(1, pymodel.R_MOD)) for _ in range(k)]
    xy, _ = G.affine_from_ints(ints)
    return ints, xy


def test_fr_bytes(czk, oracle, pymodel):
    rnd = random.Random(1)
    vals = [0, 1, pymodel.R_MOD - 1, 1 << 200] + [rnd.randrange(pymodel.R_MOD) for _ in range(50)]
    data = czk.fr_serialize(oracle.fr_from_ints(vals))
    assert len(data) == 32 * len(vals)
    assert data == b"".join(pymodel.ser_fr(v) for v in vals)
    assert oracle.fr_to_ints(czk.fr_deserialize(data)) == vals
    with pytest.raises(czk.CzkError):  # the modulus itself is not canonical
        czk.fr_deserialize(pymodel.R_MOD.to_bytes(32, "little"))
    with pytest.raises(czk.CzkError):
        czk.fr_deserialize(b"\xff" * 32)


@pytest.mark.parametrize("g", [1, 2])
@pytest.mark.parametrize("compressed", [True, False])
def test_point_bytes_match_model_and_round_trip(czk, oracle, pymodel, g, compressed):
    ints, xy = _pts(oracle, pymodel, g, 12, seed=10 * g + compressed)
    neg = pymodel.g1_neg if g == 1 else pymodel.g2_neg
    G = oracle.G1 if g == 1 else oracle.G2
    ints = ints + [neg(p) for p in ints[:4]]  # both signs of y
    xy, _ = G.affine_from_ints(ints)
    inf = np.zeros(len(ints), np.uint8)
    data = czk.point_serialize(g, xy, inf, compressed)
    size = 48 * g * (1 if compressed else 2)
    assert len(data) == size * len(ints)
    for i, P in enumerate(ints):
        assert data[size * i:size * (i + 1)] == pymodel.ser_point(g, P, compressed), i
        assert pymodel.deser_point(g, data[size * i:size * (i + 1)], compressed) == ("ok", P)
    back, binf = czk.point_deserialize(g, data, compressed)
    assert not binf.any() and (back == xy).all()
    # infinity
    zinf = czk.point_serialize(g, xy[:1], np.ones(1, np.uint8), compressed)
    assert zinf == pymodel.ser_point(g, None, compressed)
    _, binf = czk.point_deserialize(g, zinf, compressed)
    assert binf[0] == 1


@pytest.mark.parametrize("g", [1, 2])
def test_deserialize_rejects_what_the_reference_rejects(czk, oracle, pymodel, g):
    ints, xy = _pts(oracle, pymodel, g, 1, seed=77 + g)
    good = bytearray(czk.point_serialize(g, xy, None, True))
    both = bytearray(good)
    both[-1] |= 0xC0  # positive-y and infinity together: SWFlags::from_u8 -> None
    with pytest.raises(czk.CzkError):
        czk.point_deserialize(g, bytes(both))
    assert pymodel.deser_point(g, bytes(both))[0] == "err"
    # an x with no point on the curve (or outside the subgroup): walk x upwards until the model rejects it
    F = pymodel._F(g)
    x = ints[0][0]
    for _ in range(64):
        x = (x + 1) % pymodel.Q_MOD if g == 1 else ((x[0] + 1) % pymodel.Q_MOD, x[1])
        enc = pymodel._coord_bytes(g, x)
        st, why = pymodel.deser_point(g, bytes(enc))
        if st == "err":
            with pytest.raises(czk.CzkError):
                czk.point_deserialize(g, bytes(enc))
            break
        # on the curve and (by luck) in the subgroup: both must agree on the point
        got, _ = czk.point_deserialize(g, bytes(enc))
        G = oracle.G1 if g == 1 else oracle.G2
        assert G.affine_to_ints(got)[0] == why
    else:
        pytest.fail("no rejected x found")
    # non-canonical coordinate (>= q)
    big = bytearray((pymodel.Q_MOD).to_bytes(48, "little")) if g == 1 else bytearray((0).to_bytes(48, "little") + pymodel.Q_MOD.to_bytes(48, "little"))
    with pytest.raises(czk.CzkError):
        czk.point_deserialize(g, bytes(big))
    # a curve point outside the prime-order subgroup is accepted only when the check is off (deserialize_unchecked)
    if g == 1:
        x = 1
        while True:
            y = pymodel.fq_sqrt((x * x * x + 1) % pymodel.Q_MOD)
            if y is not None and pymodel._mul_raw(F, (x, y), pymodel.R_MOD) is not None:
                break
            x += 1
        enc = pymodel.ser_point(1, (x, y), False)
        with pytest.raises(czk.CzkError):
            czk.point_deserialize(1, enc, compressed=False, check_subgroup=True)
        got, _ = czk.point_deserialize(1, enc, compressed=False, check_subgroup=False)
        assert oracle.G1.affine_to_ints(got)[0] == (x, y)


def test_proof_bytes(czk, oracle, pymodel):
    (a, c), xy1 = _pts(oracle, pymodel, 1, 2, seed=5)
    (b,), xy2 = _pts(oracle, pymodel, 2, 1, seed=6)
    proof = np.concatenate([xy1[0], xy2[0], xy1[1]])
    data = czk.groth16_proof_serialize(proof, np.zeros(3, np.uint8))
    assert len(data) == 192
    assert data == pymodel.ser_point(1, a) + pymodel.ser_point(2, b) + pymodel.ser_point(1, c)
    back, inf = czk.groth16_proof_deserialize(data)
    assert (back == proof).all() and not inf.any()

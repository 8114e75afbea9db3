"""Golden parameter literals of the reference (tests/golden/bls12_377_constants.json, extracted from
curves/bls12_377/src/** by tests/golden/make_constants.py) pin:
  * the oracle's self-derived Montgomery constants (R, R2, INV),
  * the Python model's constants,
  * the product's generated parameter header (tools/gen_params.py),
and the root-of-unity derivation (SURVEY.md "omega trap")."""
import json
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
GOLD = json.loads((ROOT / "tests/golden/bls12_377_constants.json").read_text())


def g(section, name):
    return int(GOLD[section][name])


def test_moduli_and_montgomery_constants(oracle, pymodel):
    p = oracle.params()
    assert g("fr", "MODULUS") == pymodel.R_MOD and g("fq", "MODULUS") == pymodel.Q_MOD
    assert p["fr_R"] == g("fr", "R") == pow(2, 256, pymodel.R_MOD)
    assert p["fr_R2"] == g("fr", "R2")
    assert p["fr_INV"] == g("fr", "INV")
    assert p["fq_R"] == g("fq", "R") == pow(2, 384, pymodel.Q_MOD)
    assert p["fq_R2"] == g("fq", "R2")
    assert p["fq_INV"] == g("fq", "INV")
    assert g("fr", "MODULUS_BITS") == 253 and g("fq", "MODULUS_BITS") == 377
    assert g("fr", "TWO_ADICITY") == 47


def test_generator_and_roots_of_unity(oracle, pymodel):
    # GENERATOR literal decodes to 22 (the source comment says 11)
    assert pymodel.fr_from_mont(g("fr", "GENERATOR")) == 22 == pymodel.FR_GENERATOR
    assert oracle.limbs_to_ints(oracle.fr_from_ints([22]))[0] == g("fr", "GENERATOR")
    # TWO_ADIC_ROOT_OF_UNITY = 22^t, LARGE_SUBGROUP_ROOT_OF_UNITY = 11^((r-1)/(3*2^47))
    assert pymodel.fr_from_mont(g("fr", "TWO_ADIC_ROOT_OF_UNITY")) == pymodel.FR_TWO_ADIC_ROOT
    assert pymodel.fr_from_mont(g("fr", "LARGE_SUBGROUP_ROOT_OF_UNITY")) == pymodel.FR_LARGE_SUBGROUP_ROOT
    t = g("fr", "T")
    assert (pymodel.R_MOD - 1) == t << 47 and t % 2 == 1
    assert g("fr", "MODULUS_MINUS_ONE_DIV_TWO") == (pymodel.R_MOD - 1) // 2
    # the domain generator is 11^((r-1)/D), which differs from TWO_ADIC_ROOT^(2^(47-k)) for k >= 4
    for k in range(0, 25):
        d = oracle.domain_params(1 << k)
        w = oracle.fr_to_ints(d["group_gen"][None, :])[0]
        assert w == pow(11, (pymodel.R_MOD - 1) >> k, pymodel.R_MOD) == pymodel.fr_root_of_unity(1 << k)
        assert pow(w, 1 << k, pymodel.R_MOD) == 1 and (k == 0 or pow(w, 1 << (k - 1), pymodel.R_MOD) != 1)
        two_adic = pow(pymodel.FR_TWO_ADIC_ROOT, 1 << (47 - k), pymodel.R_MOD)
        assert (w == two_adic) == (k <= 3)
        assert oracle.fr_to_ints(d["size_inv"][None, :])[0] == pow(1 << k, -1, pymodel.R_MOD)
        assert oracle.fr_to_ints(d["generator_inv"][None, :])[0] == pow(22, -1, pymodel.R_MOD)


def test_curve_constants(pymodel):
    assert (g("g1", "G1_GENERATOR_X"), g("g1", "G1_GENERATOR_Y")) == pymodel.G1_GEN
    assert ((g("g2", "G2_GENERATOR_X_C0"), g("g2", "G2_GENERATOR_X_C1")),
            (g("g2", "G2_GENERATOR_Y_C0"), g("g2", "G2_GENERATOR_Y_C1"))) == pymodel.G2_GEN
    assert g("g2", "COEFF_B_C1") == pymodel.G2_B[1]
    assert g("fq2", "NONRESIDUE") == -5
    assert pymodel.g1_on_curve(pymodel.G1_GEN) and pymodel.g2_on_curve(pymodel.G2_GEN)
    # prime-order subgroup (curves/bls12_377/src/curves/tests.rs:20-62)
    assert pymodel.g1_mul(pymodel.G1_GEN, pymodel.R_MOD - 1) == pymodel.g1_neg(pymodel.G1_GEN)
    assert pymodel.g2_mul(pymodel.G2_GEN, pymodel.R_MOD - 1) == pymodel.g2_neg(pymodel.G2_GEN)
    # cofactor * r == #E(Fq) is not checked here; cofactor_inv * cofactor == 1 mod r is
    assert g("g1", "COFACTOR") * g("g1", "COFACTOR_INV") % pymodel.R_MOD == 1
    assert g("g2", "COFACTOR") * g("g2", "COFACTOR_INV") % pymodel.R_MOD == 1


def test_product_parameter_header_matches_golden(pymodel):
    """collaborative-zksnark_b200/csrc/bls12_377_params.cuh is generated; its 64-bit tables must equal
    the reference literals."""
    src = (ROOT / "collaborative-zksnark_b200/csrc/bls12_377_params.cuh").read_text()

    def table(struct, name):
        body = src[src.index(f"struct {struct}"):]
        m = re.search(name + r"\[\d+\] = \{([^}]*)\}", body)
        limbs = [int(x.replace("ull", ""), 16) for x in m.group(1).split(",")]
        return sum(l << (64 * i) for i, l in enumerate(limbs))

    assert table("FrParams", "MOD64") == g("fr", "MODULUS")
    assert table("FrParams", "ONE64") == g("fr", "R")
    assert table("FrParams", "R2_64") == g("fr", "R2")
    assert table("FrParams", "GENERATOR_64") == g("fr", "GENERATOR")
    assert table("FqParams", "MOD64") == g("fq", "MODULUS")
    assert table("FqParams", "ONE64") == g("fq", "R")
    assert table("FqParams", "R2_64") == g("fq", "R2")
    inv = int(re.search(r"struct FrParams.*?INV64 = (0x[0-9a-f]+)ull", src, re.S).group(1), 16)
    assert inv == g("fr", "INV")
    inv = int(re.search(r"struct FqParams.*?INV64 = (0x[0-9a-f]+)ull", src, re.S).group(1), 16)
    assert inv == g("fq", "INV")
    root47 = table("FrParams", "ROOT_2_47_64")
    assert pymodel.fr_from_mont(root47) == pow(pymodel.fr_from_mont(g("fr", "LARGE_SUBGROUP_ROOT_OF_UNITY")), 3, pymodel.R_MOD)
    g1 = table("CurveConsts", "G1_GEN")
    assert pymodel.fq_from_mont(g1 & ((1 << 384) - 1)) == g("g1", "G1_GENERATOR_X")
    assert pymodel.fq_from_mont(g1 >> 384) == g("g1", "G1_GENERATOR_Y")

"""ctypes binding of oracle/libczk_oracle.so.  TEST INFRASTRUCTURE ONLY.

Allowed importers: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline and
`--impl reference` legs.  The product package never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

u64p = C.POINTER(C.c_uint64)
u8p = C.POINTER(C.c_uint8)


def build(force: bool = False) -> Path:
    so = _HERE / "libczk_oracle.so"
    srcs = [_HERE / n for n in ("czk_oracle.c", "czk_oracle_groth16.inc", "czk_oracle_plonk.inc", "czk_oracle_mixed.inc", "fp_tmpl.h", "ec_tmpl.h", "Makefile")]
    if force or not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(_HERE), "-s"], check=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = _HERE / "libczk_oracle.so"
        if not so.exists():
            build()
        _LIB = C.CDLL(str(so))
        _LIB.orc_init()
    return _LIB


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


def _p8(a):
    if a is None:
        return None
    assert a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u8p)


def cpu_threads() -> int:
    return len(os.sched_getaffinity(0))


# ---------------------------------------------------------------- int <-> limb arrays
def ints_to_limbs(vals, nl: int) -> np.ndarray:
    out = np.zeros((len(vals), nl), dtype=np.uint64)
    for i, v in enumerate(vals):
        for j in range(nl):
            out[i, j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def limbs_to_ints(arr: np.ndarray):
    arr = np.asarray(arr, dtype=np.uint64)
    if arr.ndim == 1:
        arr = arr[None, :]
    return [sum(int(x) << (64 * j) for j, x in enumerate(row)) for row in arr]


# ---------------------------------------------------------------- fields
def _binop(name, nl):
    def f(a: np.ndarray, b: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, nl)
        b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, nl)
        r = np.empty_like(a)
        getattr(lib(), name)(_p(r), _p(a), _p(b), C.c_size_t(a.shape[0]))
        return r

    return f


def _unop(name, nl):
    def f(a: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, nl)
        r = np.empty_like(a)
        getattr(lib(), name)(_p(r), _p(a), C.c_size_t(a.shape[0]))
        return r

    return f


fr_mul = _binop("orc_fr_mul", 4)
fr_add = _binop("orc_fr_add", 4)
fr_sub = _binop("orc_fr_sub", 4)
fr_neg = _unop("orc_fr_neg", 4)
fr_inv = _unop("orc_fr_inv", 4)
fr_from_repr = _unop("orc_fr_from_repr", 4)
fr_into_repr = _unop("orc_fr_into_repr", 4)
fq_mul = _binop("orc_fq_mul", 6)
fq_add = _binop("orc_fq_add", 6)
fq_sub = _binop("orc_fq_sub", 6)
fq_inv = _unop("orc_fq_inv", 6)
fq_from_repr = _unop("orc_fq_from_repr", 6)
fq_into_repr = _unop("orc_fq_into_repr", 6)
fq2_mul = _binop("orc_fq2_mul", 12)
fq2_sqr = _unop("orc_fq2_sqr", 12)
fq2_inv = _unop("orc_fq2_inv", 12)


def fr_from_ints(vals) -> np.ndarray:
    """canonical ints -> Montgomery limb array (n,4)."""
    return fr_from_repr(ints_to_limbs([v for v in vals], 4))


def fr_to_ints(arr) -> list:
    return limbs_to_ints(fr_into_repr(arr))


def fq_from_ints(vals) -> np.ndarray:
    return fq_from_repr(ints_to_limbs(vals, 6))


def fq_to_ints(arr) -> list:
    return limbs_to_ints(fq_into_repr(arr))


def params():
    frR = np.zeros(4, np.uint64)
    frR2 = np.zeros(4, np.uint64)
    fqR = np.zeros(6, np.uint64)
    fqR2 = np.zeros(6, np.uint64)
    fri = C.c_uint64()
    fqi = C.c_uint64()
    lib().orc_params(_p(frR), _p(frR2), C.byref(fri), _p(fqR), _p(fqR2), C.byref(fqi))
    return dict(fr_R=limbs_to_ints(frR)[0], fr_R2=limbs_to_ints(frR2)[0], fr_INV=fri.value,
                fq_R=limbs_to_ints(fqR)[0], fq_R2=limbs_to_ints(fqR2)[0], fq_INV=fqi.value)


# ---------------------------------------------------------------- groups
class Group:
    """g = 'g1' (12 words per affine point) or 'g2' (24 words)."""

    def __init__(self, g: str):
        self.g = g
        self.w = 12 if g == "g1" else 24

    def affine_from_ints(self, pts) -> tuple:
        """pts: list of None | (x, y) with ints (G1) or ((x0,x1),(y0,y1)) (G2). -> (xy[n,w], inf[n])"""
        n = len(pts)
        flat = []
        inf = np.zeros(n, np.uint8)
        for i, p in enumerate(pts):
            if p is None:
                inf[i] = 1
                flat += [0, 1] if self.g == "g1" else [0, 0, 1, 0]
            elif self.g == "g1":
                flat += [p[0], p[1]]
            else:
                flat += [p[0][0], p[0][1], p[1][0], p[1][1]]
        xy = fq_from_ints(flat).reshape(n, self.w)
        return xy, inf

    def affine_to_ints(self, xy, inf=None):
        xy = np.ascontiguousarray(xy, dtype=np.uint64).reshape(-1, self.w)
        vals = fq_to_ints(xy.reshape(-1, 6))
        per = self.w // 6
        out = []
        for i in range(xy.shape[0]):
            if inf is not None and inf[i]:
                out.append(None)
            elif self.g == "g1":
                out.append((vals[2 * i], vals[2 * i + 1]))
            else:
                v = vals[per * i: per * i + per]
                out.append(((v[0], v[1]), (v[2], v[3])))
        return out

    def scalar_mul(self, base_xy, scalar_mont, base_inf=0):
        out = np.zeros(self.w, np.uint64)
        inf = getattr(lib(), f"orc_{self.g}_scalar_mul")(_p(out), _p(np.ascontiguousarray(base_xy, np.uint64)),
                                                         C.c_int(int(base_inf)), _p(np.ascontiguousarray(scalar_mont, np.uint64)))
        return out, inf

    def msm(self, bases_xy, inf, scalars, montgomery=True, threads=1):
        bases_xy = np.ascontiguousarray(bases_xy, np.uint64).reshape(-1, self.w)
        scalars = np.ascontiguousarray(scalars, np.uint64).reshape(-1, 4)
        n = min(bases_xy.shape[0], scalars.shape[0])
        out = np.zeros(self.w, np.uint64)
        r = getattr(lib(), f"orc_{self.g}_msm")(_p(out), _p(bases_xy), _p8(inf), _p(scalars), C.c_int(int(montgomery)),
                                                C.c_size_t(n), C.c_int(threads))
        return out, r

    def msm_naive(self, bases_xy, inf, scalars_mont):
        bases_xy = np.ascontiguousarray(bases_xy, np.uint64).reshape(-1, self.w)
        scalars_mont = np.ascontiguousarray(scalars_mont, np.uint64).reshape(-1, 4)
        n = min(bases_xy.shape[0], scalars_mont.shape[0])
        out = np.zeros(self.w, np.uint64)
        r = getattr(lib(), f"orc_{self.g}_msm_naive")(_p(out), _p(bases_xy), _p8(inf), _p(scalars_mont), C.c_size_t(n))
        return out, r

    def gen_progression(self, base_xy, k0_mont, kstep_mont, n, threads=1):
        out = np.zeros((n, self.w), np.uint64)
        getattr(lib(), f"orc_{self.g}_gen_progression")(_p(out), _p(np.ascontiguousarray(base_xy, np.uint64)),
                                                        _p(np.ascontiguousarray(k0_mont, np.uint64)),
                                                        _p(np.ascontiguousarray(kstep_mont, np.uint64)), C.c_size_t(n), C.c_int(threads))
        return out


G1 = Group("g1")
G2 = Group("g2")


# ---------------------------------------------------------------- NTT
def ntt(data: np.ndarray, inverse=False, coset=False, threads=1) -> np.ndarray:
    """data: (2^k, 4) Montgomery Fr.  Returns a transformed copy."""
    a = np.array(data, dtype=np.uint64, order="C").reshape(-1, 4)
    n = a.shape[0]
    log_d = n.bit_length() - 1
    assert 1 << log_d == n
    ok = lib().orc_ntt(_p(a), C.c_uint(log_d), C.c_int(int(inverse)), C.c_int(int(coset)), C.c_int(threads))
    assert ok
    return a


def ntt_mixed(data: np.ndarray, inverse=False, coset=False) -> np.ndarray:
    """MixedRadixEvaluationDomain transforms (algebra/poly/src/domain/mixed_radix.rs): data is (2^k, 4) or (3 * 2^k, 4)
    Montgomery Fr, natural order in and out.  Returns a transformed copy."""
    a = np.array(data, dtype=np.uint64, order="C").reshape(-1, 4)
    ok = lib().orc_ntt_mixed(_p(a), C.c_size_t(a.shape[0]), C.c_int(int(inverse)), C.c_int(int(coset)))
    assert ok, "not a mixed-radix domain size"
    return a


def mixed_domain_params(size: int):
    """(group_gen, group_gen_inv, size_inv, generator_inv) of MixedRadixEvaluationDomain::new(size), Montgomery limbs."""
    out = [np.zeros(4, np.uint64) for _ in range(4)]
    assert lib().orc_mixed_domain_params(C.c_size_t(size), *[_p(o) for o in out])
    return tuple(out)


def serial_radix2_fft(data: np.ndarray, inverse=False) -> np.ndarray:
    a = np.array(data, dtype=np.uint64, order="C").reshape(-1, 4)
    n = a.shape[0]
    log_n = n.bit_length() - 1
    assert lib().orc_serial_radix2_fft(_p(a), C.c_uint(log_n), C.c_int(int(inverse)))
    return a


def poly_eval(coeffs: np.ndarray, x_mont: np.ndarray) -> np.ndarray:
    coeffs = np.ascontiguousarray(coeffs, np.uint64).reshape(-1, 4)
    out = np.zeros(4, np.uint64)
    lib().orc_poly_eval(_p(out), _p(coeffs), C.c_size_t(coeffs.shape[0]), _p(np.ascontiguousarray(x_mont, np.uint64)))
    return out


def fr_pow_u64(a_mont, e: int) -> np.ndarray:
    out = np.zeros(4, np.uint64)
    lib().orc_fr_pow_u64(_p(out), _p(np.ascontiguousarray(a_mont, np.uint64)), C.c_uint64(e))
    return out


def domain_params(num_coeffs: int):
    size = C.c_uint64()
    gg = np.zeros(4, np.uint64)
    ggi = np.zeros(4, np.uint64)
    si = np.zeros(4, np.uint64)
    gi = np.zeros(4, np.uint64)
    ok = lib().orc_domain_params(C.c_size_t(num_coeffs), C.byref(size), _p(gg), _p(ggi), _p(si), _p(gi))
    assert ok
    return dict(size=size.value, group_gen=gg, group_gen_inv=ggi, size_inv=si, generator_inv=gi)


def divide_by_vanishing_on_coset(data: np.ndarray, threads=1) -> np.ndarray:
    a = np.array(data, dtype=np.uint64, order="C").reshape(-1, 4)
    log_d = a.shape[0].bit_length() - 1
    assert lib().orc_divide_by_vanishing_on_coset(_p(a), C.c_uint(log_d), C.c_int(threads))
    return a


# ---------------------------------------------------------------- seeded inputs (SplitMix64 rejection sampling)
def splitmix64(seed: int, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.uint64)
    x = seed & 0xFFFFFFFFFFFFFFFF
    M = 0xFFFFFFFFFFFFFFFF
    for i in range(n):
        x = (x + 0x9E3779B97F4A7C15) & M
        z = x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        out[i] = z ^ (z >> 31)
    return out


def random_fr_canonical(seed: int, n: int) -> np.ndarray:
    """n uniform values in [0, r) as canonical limb array (n,4): vectorised rejection sampling."""
    from . import pymodel

    rng = np.random.Generator(np.random.PCG64(seed))
    mod_limbs = ints_to_limbs([pymodel.R_MOD], 4)[0]
    out = np.zeros((n, 4), dtype=np.uint64)
    need = np.arange(n)
    while need.size:
        cand = rng.integers(0, 1 << 64, size=(need.size, 4), dtype=np.uint64)
        cand[:, 3] &= np.uint64((1 << 61) - 1)  # 253-bit candidates
        lt = np.zeros(need.size, dtype=bool)
        decided = np.zeros(need.size, dtype=bool)
        for j in (3, 2, 1, 0):
            less = (cand[:, j] < mod_limbs[j]) & ~decided
            greater = (cand[:, j] > mod_limbs[j]) & ~decided
            lt |= less
            decided |= less | greater
        out[need[lt]] = cand[lt]
        need = need[~lt]
    return out


def random_fr_mont(seed: int, n: int) -> np.ndarray:
    return fr_from_repr(random_fr_canonical(seed, n))


# ---------------------------------------------------------------- Groth16 (squaring circuit) + shares
SCHEME_PLAIN, SCHEME_ADDITIVE, SCHEME_SPDZ = 0, 1, 2


def generators():
    g1 = np.zeros(12, np.uint64)
    g2 = np.zeros(24, np.uint64)
    lib().orc_generators(_p(g1), _p(g2))
    return g1, g2


def groth16_domain_size(n_sq: int) -> int:
    f = lib().orc_groth16_domain_size
    f.restype = C.c_size_t
    return f(C.c_size_t(n_sq))


def squaring_chain(start_mont, n_sq: int) -> np.ndarray:
    out = np.zeros((n_sq + 1, 4), np.uint64)
    lib().orc_squaring_chain(_p(out), _p(np.ascontiguousarray(start_mont, np.uint64)), C.c_size_t(n_sq))
    return out


def groth16_setup(n_sq: int, toxic_mont: np.ndarray, threads=1) -> dict:
    """toxic_mont: (7,4) Montgomery Fr = alpha, beta, gamma, delta, tau, g1_scalar, g2_scalar."""
    D = groth16_domain_size(n_sq)
    nv = n_sq + 2
    pk = dict(n_sq=n_sq, D=D,
              a_query=np.zeros((nv, 12), np.uint64), a_inf=np.zeros(nv, np.uint8),
              b_g1_query=np.zeros((nv, 12), np.uint64), b1_inf=np.zeros(nv, np.uint8),
              b_g2_query=np.zeros((nv, 24), np.uint64), b2_inf=np.zeros(nv, np.uint8),
              h_query=np.zeros((D - 1, 12), np.uint64), h_inf=np.zeros(D - 1, np.uint8),
              l_query=np.zeros((n_sq, 12), np.uint64), l_inf=np.zeros(n_sq, np.uint8),
              vk_g1=np.zeros((3, 12), np.uint64), vk_g2=np.zeros((3, 24), np.uint64),
              gamma_abc_g1=np.zeros((2, 12), np.uint64))
    toxic_mont = np.ascontiguousarray(toxic_mont, np.uint64).reshape(7, 4)
    ok = lib().orc_groth16_setup(C.c_size_t(n_sq), _p(toxic_mont), _p(pk["a_query"]), _p8(pk["a_inf"]), _p(pk["b_g1_query"]),
                                 _p8(pk["b1_inf"]), _p(pk["b_g2_query"]), _p8(pk["b2_inf"]), _p(pk["h_query"]), _p8(pk["h_inf"]),
                                 _p(pk["l_query"]), _p8(pk["l_inf"]), _p(pk["vk_g1"]), _p(pk["vk_g2"]), _p(pk["gamma_abc_g1"]),
                                 C.c_int(threads))
    assert ok
    return pk


def king_share_batch(values_mont: np.ndarray, n_parties: int, seed: int):
    """Additive shares in the shape of king_share_batch (add.rs:105-117 / spdz.rs:150-162): parties
    0..n-2 get uniform values, party n-1 gets f - sum.  (The reference's ChaCha stream is not reproduced.)"""
    values_mont = np.ascontiguousarray(values_mont, np.uint64).reshape(-1, 4)
    k = values_mont.shape[0]
    shares = [random_fr_mont(seed + 1000 * p, k) for p in range(n_parties - 1)]
    last = values_mont.copy()
    for s in shares:
        last = fr_sub(last, s)
    return shares + [last]


def _chain_ptrs(chain_shares):
    arrs = [np.ascontiguousarray(c, np.uint64).reshape(-1, 4) for c in chain_shares]
    ptrs = (u64p * len(arrs))(*[_p(a) for a in arrs])
    return arrs, ptrs


def groth16_witness_map(scheme, n_sq, chain_shares, threads=1):
    n = len(chain_shares)
    D = groth16_domain_size(n_sq)
    arrs, ptrs = _chain_ptrs(chain_shares)
    h = np.zeros((n, D, 4), np.uint64)
    ok = lib().orc_groth16_witness_map(C.c_int(scheme), C.c_int(n), C.c_size_t(n_sq), ptrs, _p(h), C.c_int(threads))
    return h, bool(ok)


def groth16_prove(scheme, n_sq, chain_shares, r_sh, s_sh, pk, threads=1, want_h=True):
    n = len(chain_shares)
    D = pk["D"]
    arrs, ptrs = _chain_ptrs(chain_shares)
    r_sh = np.ascontiguousarray(r_sh, np.uint64).reshape(n, 4)
    s_sh = np.ascontiguousarray(s_sh, np.uint64).reshape(n, 4)
    h = np.zeros((n, D, 4), np.uint64) if want_h else None
    proof_sh = np.zeros((n, 48), np.uint64)
    proof_sh_inf = np.zeros((n, 3), np.uint8)
    proof = np.zeros(48, np.uint64)
    proof_inf = np.zeros(3, np.uint8)
    ok = lib().orc_groth16_prove(C.c_int(scheme), C.c_int(n), C.c_size_t(n_sq), ptrs, _p(r_sh), _p(s_sh),
                                 _p(pk["a_query"]), _p8(pk["a_inf"]), _p(pk["b_g1_query"]), _p8(pk["b1_inf"]),
                                 _p(pk["b_g2_query"]), _p8(pk["b2_inf"]), _p(pk["h_query"]), _p8(pk["h_inf"]),
                                 _p(pk["l_query"]), _p8(pk["l_inf"]), _p(pk["vk_g1"]), _p(pk["vk_g2"]),
                                 _p(h) if want_h else None, _p(proof_sh), _p8(proof_sh_inf), _p(proof), _p8(proof_inf),
                                 C.c_int(threads))
    return dict(ok=bool(ok), h=h, proof_sh=proof_sh, proof_sh_inf=proof_sh_inf, proof=proof, proof_inf=proof_inf)


SCHEME_GSZ = 3


def groth16_prove_gsz(n_parties, n_sq, chain_mont, r_val, s_val, pk, threads=1):
    """Groth16 under GSZ20 shares with the reference's stubbed preprocessing (every party holds the plaintext)."""
    D = pk["D"]
    chain_mont = np.ascontiguousarray(chain_mont, np.uint64).reshape(n_sq + 1, 4)
    h = np.zeros((D, 4), np.uint64)
    proof = np.zeros(48, np.uint64)
    proof_inf = np.zeros(3, np.uint8)
    check = np.zeros(16, np.uint64)
    check_g1 = np.zeros(24, np.uint64)
    check_g1_inf = np.zeros(2, np.uint8)
    counts = np.zeros(2, np.uint64)
    ok = lib().orc_groth16_prove_gsz(C.c_int(n_parties), C.c_size_t(n_sq), _p(chain_mont), _p(np.ascontiguousarray(r_val, np.uint64)),
                                     _p(np.ascontiguousarray(s_val, np.uint64)), _p(pk["a_query"]), _p8(pk["a_inf"]),
                                     _p(pk["b_g1_query"]), _p8(pk["b1_inf"]), _p(pk["b_g2_query"]), _p8(pk["b2_inf"]),
                                     _p(pk["h_query"]), _p8(pk["h_inf"]), _p(pk["l_query"]), _p8(pk["l_inf"]), _p(pk["vk_g1"]),
                                     _p(pk["vk_g2"]), _p(h), _p(proof), _p8(proof_inf), _p(check), _p(check_g1), _p8(check_g1_inf),
                                     _p(counts), C.c_int(threads))
    return dict(ok=bool(ok), h=h, proof=proof, proof_inf=proof_inf, field_check=check[:12].reshape(3, 4), group_check_x=check[12:16],
                group_check_yz=check_g1.reshape(2, 12), group_check_inf=check_g1_inf, king_computes=int(counts[0]), opens=int(counts[1]))


def gsz_share(n_parties, coeffs_mont):
    coeffs_mont = np.ascontiguousarray(coeffs_mont, np.uint64).reshape(-1, 4)
    out = np.zeros((n_parties, 4), np.uint64)
    assert lib().orc_gsz_share(C.c_int(n_parties), _p(coeffs_mont), C.c_int(coeffs_mont.shape[0]), _p(out))
    return out


def gsz_open(shares_mont, degree):
    shares_mont = np.ascontiguousarray(shares_mont, np.uint64).reshape(-1, 4)
    out = np.zeros(4, np.uint64)
    ok = lib().orc_gsz_open(C.c_int(shares_mont.shape[0]), _p(shares_mont), C.c_int(degree), _p(out))
    return out, ok


# ------------------------------------------------------------------ Plonk / KZG10 share-level leaves
SHARE_BATCH_INV, SHARE_BATCH_DIV, SHARE_PARTIAL_PRODUCTS, SHARE_BATCH_MUL = 0, 1, 2, 3


def share_op(op, scheme, x_sh, x_mac=None, y_sh=None, y_mac=None, threads=1):
    """x_sh: (n_parties, k, 4) Montgomery shares (x_mac likewise for SPDZ).  Returns (status, x_sh', x_mac')."""
    x_sh = np.ascontiguousarray(x_sh, np.uint64).copy()
    n, k = x_sh.shape[0], x_sh.shape[1]

    def ptrs(a):
        if a is None:
            return None, None
        a = np.ascontiguousarray(a, np.uint64).copy()
        arr = (C.c_void_p * n)(*[a[p].ctypes.data for p in range(n)])
        return a, arr

    x_sh, xp = ptrs(x_sh)
    x_mac, xmp = ptrs(x_mac)
    y_sh, yp = ptrs(y_sh)
    y_mac, ymp = ptrs(y_mac)
    f = lib().orc_share_op
    f.restype = C.c_int
    st = f(C.c_int(op), C.c_int(scheme), C.c_int(n), C.c_size_t(k), xp, xmp, yp, ymp, C.c_int(threads))
    return st, x_sh, x_mac


def plonk_prove_wiring(scheme, p_shares, w_pub, powers_xy, powers_inf=None, seed=0, threads=1):
    """Prover::prove_wiring on n simulated parties (oracle/czk_oracle_plonk.inc).  p_shares: (n_parties, D, 4) coefficient
    shares of the wire polynomial, w_pub: (D, 4) public wiring polynomial, powers_xy: >= D KZG10 powers_of_g.
    Returns the revealed proof, every party's opening-proof shares and the status (1 ok, 0 zero divisor, -1 MAC failure)."""
    p_shares = np.ascontiguousarray(p_shares, np.uint64)
    n, D = p_shares.shape[0], p_shares.shape[1]  # D = 2^k, or 3 * 2^k (the reference's mixed-radix wire domain)
    w_pub = np.ascontiguousarray(w_pub, np.uint64).reshape(D, 4)
    powers_xy = np.ascontiguousarray(powers_xy, np.uint64)
    assert powers_xy.shape[0] >= D
    ptrs = (C.c_void_p * n)(*[p_shares[q].ctypes.data for q in range(n)])
    out = dict(cmt_xy=np.zeros((4, 12), np.uint64), cmt_inf=np.zeros(4, np.uint8), open_val=np.zeros((9, 4), np.uint64),
               open_pf_xy=np.zeros((9, 12), np.uint64), open_pf_inf=np.zeros(9, np.uint8), challenges=np.zeros((4, 4), np.uint64))
    sh_xy, sh_inf = np.zeros((n, 9, 12), np.uint64), np.zeros((n, 9), np.uint8)
    f = lib().orc_plonk_prove_wiring_size
    f.restype = C.c_int
    st = f(C.c_int(scheme), C.c_int(n), C.c_size_t(D), _p(powers_xy), _p8(powers_inf) if powers_inf is not None else None, ptrs, _p(w_pub),
           C.c_uint64(seed), _p(out["cmt_xy"]), _p8(out["cmt_inf"]), _p(out["open_val"]), _p(out["open_pf_xy"]), _p8(out["open_pf_inf"]),
           _p(sh_xy), _p8(sh_inf), _p(out["challenges"]), C.c_int(threads))
    return dict(status=st, proof=out, share_pf_xy=sh_xy, share_pf_inf=sh_inf)


def poly_div_linear(p_mont, z_mont):
    p_mont = np.ascontiguousarray(p_mont, np.uint64).reshape(-1, 4)
    n = p_mont.shape[0]
    q = np.zeros((max(n - 1, 0), 4), np.uint64)
    rem = np.zeros(4, np.uint64)
    qbuf = q if n > 1 else np.zeros((1, 4), np.uint64)
    lib().orc_poly_div_linear(_p(qbuf), _p(rem), _p(p_mont), C.c_size_t(n), _p(np.ascontiguousarray(z_mont, np.uint64)))
    return q, rem


def kzg_open(powers_xy, powers_inf, p_mont, z_mont, threads=1):
    p_mont = np.ascontiguousarray(p_mont, np.uint64).reshape(-1, 4)
    w = np.zeros(12, np.uint64)
    winf = np.zeros(1, np.uint8)
    ev = np.zeros(4, np.uint64)
    lib().orc_kzg_open(_p(powers_xy), _p8(powers_inf), _p(p_mont), C.c_size_t(p_mont.shape[0]),
                       _p(np.ascontiguousarray(z_mont, np.uint64)), _p(w), _p8(winf), _p(ev), C.c_int(threads))
    return w, int(winf[0]), ev


# ------------------------------------------------------------------ Groth16 on an arbitrary R1CS (CSR matrices)
def random_r1cs(seed: int, n_inst: int, n_free: int, n_cons: int, modulus: int):
    """A random satisfiable circuit in the shape of ConstraintMatrices: variables [1, inst.., free witnesses.., one product
    witness per constraint]; constraint i: <A_i, z> * <B_i, z> = z[product_i].  Returns (r1cs dict, assignment ints)."""
    import random

    rnd = random.Random(seed)
    z = [1] + [rnd.randrange(modulus) for _ in range(n_inst - 1 + n_free)]
    rows = {"a": [], "b": [], "c": []}
    for i in range(n_cons):
        avail = len(z)
        vals = {}
        for m in ("a", "b"):
            k = rnd.randrange(1, 4)
            idx = rnd.sample(range(avail), min(k, avail))
            row = [(1 if rnd.random() < 0.5 else rnd.randrange(1, modulus), j) for j in idx]
            rows[m].append(row)
            vals[m] = sum(c * z[j] for c, j in row) % modulus
        z.append(vals["a"] * vals["b"] % modulus)
        rows["c"].append([(1, len(z) - 1)])
    n_wit = len(z) - n_inst
    cs = dict(ncons=n_cons, ninst=n_inst, nwit=n_wit)
    for m in ("a", "b", "c"):
        rp = np.zeros(n_cons + 1, np.uint64)
        col, cf = [], []
        for i, row in enumerate(rows[m]):
            for c, j in row:
                col.append(j)
                cf.append(c)
            rp[i + 1] = len(col)
        cs[m] = (rp, np.array(col, np.uint32), fr_from_ints(cf))
    return cs, z


def _r1cs_ptrs(cs):
    keep = []
    rp = (C.c_void_p * 3)(*[cs[m][0].ctypes.data for m in ("a", "b", "c")])
    col = (C.c_void_p * 3)(*[cs[m][1].ctypes.data for m in ("a", "b", "c")])
    cf = (C.c_void_p * 3)(*[cs[m][2].ctypes.data for m in ("a", "b", "c")])
    return rp, col, cf, keep


def groth16_setup_r1cs(cs: dict, toxic_mont: np.ndarray, threads=1) -> dict:
    nv = cs["ninst"] + cs["nwit"]
    D = 1
    while D < cs["ncons"] + cs["ninst"]:
        D <<= 1
    pk = dict(D=D, ncons=cs["ncons"], ninst=cs["ninst"], nwit=cs["nwit"],
              a_query=np.zeros((nv, 12), np.uint64), a_inf=np.zeros(nv, np.uint8),
              b_g1_query=np.zeros((nv, 12), np.uint64), b1_inf=np.zeros(nv, np.uint8),
              b_g2_query=np.zeros((nv, 24), np.uint64), b2_inf=np.zeros(nv, np.uint8),
              h_query=np.zeros((D - 1, 12), np.uint64), h_inf=np.zeros(D - 1, np.uint8),
              l_query=np.zeros((cs["nwit"], 12), np.uint64), l_inf=np.zeros(cs["nwit"], np.uint8),
              vk_g1=np.zeros((3, 12), np.uint64), vk_g2=np.zeros((3, 24), np.uint64),
              gamma_abc_g1=np.zeros((cs["ninst"], 12), np.uint64))
    rp, col, cf, _ = _r1cs_ptrs(cs)
    toxic_mont = np.ascontiguousarray(toxic_mont, np.uint64).reshape(7, 4)
    ok = lib().orc_groth16_setup_r1cs(C.c_size_t(cs["ncons"]), C.c_size_t(cs["ninst"]), C.c_size_t(cs["nwit"]), rp, col, cf,
                                      _p(toxic_mont), _p(pk["a_query"]), _p8(pk["a_inf"]), _p(pk["b_g1_query"]), _p8(pk["b1_inf"]),
                                      _p(pk["b_g2_query"]), _p8(pk["b2_inf"]), _p(pk["h_query"]), _p8(pk["h_inf"]),
                                      _p(pk["l_query"]), _p8(pk["l_inf"]), _p(pk["vk_g1"]), _p(pk["vk_g2"]), _p(pk["gamma_abc_g1"]),
                                      C.c_int(threads))
    assert ok
    return pk


def groth16_prove_r1cs(scheme, cs: dict, full_shares, r_sh, s_sh, pk, threads=1, want_h=True):
    """full_shares: per party, (ninst + nwit, 4) shares of [instance, witness]; entry 0 is the lowered constant one."""
    n = len(full_shares)
    D = pk["D"]
    arrs, ptrs = _chain_ptrs(full_shares)
    r_sh = np.ascontiguousarray(r_sh, np.uint64).reshape(n, 4)
    s_sh = np.ascontiguousarray(s_sh, np.uint64).reshape(n, 4)
    h = np.zeros((n, D, 4), np.uint64) if want_h else None
    proof_sh = np.zeros((n, 48), np.uint64)
    proof_sh_inf = np.zeros((n, 3), np.uint8)
    proof = np.zeros(48, np.uint64)
    proof_inf = np.zeros(3, np.uint8)
    rp, col, cf, _ = _r1cs_ptrs(cs)
    ok = lib().orc_groth16_prove_r1cs(C.c_int(scheme), C.c_int(n), C.c_size_t(cs["ncons"]), C.c_size_t(cs["ninst"]),
                                      C.c_size_t(cs["nwit"]), rp, col, cf, ptrs, _p(r_sh), _p(s_sh),
                                      _p(pk["a_query"]), _p8(pk["a_inf"]), _p(pk["b_g1_query"]), _p8(pk["b1_inf"]),
                                      _p(pk["b_g2_query"]), _p8(pk["b2_inf"]), _p(pk["h_query"]), _p8(pk["h_inf"]),
                                      _p(pk["l_query"]), _p8(pk["l_inf"]), _p(pk["vk_g1"]), _p(pk["vk_g2"]),
                                      _p(h) if want_h else None, _p(proof_sh), _p8(proof_sh_inf), _p(proof), _p8(proof_inf),
                                      C.c_int(threads))
    return dict(ok=bool(ok), h=h, proof_sh=proof_sh, proof_sh_inf=proof_sh_inf, proof=proof, proof_inf=proof_inf)


def r1cs_full_shares(z_ints, n_parties: int, seed: int, scheme):
    """Shares of the full assignment in the reference's lowering: additive shares of every variable except the constant one,
    which is Public(1) lowered to 'the king holds 1' (plain: the values themselves)."""
    z = fr_from_ints(z_ints)
    if scheme == SCHEME_PLAIN:
        return [z]
    sh = [a.copy() for a in king_share_batch(z, n_parties, seed)]
    one = fr_from_ints([1])[0]
    for p in range(n_parties):
        sh[p][0] = one if p == 0 else 0
    return sh

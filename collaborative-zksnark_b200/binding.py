"""ctypes binding of libczk_b200.so (include/czk.h) with a thin object layer.

Names follow the reference's interface for this path:
  Context.msm_g1 / msm_g2      <- AffineCurve::multi_scalar_mul   (algebra/ec/src/lib.rs:302-311)
  Context.fft / ifft / coset_fft / coset_ifft  (+ *_in_place on DeviceVec)
                               <- EvaluationDomain               (algebra/poly/src/domain/mod.rs:79-158)
  Context.batch_open / batch_mul <- FieldShare                   (mpc-algebra/src/share/field.rs:44-46,97-127)
  Context.net_*                <- MpcNet                         (mpc-net/src/lib.rs:28-70)
All arrays are numpy uint64 limb arrays in the reference's Montgomery in-memory form.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

SCHEME_PLAIN, SCHEME_ADDITIVE, SCHEME_SPDZ, SCHEME_GSZ = 0, 1, 2, 3
NTT_FFT, NTT_IFFT, NTT_COSET_FFT, NTT_COSET_IFFT, NTT_IFFT_COSET_FFT = 0, 1, 2, 3, 4

u64p = C.POINTER(C.c_uint64)
u8p = C.POINTER(C.c_uint8)


class CzkError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"czk error {code}: {msg}")
        self.code = code


def library_path() -> Path:
    # CZK_B200_LIB: another build of the same library (A/B runs of kernel variants on one GPU box)
    override = os.environ.get("CZK_B200_LIB")
    return Path(override) if override else _HERE / "libczk_b200.so"


_SIGS = {
    "czk_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "czk_ctx_destroy": (None, [C.c_void_p]),
    "czk_last_error": (C.c_char_p, [C.c_void_p]),
    "czk_ctx_sync": (C.c_int, [C.c_void_p]),
    "czk_ctx_stream": (C.c_void_p, [C.c_void_p]),
    "czk_ctx_launches": (C.c_uint64, [C.c_void_p]),
    "czk_version": (C.c_char_p, []),
    "czk_vec_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "czk_vec_free": (None, [C.c_void_p, C.c_void_p]),
    "czk_vec_len": (C.c_size_t, [C.c_void_p]),
    "czk_vec_device_ptr": (C.c_void_p, [C.c_void_p]),
    "czk_vec_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "czk_vec_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "czk_vec_copy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t]),
    "czk_vec_zero": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]),
    "czk_ntt_fr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint, C.c_int, C.c_int]),
    "czk_ntt_fr_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint, C.c_int, C.c_int]),
    "czk_ntt_vec": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint, C.c_int, C.c_int]),
    "czk_ntt_fr_batch": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_uint, C.c_int]),
    "czk_ntt_vec_batch": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_uint, C.c_int]),
    "czk_domain_params": (C.c_int, [C.c_uint, u64p, u64p, u64p, u64p]),
    "czk_ntt_mixed_fr": (C.c_int, [C.c_void_p, u64p, C.c_uint, C.c_int, C.c_int]),
    "czk_ntt_mixed_fr_batch": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_uint, C.c_int]),
    "czk_ntt_mixed_vec_batch": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_uint, C.c_int]),
    "czk_mixed_domain_params": (C.c_int, [C.c_uint, u64p, u64p, u64p, u64p]),
    "czk_vec_add": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "czk_vec_sub": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "czk_vec_mul": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "czk_vec_scale": (C.c_int, [C.c_void_p, C.c_void_p, u64p, C.c_size_t]),
    "czk_vec_distribute_powers": (C.c_int, [C.c_void_p, C.c_void_p, u64p, u64p, C.c_size_t]),
    "czk_vec_divide_by_vanishing_on_coset": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint]),
    "czk_msm_g1": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, u64p]),
    "czk_msm_g2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, u64p]),
    "czk_bases_upload": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "czk_bases_free": (None, [C.c_void_p, C.c_void_p]),
    "czk_bases_len": (C.c_size_t, [C.c_void_p]),
    "czk_bases_precompute": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint]),
    "czk_bases_device_bytes": (C.c_int, [C.c_void_p, u64p]),
    "czk_msm_bases": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_size_t, u64p]),
    "czk_msm_bases_multi": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_size_t,
                                      C.POINTER(u64p), C.POINTER(C.c_double)]),
    "czk_bases_synthetic": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_size_t, C.c_size_t, C.POINTER(C.c_void_p)]),
    "czk_bases_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p]),
    "czk_net_unique_id": (C.c_int, [C.c_void_p]),
    "czk_net_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "czk_net_init_single": (C.c_int, [C.c_void_p]),
    "czk_net_deinit": (None, [C.c_void_p]),
    "czk_net_party_id": (C.c_int, [C.c_void_p]),
    "czk_net_n_parties": (C.c_int, [C.c_void_p]),
    "czk_net_allgather_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "czk_net_allgather_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "czk_net_bcast_from_king_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "czk_net_gather_to_king_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "czk_net_stats": (C.c_int, [C.c_void_p, u64p]),
    "czk_net_reset_stats": (None, [C.c_void_p]),
    "czk_batch_open": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "czk_beaver_batch_mul": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "czk_vec_prefix_products": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "czk_vec_batch_inverse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "czk_poly_div_linear": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, u64p, C.c_void_p, u64p]),
    "czk_share_batch_inv": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]),
    "czk_share_batch_div": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "czk_share_partial_products": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]),
    "czk_kzg_open": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, u64p, u64p, u64p]),
    "czk_fr_serialize": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p]),
    "czk_fr_deserialize": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p]),
    "czk_g1_serialize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "czk_g2_serialize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "czk_g1_deserialize": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "czk_g2_deserialize": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "czk_gsz_open": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p, C.c_size_t]),
    "czk_gsz_king_compute": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint, C.c_size_t]),
    "czk_gsz_batch_mul": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]),
    "czk_gsz_check_products": (C.c_int, [C.c_void_p, u64p]),
    "czk_gsz_stats": (C.c_int, [C.c_void_p, u64p]),
    "czk_net_link_bytes": (C.c_int, [C.c_void_p, u64p]),
    "czk_net_share_transport": (C.c_int, [C.c_void_p]),
    "czk_diag_sim_batch_open": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                          C.c_size_t, C.POINTER(C.c_uint32)]),
    "czk_diag_sim_beaver_mul": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                          C.POINTER(C.c_void_p), C.c_size_t, C.POINTER(C.c_uint32)]),
    "czk_diag_gsz_open_gathered": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_uint, C.c_size_t, C.c_void_p, C.POINTER(C.c_uint32)]),
    "czk_msm_set_batched": (C.c_int, [C.c_void_p, C.c_int]),
    "czk_fq_inverse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "czk_msm_stats": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_int]),
    "czk_microbench": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}

class PlonkTranscript(C.Structure):
    """czk_plonk_transcript (include/czk_plonk.h): the caller's Fiat-Shamir transcript as two callbacks."""
    _fields_ = [("user", C.c_void_p), ("absorb_g1", C.c_void_p), ("challenge", C.c_void_p)]


class PlonkWiringProof(C.Structure):
    """czk_plonk_wiring_proof: commitments l1, t, q, l2_q; openings t(wr) t(r) t(w^(k-1)) l1(wr) q(r) l2_q(x) w(x) l1(x) p(x)."""
    _fields_ = [("cmt_xy", (C.c_uint64 * 12) * 4), ("cmt_inf", C.c_uint8 * 4), ("open_val", (C.c_uint64 * 4) * 9),
                ("open_pf_xy", (C.c_uint64 * 12) * 9), ("open_pf_inf", C.c_uint8 * 9), ("challenges", (C.c_uint64 * 4) * 4)]

    def to_dict(self):
        return dict(cmt_xy=np.array(self.cmt_xy, np.uint64).reshape(4, 12), cmt_inf=np.array(self.cmt_inf, np.uint8),
                    open_val=np.array(self.open_val, np.uint64).reshape(9, 4), open_pf_xy=np.array(self.open_pf_xy, np.uint64).reshape(9, 12),
                    open_pf_inf=np.array(self.open_pf_inf, np.uint8), challenges=np.array(self.challenges, np.uint64).reshape(4, 4))


# include/czk_groth16.h, include/czk_plonk.h
_OPTIONAL_SIGS = {
    "czk_plonk_standin_transcript": (None, [C.POINTER(C.c_uint64), C.c_uint64, C.POINTER(PlonkTranscript)]),
    "czk_plonk_prove_wiring": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.POINTER(PlonkTranscript), C.POINTER(PlonkWiringProof), C.POINTER(PlonkWiringProof),
                                         C.POINTER(C.c_double)]),
    "czk_plonk_prove_wiring_mixed": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.POINTER(PlonkTranscript), C.POINTER(PlonkWiringProof), C.POINTER(PlonkWiringProof),
                                               C.POINTER(C.c_double)]),
    "czk_groth16_pk_upload": (C.c_int, [C.c_void_p, C.c_size_t] + [C.c_void_p] * 12 + [C.POINTER(C.c_void_p)]),
    "czk_groth16_pk_synthetic": (C.c_int, [C.c_void_p, C.c_size_t, C.c_uint64, C.POINTER(C.c_void_p)]),
    "czk_groth16_pk_free": (None, [C.c_void_p, C.c_void_p]),
    "czk_groth16_pk_domain_size": (C.c_size_t, [C.c_void_p]),
    "czk_groth16_pk_query": (C.c_void_p, [C.c_void_p, C.c_int]),
    "czk_groth16_pk_vk": (C.c_int, [C.c_void_p, u64p, u64p]),
    "czk_groth16_witness_map": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p]),
    "czk_groth16_prove": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, u64p, u64p, u64p, u8p, u64p, u8p]),
    "czk_groth16_prove_vec": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, u64p, u64p, u64p, u8p, u64p, u8p]),
    "czk_squaring_chain": (C.c_int, [u64p, C.c_size_t, C.c_void_p]),
    "czk_king_share_batch": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_uint64, C.c_void_p]),
    "czk_groth16_last_phases": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "czk_groth16_proof_serialize": (C.c_int, [u64p, u8p, C.c_void_p]),
    "czk_groth16_proof_deserialize": (C.c_int, [C.c_void_p, u64p, u8p]),
    "czk_pairing_product_is_one": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int)]),
    "czk_groth16_verify": (C.c_int, [u64p, u64p, C.c_void_p, C.c_size_t, C.c_void_p, u64p, u8p, C.POINTER(C.c_int)]),
    "czk_groth16_setup": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_void_p)]),
    "czk_groth16_setup_r1cs": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "czk_groth16_pk_gamma_abc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "czk_fixed_base_msm": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.POINTER(C.c_void_p)]),
    "czk_r1cs_upload": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "czk_r1cs_free": (None, [C.c_void_p, C.c_void_p]),
    "czk_groth16_pk_upload_r1cs": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t] + [C.c_void_p] * 12 + [C.POINTER(C.c_void_p)]),
    "czk_groth16_prove_r1cs": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, u64p, u64p, u64p, u8p, u64p, u8p]),
    "czk_groth16_gsz_last_checks": (C.c_int, [C.c_void_p, u64p, u64p, u64p, u8p, u64p]),
}


def exported_symbols():
    """Every symbol include/czk.h declares (used by the CPU-tier load test)."""
    return sorted(_SIGS)


def load_library():
    """Load libczk_b200.so.  Raises if it has not been built: there is no fallback implementation."""
    global _LIB
    if _LIB is not None:
        return _LIB
    so = library_path()
    if not so.exists():
        raise FileNotFoundError(
            f"{so} is missing - build it with `python collaborative-zksnark_b200/build.py` "
            "(__graft_entry__.build()); czk_b200 has no CPU fallback")
    lib = C.CDLL(str(so))
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    for name, (res, args) in _OPTIONAL_SIGS.items():
        if hasattr(lib, name):
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
    _LIB = lib
    return lib


def _np_u64(a, shape_last=None):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if shape_last is not None:
        a = a.reshape(-1, shape_last)
    return a


def domain_params(log_d: int):
    lib = load_library()
    outs = [np.zeros(4, np.uint64) for _ in range(4)]
    rc = lib.czk_domain_params(log_d, *[o.ctypes.data_as(u64p) for o in outs])
    if rc:
        raise CzkError(rc, lib.czk_last_error(None).decode())
    return dict(group_gen=outs[0], group_gen_inv=outs[1], size_inv=outs[2], generator_inv=outs[3])


def mixed_domain_params(log_m: int):
    """Constants of MixedRadixEvaluationDomain::new(3 * 2^log_m)."""
    lib = load_library()
    outs = [np.zeros(4, np.uint64) for _ in range(4)]
    rc = lib.czk_mixed_domain_params(log_m, *[o.ctypes.data_as(u64p) for o in outs])
    if rc:
        raise CzkError(rc, lib.czk_last_error(None).decode())
    return dict(group_gen=outs[0], group_gen_inv=outs[1], size_inv=outs[2], generator_inv=outs[3])


class DeviceVec:
    """Device-resident Vec<Fr> (Montgomery limbs)."""

    def __init__(self, ctx: "Context", n: int):
        self.ctx = ctx
        self.n = n
        h = C.c_void_p()
        ctx._chk(ctx.lib.czk_vec_alloc(ctx.h, n, C.byref(h)))
        self.h = h

    @classmethod
    def from_numpy(cls, ctx, arr, n=None):
        arr = _np_u64(arr, 4)
        v = cls(ctx, n if n is not None else arr.shape[0])
        v.upload(arr)
        return v

    def upload(self, arr, offset=0):
        arr = _np_u64(arr, 4)
        self.ctx._chk(self.ctx.lib.czk_vec_upload(self.ctx.h, self.h, offset, arr.ctypes.data, arr.shape[0]))

    def numpy(self, offset=0, n=None):
        n = self.n - offset if n is None else n
        out = np.empty((n, 4), np.uint64)
        self.ctx._chk(self.ctx.lib.czk_vec_download(self.ctx.h, self.h, offset, out.ctypes.data, n))
        return out

    def clone(self):
        v = DeviceVec(self.ctx, self.n)
        self.ctx._chk(self.ctx.lib.czk_vec_copy(self.ctx.h, v.h, 0, self.h, 0, self.n))
        return v

    @property
    def device_ptr(self) -> int:
        return self.ctx.lib.czk_vec_device_ptr(self.h)

    def free(self):
        if self.h:
            self.ctx.lib.czk_vec_free(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Bases:
    """Device-resident affine bases (one CRS query)."""

    def __init__(self, ctx, h, curve):
        self.ctx, self.h, self.curve = ctx, h, curve

    def __len__(self):
        return self.ctx.lib.czk_bases_len(self.h)

    def numpy(self, off=0, n=None):
        n = len(self) - off if n is None else n
        w = 12 if self.curve == 1 else 24
        xy = np.empty((n, w), np.uint64)
        inf = np.empty(n, np.uint8)
        self.ctx._chk(self.ctx.lib.czk_bases_download(self.ctx.h, self.h, off, n, xy.ctypes.data, inf.ctypes.data))
        return xy, inf

    def device_bytes(self) -> dict:
        out = np.zeros(2, np.uint64)
        self.ctx._chk(self.ctx.lib.czk_bases_device_bytes(self.h, out.ctypes.data_as(u64p)))
        return {"points": int(out[0]), "table": int(out[1])}

    def precompute(self, c: int = 0):
        """Build the merged-window table (2^(c w) * P_i); later msm_bases calls use it."""
        self.ctx._chk(self.ctx.lib.czk_bases_precompute(self.ctx.h, self.h, c))
        return self

    def free(self):
        if self.h:
            self.ctx.lib.czk_bases_free(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One party pinned to one GPU."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.czk_ctx_create(device, C.byref(h))
        if rc:
            raise CzkError(rc, self.lib.czk_last_error(None).decode())
        self.h = h
        self.device = device

    def _chk(self, rc):
        if rc:
            raise CzkError(rc, self.lib.czk_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.lib.czk_ctx_destroy(self.h)
            self.h = None

    def sync(self):
        self._chk(self.lib.czk_ctx_sync(self.h))

    @property
    def stream(self) -> int:
        return self.lib.czk_ctx_stream(self.h)

    @property
    def launches(self) -> int:
        return self.lib.czk_ctx_launches(self.h)

    # ------------------------------------------------------------------ vectors
    def vec(self, n):
        return DeviceVec(self, n)

    def vec_from(self, arr, n=None):
        return DeviceVec.from_numpy(self, arr, n)

    # ------------------------------------------------------------------ EvaluationDomain (host buffers: e2e path)
    def _ntt_host(self, data, inverse, coset):
        a = np.array(data, dtype=np.uint64, order="C").reshape(-1, 4)
        n = a.shape[0]
        log_d = max(n - 1, 0).bit_length()
        if (1 << log_d) != n:  # resize with zeros (radix2/mod.rs:100-101)
            a = np.concatenate([a, np.zeros(((1 << log_d) - n, 4), np.uint64)])
        self._chk(self.lib.czk_ntt_fr(self.h, a.ctypes.data, log_d, int(inverse), int(coset)))
        return a

    def fft(self, data):
        return self._ntt_host(data, False, False)

    def ifft(self, data):
        return self._ntt_host(data, True, False)

    def coset_fft(self, data):
        return self._ntt_host(data, False, True)

    def coset_ifft(self, data):
        return self._ntt_host(data, True, True)

    # device-resident, in place
    def ntt_in_place(self, v: DeviceVec, log_d: int, inverse=False, coset=False):
        self._chk(self.lib.czk_ntt_vec(self.h, v.h, log_d, int(inverse), int(coset)))

    def vec_add(self, a, b, n=None):
        self._chk(self.lib.czk_vec_add(self.h, a.h, b.h, a.n if n is None else n))

    def vec_sub(self, a, b, n=None):
        self._chk(self.lib.czk_vec_sub(self.h, a.h, b.h, a.n if n is None else n))

    def vec_mul(self, a, b, n=None):
        self._chk(self.lib.czk_vec_mul(self.h, a.h, b.h, a.n if n is None else n))

    def vec_scale(self, a, c, n=None):
        c = _np_u64(c)
        self._chk(self.lib.czk_vec_scale(self.h, a.h, c.ctypes.data_as(u64p), a.n if n is None else n))

    def distribute_powers(self, a, g, c, n=None):
        g, c = _np_u64(g), _np_u64(c)
        self._chk(self.lib.czk_vec_distribute_powers(self.h, a.h, g.ctypes.data_as(u64p), c.ctypes.data_as(u64p), a.n if n is None else n))

    def divide_by_vanishing_on_coset(self, a, log_d):
        self._chk(self.lib.czk_vec_divide_by_vanishing_on_coset(self.h, a.h, log_d))

    # ------------------------------------------------------------------ MSM
    def _msm_host(self, fn, w, bases_xy, inf, scalars, montgomery):
        bases_xy = _np_u64(bases_xy, w)
        scalars = _np_u64(scalars, 4)
        n = min(bases_xy.shape[0], scalars.shape[0])  # variable_base.rs:16
        infp = None
        if inf is not None:
            inf = np.ascontiguousarray(inf, dtype=np.uint8)
            infp = inf.ctypes.data
        out = np.zeros(3 * (w // 2), np.uint64)
        self._chk(fn(self.h, bases_xy.ctypes.data, infp, scalars.ctypes.data, int(montgomery), n, out.ctypes.data_as(u64p)))
        return out

    def msm_g1(self, bases_xy, inf, scalars, montgomery=True):
        """-> Jacobian (x, y, z) limbs, affine-normalised (z = 1) or (1, 1, 0)."""
        return self._msm_host(self.lib.czk_msm_g1, 12, bases_xy, inf, scalars, montgomery)

    def msm_g2(self, bases_xy, inf, scalars, montgomery=True):
        return self._msm_host(self.lib.czk_msm_g2, 24, bases_xy, inf, scalars, montgomery)

    def bases_upload(self, curve, bases_xy, inf=None):
        w = 12 if curve == 1 else 24
        bases_xy = _np_u64(bases_xy, w)
        infp = None
        if inf is not None:
            inf = np.ascontiguousarray(inf, dtype=np.uint8)
            infp = inf.ctypes.data
        h = C.c_void_p()
        self._chk(self.lib.czk_bases_upload(self.h, curve, bases_xy.ctypes.data, infp, bases_xy.shape[0], C.byref(h)))
        return Bases(self, h, curve)

    def bases_synthetic(self, curve, seed, n, inf_every=0):
        h = C.c_void_p()
        self._chk(self.lib.czk_bases_synthetic(self.h, curve, seed, n, inf_every, C.byref(h)))
        return Bases(self, h, curve)

    def msm_bases(self, bases: Bases, scalars: DeviceVec, n=None, base_off=0, sc_off=0, montgomery=True):
        n = min(len(bases) - base_off, scalars.n - sc_off) if n is None else n
        w = 12 if bases.curve == 1 else 24
        out = np.zeros(3 * (w // 2), np.uint64)
        self._chk(self.lib.czk_msm_bases(self.h, bases.h, base_off, scalars.h, sc_off, int(montgomery), n, out.ctypes.data_as(u64p)))
        return out

    def msm_bases_multi(self, bases_list, scalars: DeviceVec, n=None, base_off=0, sc_off=0, montgomery=True):
        """One scalar vector against several resident base sets (shared digit sort where the sets allow it)."""
        n = min(min(len(b) for b in bases_list) - base_off, scalars.n - sc_off) if n is None else n
        outs = [np.zeros(18 if b.curve == 1 else 36, np.uint64) for b in bases_list]
        hs = (C.c_void_p * len(bases_list))(*[b.h for b in bases_list])
        ps = (u64p * len(bases_list))(*[o.ctypes.data_as(u64p) for o in outs])
        self._chk(self.lib.czk_msm_bases_multi(self.h, hs, len(bases_list), base_off, scalars.h, sc_off, int(montgomery), n, ps, None))
        return outs

    # ------------------------------------------------------------------ MpcNet
    def net_init(self, rank, nranks, unique_id: bytes | None):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id) if unique_id is not None else None
        self._chk(self.lib.czk_net_init(self.h, rank, nranks, buf))

    def net_unique_id(self) -> bytes:
        buf = (C.c_uint8 * 128)()
        rc = self.lib.czk_net_unique_id(buf)
        if rc:
            raise CzkError(rc, self.lib.czk_last_error(None).decode())
        return bytes(buf)

    @property
    def party_id(self):
        return self.lib.czk_net_party_id(self.h)

    @property
    def n_parties(self):
        return self.lib.czk_net_n_parties(self.h)

    def am_king(self):
        return self.party_id == 0

    def broadcast_bytes(self, data: bytes):
        """MpcNet::broadcast_bytes: every party's message, in party order."""
        n = len(data)
        send = np.frombuffer(data, dtype=np.uint8).copy()
        recv = np.empty(n * self.n_parties, np.uint8)
        self._chk(self.lib.czk_net_allgather_host(self.h, send.ctypes.data, recv.ctypes.data, n))
        return [recv[i * n:(i + 1) * n].tobytes() for i in range(self.n_parties)]

    def net_stats(self):
        out = np.zeros(5, np.uint64)
        self.lib.czk_net_stats(self.h, out.ctypes.data_as(u64p))
        return dict(zip(("bytes_sent", "bytes_recv", "broadcasts", "to_king", "from_king"), (int(x) for x in out)))

    def net_reset_stats(self):
        self.lib.czk_net_reset_stats(self.h)

    # ------------------------------------------------------------------ FieldShare
    def batch_open(self, scheme, sh: DeviceVec, mac: DeviceVec | None = None, n=None):
        n = sh.n if n is None else n
        out = DeviceVec(self, n)
        self._chk(self.lib.czk_batch_open(self.h, scheme, sh.h, mac.h if mac is not None else None, out.h, n))
        return out

    def batch_mul(self, scheme, x_sh, x_mac, y_sh, y_mac, n=None):
        """x *= y on shares (Field::batch_product_in_place on MpcField, wire/field.rs:358-393)."""
        n = x_sh.n if n is None else n
        self._chk(self.lib.czk_beaver_batch_mul(self.h, scheme, x_sh.h, x_mac.h if x_mac is not None else None, y_sh.h,
                                                y_mac.h if y_mac is not None else None, n))

    @property
    def share_transport(self) -> str:
        return {1: "nvlink peer memory (CUDA IPC)", -1: "nccl send/recv + all-gather", 0: "undecided / single party"}[
            self.lib.czk_net_share_transport(self.h)]

    def ntt_batch(self, vecs, log_d: int, op: int):
        """The same transform over several device vectors in one grid per pass (czk_ntt_vec_batch).
        op: NTT_FFT / NTT_IFFT / NTT_COSET_FFT / NTT_COSET_IFFT / NTT_IFFT_COSET_FFT."""
        self._chk(self.lib.czk_ntt_vec_batch(self.h, self._handles(vecs), len(vecs), log_d, op))

    def ntt_mixed(self, data, inverse=False, coset=False) -> np.ndarray:
        """Host array of 3 * 2^k Montgomery Fr elements through czk_ntt_mixed_fr; returns a transformed copy."""
        a = np.array(data, dtype=np.uint64, order="C").reshape(-1, 4)
        n = a.shape[0]
        log_m = (n // 3).bit_length() - 1
        assert n == 3 << log_m, "length must be 3 * 2^k"
        self._chk(self.lib.czk_ntt_mixed_fr(self.h, a.ctypes.data_as(u64p), log_m, int(inverse), int(coset)))
        return a

    def ntt_mixed_batch(self, vecs, log_m: int, op: int):
        """MixedRadixEvaluationDomain transforms over 3 * 2^log_m points (czk_ntt_mixed_vec_batch), same ops."""
        self._chk(self.lib.czk_ntt_mixed_vec_batch(self.h, self._handles(vecs), len(vecs), log_m, op))

    def net_link_bytes(self):
        out = np.zeros(2, np.uint64)
        self.lib.czk_net_link_bytes(self.h, out.ctypes.data_as(u64p))
        return {"sent": int(out[0]), "received": int(out[1])}

    @staticmethod
    def _handles(vecs):
        if vecs is None:
            return None
        return (C.c_void_p * len(vecs))(*[v.h for v in vecs])

    def sim_batch_open(self, scheme, sh, mac=None, n=None):
        """N-party batch_open simulated on this GPU (diagnostics): sh / mac are lists of DeviceVec, one per party.
        Returns (opened vector of every party, per-party MAC-check flags)."""
        parties = len(sh)
        n = sh[0].n if n is None else n
        out = [DeviceVec(self, n) for _ in range(parties)]
        flags = (C.c_uint32 * parties)()
        self._chk(self.lib.czk_diag_sim_batch_open(self.h, scheme, parties, self._handles(sh), self._handles(mac), self._handles(out), n, flags))
        return out, list(flags)

    def sim_batch_mul(self, scheme, x_sh, x_mac, y_sh, y_mac, n=None):
        """N-party Beaver product simulated on this GPU (diagnostics): x_sh[q] (x_mac[q]) *= y for every party q."""
        parties = len(x_sh)
        n = x_sh[0].n if n is None else n
        flags = (C.c_uint32 * parties)()
        self._chk(self.lib.czk_diag_sim_beaver_mul(self.h, scheme, parties, self._handles(x_sh), self._handles(x_mac), self._handles(y_sh),
                                                   self._handles(y_mac), n, flags))
        return list(flags)

    def gsz_open_gathered(self, gathered: DeviceVec, parties: int, degree: int, k: int):
        """open_degree_vec on a party-major (parties x k) matrix of gathered shares; returns (values, degree-check flag)."""
        out = DeviceVec(self, k)
        flag = C.c_uint32(0)
        self._chk(self.lib.czk_diag_gsz_open_gathered(self.h, gathered.h, parties, degree, k, out.h, C.byref(flag)))
        return out, int(flag.value)

    # ------------------------------------------------------------------ Plonk / KZG10 leaves
    def prefix_products(self, v: DeviceVec, n=None):
        self._chk(self.lib.czk_vec_prefix_products(self.h, v.h, v.n if n is None else n))

    def batch_inverse(self, v: DeviceVec, n=None):
        self._chk(self.lib.czk_vec_batch_inverse(self.h, v.h, v.n if n is None else n))

    def poly_div_linear(self, p: DeviceVec, z, n=None):
        """(q, rem) = p / (X - z): q a DeviceVec of n - 1 coefficients, rem = p(z) as 4 limbs."""
        n = p.n if n is None else n
        q = DeviceVec(self, max(n - 1, 1))
        rem = np.zeros(4, np.uint64)
        self._chk(self.lib.czk_poly_div_linear(self.h, p.h, n, _np_u64(z).ctypes.data_as(u64p), q.h, rem.ctypes.data_as(u64p)))
        return q, rem

    def share_batch_inv(self, scheme, x_sh, x_mac=None, n=None):
        self._chk(self.lib.czk_share_batch_inv(self.h, scheme, x_sh.h, x_mac.h if x_mac is not None else None, x_sh.n if n is None else n))

    def share_batch_div(self, scheme, x_sh, x_mac, y_sh, y_mac, n=None):
        self._chk(self.lib.czk_share_batch_div(self.h, scheme, x_sh.h, x_mac.h if x_mac is not None else None, y_sh.h,
                                               y_mac.h if y_mac is not None else None, x_sh.n if n is None else n))

    def share_partial_products(self, scheme, x_sh, x_mac=None, n=None):
        self._chk(self.lib.czk_share_partial_products(self.h, scheme, x_sh.h, x_mac.h if x_mac is not None else None,
                                                      x_sh.n if n is None else n))

    def kzg_open(self, powers: "Bases", p: DeviceVec, z, n=None):
        """KZG10::open without hiding on one coefficient vector: (w as Jacobian limbs, eval)."""
        n = p.n if n is None else n
        w = np.zeros(18, np.uint64)
        ev = np.zeros(4, np.uint64)
        self._chk(self.lib.czk_kzg_open(self.h, powers.h, p.h, n, _np_u64(z).ctypes.data_as(u64p), w.ctypes.data_as(u64p),
                                        ev.ctypes.data_as(u64p)))
        return w, ev

    # ------------------------------------------------------------------ GSZ20 shares (share/gsz20/mod.rs)
    def gsz_open(self, sh: DeviceVec, degree: int, n=None):
        n = sh.n if n is None else n
        out = DeviceVec(self, n)
        self._chk(self.lib.czk_gsz_open(self.h, sh.h, degree, out.h, n))
        return out

    def gsz_king_compute(self, v: DeviceVec, degree: int, n=None):
        self._chk(self.lib.czk_gsz_king_compute(self.h, v.h, degree, v.n if n is None else n))

    def gsz_batch_mul(self, x: DeviceVec, y: DeviceVec, n=None, queue_check=True):
        self._chk(self.lib.czk_gsz_batch_mul(self.h, x.h, y.h, x.n if n is None else n, int(queue_check)))

    def gsz_check_products(self):
        """hadamard_check -> ip_check over the queued triples; returns the opened (x, y, z) of the last step."""
        out = np.zeros(12, np.uint64)
        self._chk(self.lib.czk_gsz_check_products(self.h, out.ctypes.data_as(u64p)))
        return out.reshape(3, 4)

    def gsz_stats(self):
        out = np.zeros(2, np.uint64)
        self._chk(self.lib.czk_gsz_stats(self.h, out.ctypes.data_as(u64p)))
        return dict(king_computes=int(out[0]), opens=int(out[1]))

    # ------------------------------------------------------------------ diagnostics
    def msm_set_batched(self, enabled, always: bool = False):
        """enabled: use the batched-affine tree (from ~2^19 terms); always=True: at every size."""
        self._chk(self.lib.czk_msm_set_batched(self.h, 2 if (enabled and always) else int(bool(enabled))))

    def fq_inverse(self, a):
        a = _np_u64(a, 6)
        out = np.zeros_like(a)
        self._chk(self.lib.czk_fq_inverse(self.h, a.ctypes.data, out.ctypes.data, a.shape[0]))
        return out

    def msm_stats(self, curve=1, reset=False):
        out = (C.c_double * 5)()
        self._chk(self.lib.czk_msm_stats(self.h, curve, out, int(reset)))
        return dict(accumulate_ms=out[0], launches=int(out[1]), terms=out[2], msm_ms=out[3], entries=out[4])

    def microbench(self, kind, blocks_per_sm=8, threads=128, iters=2000):
        ops = C.c_double()
        ms = C.c_double()
        self._chk(self.lib.czk_microbench(self.h, kind, blocks_per_sm, threads, iters, C.byref(ops), C.byref(ms)))
        return ops.value, ms.value


class ProvingKey:
    """Device-resident groth16 ProvingKey for the squaring circuit (include/czk_groth16.h)."""

    QUERIES = ("a_query", "b_g1_query", "b_g2_query", "h_query", "l_query")

    def __init__(self, ctx: Context, h, n_sq: int):
        self.ctx, self.h, self.n_sq = ctx, h, n_sq

    @classmethod
    def upload(cls, ctx: Context, pk: dict):
        """pk: dict with the arrays of groth16 setup (a_query, a_inf, ..., vk_g1, vk_g2)."""
        def p(a):
            return np.ascontiguousarray(a).ctypes.data if a is not None else None
        keep = [np.ascontiguousarray(pk[k]) for k in ("a_query", "a_inf", "b_g1_query", "b1_inf", "b_g2_query", "b2_inf",
                                                       "h_query", "h_inf", "l_query", "l_inf", "vk_g1", "vk_g2")]
        h = C.c_void_p()
        ctx._chk(ctx.lib.czk_groth16_pk_upload(ctx.h, pk["n_sq"], *[k.ctypes.data for k in keep], C.byref(h)))
        return cls(ctx, h, pk["n_sq"])

    @classmethod
    def synthetic(cls, ctx: Context, n_sq: int, seed: int = 1):
        h = C.c_void_p()
        ctx._chk(ctx.lib.czk_groth16_pk_synthetic(ctx.h, n_sq, seed, C.byref(h)))
        return cls(ctx, h, n_sq)

    @property
    def domain_size(self):
        return self.ctx.lib.czk_groth16_pk_domain_size(self.h)

    def query(self, which: int) -> "Bases":
        b = Bases(self.ctx, C.c_void_p(self.ctx.lib.czk_groth16_pk_query(self.h, which)), 2 if which == 2 else 1)
        b.free = lambda: None  # owned by the key
        return b

    def to_host(self) -> dict:
        """Download everything (for comparisons in tests)."""
        out = dict(n_sq=self.n_sq, D=self.domain_size)
        for i, (name, inf) in enumerate(zip(self.QUERIES, ("a_inf", "b1_inf", "b2_inf", "h_inf", "l_inf"))):
            xy, f = self.query(i).numpy()
            out[name], out[inf] = xy, f
        v1 = np.zeros((3, 12), np.uint64)
        v2 = np.zeros((3, 24), np.uint64)
        self.ctx._chk(self.ctx.lib.czk_groth16_pk_vk(self.h, v1.ctypes.data_as(u64p), v2.ctypes.data_as(u64p)))
        out["vk_g1"], out["vk_g2"] = v1, v2
        return out

    def free(self):
        if self.h:
            self.ctx.lib.czk_groth16_pk_free(self.ctx.h, self.h)
            self.h = None


def plonk_prove_wiring(ctx: "Context", scheme: int, powers: "Bases", log_d: int, p_sh: "DeviceVec", p_mac, w_pub: "DeviceVec",
                       seed: int = 0, transcript: "PlonkTranscript | None" = None, mixed: bool = False) -> dict:
    """Prover::prove_wiring (mpc-plonk/src/lib.rs:199-258) on this party's shares; the stand-in transcript seeded with
    `seed` unless the caller supplies its own callbacks.  mixed: the domain has 3 * 2^log_d points (the reference's
    MixedRadixEvaluationDomain wire domain) instead of 2^log_d.  Returns the revealed proof, this party's proof shares and
    the phase times (transforms + share protocols, commitments, openings, reveal)."""
    state = C.c_uint64(0)
    tr = transcript
    if tr is None:
        tr = PlonkTranscript()
        ctx.lib.czk_plonk_standin_transcript(C.byref(state), seed, C.byref(tr))
    share, out = PlonkWiringProof(), PlonkWiringProof()
    ph = (C.c_double * 4)()
    fn = ctx.lib.czk_plonk_prove_wiring_mixed if mixed else ctx.lib.czk_plonk_prove_wiring
    ctx._chk(fn(ctx.h, scheme, powers.h, log_d, p_sh.h, p_mac.h if p_mac is not None else None, w_pub.h,
                C.byref(tr), C.byref(share), C.byref(out), ph))
    return dict(proof=out.to_dict(), proof_share=share.to_dict(),
                phases_ms=dict(zip(("transforms_and_shares", "commitments", "openings", "reveal"), list(ph))))


def groth16_witness_map(ctx: Context, scheme: int, n_sq: int, chain_sh) -> np.ndarray:
    chain_sh = _np_u64(chain_sh, 4)
    assert chain_sh.shape[0] == n_sq + 1
    d = 1
    while d < n_sq + 2:
        d <<= 1
    h = np.empty((d, 4), np.uint64)
    ctx._chk(ctx.lib.czk_groth16_witness_map(ctx.h, scheme, n_sq, chain_sh.ctypes.data, h.ctypes.data))
    return h


def squaring_chain(start_mont, n_sq: int) -> np.ndarray:
    lib = load_library()
    out = np.empty((n_sq + 1, 4), np.uint64)
    rc = lib.czk_squaring_chain(_np_u64(start_mont).ctypes.data_as(u64p), n_sq, out.ctypes.data)
    if rc:
        raise CzkError(rc, lib.czk_last_error(None).decode())
    return out


def king_share_batch(values_mont, n_parties: int, seed: int) -> np.ndarray:
    """-> (n_parties, k, 4) additive shares (Reveal::king_share_batch)."""
    lib = load_library()
    values_mont = _np_u64(values_mont, 4)
    k = values_mont.shape[0]
    out = np.empty((n_parties, k, 4), np.uint64)
    rc = lib.czk_king_share_batch(values_mont.ctypes.data, k, n_parties, seed, out.ctypes.data)
    if rc:
        raise CzkError(rc, lib.czk_last_error(None).decode())
    return out


def groth16_prove(ctx: Context, scheme: int, pk: ProvingKey, chain_sh, r_sh, s_sh) -> dict:
    """create_random_proof + reveal for this party (mpc-snarks/src/proof.rs:130-139).
    chain_sh: host array (n_sq + 1, 4) - copied to the device inside the call - or a DeviceVec."""
    dev = isinstance(chain_sh, DeviceVec)
    if not dev:
        chain_sh = _np_u64(chain_sh, 4)
        assert chain_sh.shape[0] == pk.n_sq + 1
    r_sh, s_sh = _np_u64(r_sh), _np_u64(s_sh)
    proof_sh = np.zeros(48, np.uint64)
    proof = np.zeros(48, np.uint64)
    sh_inf = np.zeros(3, np.uint8)
    inf = np.zeros(3, np.uint8)
    fn = ctx.lib.czk_groth16_prove_vec if dev else ctx.lib.czk_groth16_prove
    ctx._chk(fn(ctx.h, scheme, pk.h, chain_sh.h if dev else chain_sh.ctypes.data, r_sh.ctypes.data_as(u64p),
                s_sh.ctypes.data_as(u64p), proof_sh.ctypes.data_as(u64p), sh_inf.ctypes.data_as(u8p), proof.ctypes.data_as(u64p),
                inf.ctypes.data_as(u8p)))
    ph = (C.c_double * 8)()
    ctx.lib.czk_groth16_last_phases(ctx.h, ph)
    names = ("upload", "witness_map", "msm_h", "msm_l", "msm_a", "msm_b_g1", "msm_b_g2", "group_tail_reveal")
    res = dict(proof_sh=proof_sh, proof_sh_inf=sh_inf, proof=proof, proof_inf=inf, phases_ms=dict(zip(names, list(ph))))
    if scheme == SCHEME_GSZ:
        f3, gx, gyz = np.zeros(12, np.uint64), np.zeros(4, np.uint64), np.zeros(24, np.uint64)
        ginf, cnt = np.zeros(2, np.uint8), np.zeros(2, np.uint64)
        ctx._chk(ctx.lib.czk_groth16_gsz_last_checks(ctx.h, f3.ctypes.data_as(u64p), gx.ctypes.data_as(u64p), gyz.ctypes.data_as(u64p),
                                                     ginf.ctypes.data_as(u8p), cnt.ctypes.data_as(u64p)))
        res.update(field_check=f3.reshape(3, 4), group_check_x=gx, group_check_yz=gyz.reshape(2, 12), group_check_inf=ginf,
                   king_computes=int(cnt[0]), opens=int(cnt[1]))
    return res


class R1cs:
    """Device-resident constraint matrices (czk_r1cs): cs = dict(ncons, ninst, nwit, a=(row_ptr u64, col u32, coeff (nnz,4)), b=..., c=...)."""

    def __init__(self, ctx: Context, cs: dict):
        self.ctx, self.dims = ctx, (cs["ncons"], cs["ninst"], cs["nwit"])
        mats = [tuple(np.ascontiguousarray(x) for x in cs[m]) for m in ("a", "b", "c")]
        rp = (C.c_void_p * 3)(*[m[0].astype(np.uint64, copy=False).ctypes.data for m in mats])
        col = (C.c_void_p * 3)(*[m[1].astype(np.uint32, copy=False).ctypes.data for m in mats])
        cf = (C.c_void_p * 3)(*[m[2].astype(np.uint64, copy=False).ctypes.data for m in mats])
        self._keep = mats
        h = C.c_void_p()
        ctx._chk(ctx.lib.czk_r1cs_upload(ctx.h, cs["ncons"], cs["ninst"], cs["nwit"], rp, col, cf, C.byref(h)))
        self.h = h

    def free(self):
        if self.h:
            self.ctx.lib.czk_r1cs_free(self.ctx.h, self.h)
            self.h = None


def groth16_pk_upload_r1cs(ctx: Context, pk: dict) -> "ProvingKey":
    def p(a):
        return None if a is None else np.ascontiguousarray(a).ctypes.data

    h = C.c_void_p()
    ctx._chk(ctx.lib.czk_groth16_pk_upload_r1cs(ctx.h, pk["ncons"], pk["ninst"], pk["nwit"], p(pk["a_query"]), p(pk["a_inf"]),
                                                p(pk["b_g1_query"]), p(pk["b1_inf"]), p(pk["b_g2_query"]), p(pk["b2_inf"]),
                                                p(pk["h_query"]), p(pk["h_inf"]), p(pk["l_query"]), p(pk["l_inf"]), p(pk["vk_g1"]),
                                                p(pk["vk_g2"]), C.byref(h)))
    return ProvingKey(ctx, h, 0)


def groth16_prove_r1cs(ctx: Context, scheme: int, pk: "ProvingKey", cs: R1cs, full_sh, r_sh, s_sh) -> dict:
    """create_proof + reveal for any circuit: full_sh = this party's shares of [instance, witness]."""
    full_sh = _np_u64(full_sh, 4)
    assert full_sh.shape[0] == cs.dims[1] + cs.dims[2]
    r_sh, s_sh = _np_u64(r_sh), _np_u64(s_sh)
    proof_sh, proof = np.zeros(48, np.uint64), np.zeros(48, np.uint64)
    sh_inf, inf = np.zeros(3, np.uint8), np.zeros(3, np.uint8)
    ctx._chk(ctx.lib.czk_groth16_prove_r1cs(ctx.h, scheme, pk.h, cs.h, full_sh.ctypes.data, r_sh.ctypes.data_as(u64p),
                                            s_sh.ctypes.data_as(u64p), proof_sh.ctypes.data_as(u64p), sh_inf.ctypes.data_as(u8p),
                                            proof.ctypes.data_as(u64p), inf.ctypes.data_as(u8p)))
    return dict(proof_sh=proof_sh, proof_sh_inf=sh_inf, proof=proof, proof_inf=inf)


# ------------------------------------------------------------------ wire format (ark-serialize canonical encodings; host-side)
def _ser_chk(rc):
    if rc:
        raise CzkError(rc, load_library().czk_last_error(None).decode())


def fr_serialize(a) -> bytes:
    a = _np_u64(a, 4)
    out = np.zeros(32 * a.shape[0], np.uint8)
    _ser_chk(load_library().czk_fr_serialize(a.ctypes.data, a.shape[0], out.ctypes.data))
    return out.tobytes()


def fr_deserialize(data: bytes) -> np.ndarray:
    buf = np.frombuffer(data, np.uint8).copy()
    n = buf.size // 32
    out = np.zeros((n, 4), np.uint64)
    _ser_chk(load_library().czk_fr_deserialize(buf.ctypes.data, n, out.ctypes.data))
    return out


def point_serialize(curve: int, xy, inf=None, compressed=True) -> bytes:
    w = 12 if curve == 1 else 24
    xy = _np_u64(xy, w)
    n = xy.shape[0]
    sz = (48 if curve == 1 else 96) * (1 if compressed else 2)
    out = np.zeros(sz * n, np.uint8)
    infp = None if inf is None else np.ascontiguousarray(inf, np.uint8)
    fn = load_library().czk_g1_serialize if curve == 1 else load_library().czk_g2_serialize
    _ser_chk(fn(xy.ctypes.data, None if infp is None else infp.ctypes.data, n, int(compressed), out.ctypes.data))
    return out.tobytes()


def point_deserialize(curve: int, data: bytes, compressed=True, check_subgroup=True):
    w = 12 if curve == 1 else 24
    sz = (48 if curve == 1 else 96) * (1 if compressed else 2)
    buf = np.frombuffer(data, np.uint8).copy()
    n = buf.size // sz
    xy, inf = np.zeros((n, w), np.uint64), np.zeros(n, np.uint8)
    fn = load_library().czk_g1_deserialize if curve == 1 else load_library().czk_g2_deserialize
    _ser_chk(fn(buf.ctypes.data, n, int(compressed), int(check_subgroup), xy.ctypes.data, inf.ctypes.data))
    return xy, inf


def groth16_proof_serialize(proof, proof_inf) -> bytes:
    proof, proof_inf = _np_u64(proof), np.ascontiguousarray(proof_inf, np.uint8)
    out = np.zeros(192, np.uint8)
    _ser_chk(load_library().czk_groth16_proof_serialize(proof.ctypes.data_as(u64p), proof_inf.ctypes.data_as(u8p), out.ctypes.data))
    return out.tobytes()


def groth16_proof_deserialize(data: bytes):
    buf = np.frombuffer(data, np.uint8).copy()
    proof, inf = np.zeros(48, np.uint64), np.zeros(3, np.uint8)
    _ser_chk(load_library().czk_groth16_proof_deserialize(buf.ctypes.data, proof.ctypes.data_as(u64p), inf.ctypes.data_as(u8p)))
    return proof, inf


def pairing_product_is_one(g1_xy, g2_xy, g1_inf=None, g2_inf=None) -> bool:
    """prod e(P_i, Q_i) == 1 (host-side BLS12-377 pairing of libczk_b200)."""
    g1_xy, g2_xy = _np_u64(g1_xy, 12), _np_u64(g2_xy, 24)
    n = g1_xy.shape[0]
    assert g2_xy.shape[0] == n
    i1 = None if g1_inf is None else np.ascontiguousarray(g1_inf, np.uint8)
    i2 = None if g2_inf is None else np.ascontiguousarray(g2_inf, np.uint8)
    res = C.c_int()
    _ser_chk(load_library().czk_pairing_product_is_one(g1_xy.ctypes.data, None if i1 is None else i1.ctypes.data, g2_xy.ctypes.data,
                                                       None if i2 is None else i2.ctypes.data, n, C.byref(res)))
    return bool(res.value)


def groth16_verify(pk: dict, public_inputs, proof, proof_inf) -> bool:
    """verify_proof on a key dict (vk_g1 = alpha | beta | delta, vk_g2 = beta | gamma | delta, gamma_abc_g1)."""
    alpha = _np_u64(pk["vk_g1"]).reshape(-1)[:12].copy()
    vk_g2 = _np_u64(pk["vk_g2"]).reshape(-1).copy()
    abc = _np_u64(pk["gamma_abc_g1"], 12)
    pub = _np_u64(public_inputs, 4) if len(public_inputs) else np.zeros((0, 4), np.uint64)
    assert pub.shape[0] == abc.shape[0] - 1
    proof, proof_inf = _np_u64(proof), np.ascontiguousarray(proof_inf, np.uint8)
    ok = C.c_int()
    _ser_chk(load_library().czk_groth16_verify(alpha.ctypes.data_as(u64p), vk_g2.ctypes.data_as(u64p), abc.ctypes.data, abc.shape[0],
                                               pub.ctypes.data, proof.ctypes.data_as(u64p), proof_inf.ctypes.data_as(u8p), C.byref(ok)))
    return bool(ok.value)


def fixed_base_msm(ctx: Context, curve: int, base_xy, scalars: DeviceVec, n=None, sc_off=0) -> "Bases":
    """scalars[i] * base as a resident base set (FixedBaseMSM::multi_scalar_mul on the device)."""
    base_xy = _np_u64(base_xy)
    n = scalars.n - sc_off if n is None else n
    h = C.c_void_p()
    ctx._chk(ctx.lib.czk_fixed_base_msm(ctx.h, curve, base_xy.ctypes.data, scalars.h, sc_off, n, C.byref(h)))
    return Bases(ctx, h, curve)


def groth16_setup(ctx: Context, n_sq: int, toxic_mont) -> "ProvingKey":
    """generate_parameters for the squaring circuit on the device (toxic: 7 Montgomery Fr)."""
    toxic = _np_u64(toxic_mont, 4)
    assert toxic.shape[0] == 7
    h = C.c_void_p()
    ctx._chk(ctx.lib.czk_groth16_setup(ctx.h, n_sq, toxic.ctypes.data, C.byref(h)))
    pk = ProvingKey(ctx, h, n_sq)
    pk.ninst = 2
    return pk


def groth16_setup_r1cs(ctx: Context, cs: dict, toxic_mont) -> "ProvingKey":
    toxic = _np_u64(toxic_mont, 4)
    mats = [tuple(np.ascontiguousarray(x) for x in cs[m]) for m in ("a", "b", "c")]
    rp = (C.c_void_p * 3)(*[m[0].ctypes.data for m in mats])
    col = (C.c_void_p * 3)(*[m[1].ctypes.data for m in mats])
    cf = (C.c_void_p * 3)(*[m[2].ctypes.data for m in mats])
    h = C.c_void_p()
    ctx._chk(ctx.lib.czk_groth16_setup_r1cs(ctx.h, cs["ncons"], cs["ninst"], cs["nwit"], rp, col, cf, toxic.ctypes.data, C.byref(h)))
    pk = ProvingKey(ctx, h, 0)
    pk.ninst = cs["ninst"]
    return pk


def pk_verifying_key(pk: "ProvingKey") -> dict:
    """The verifier's view of a key generated on the device: vk_g1 (alpha | beta | delta), vk_g2 (beta | gamma | delta),
    gamma_abc_g1 - the dict czk_b200.groth16_verify takes."""
    vk1, vk2 = np.zeros(36, np.uint64), np.zeros(72, np.uint64)
    pk.ctx._chk(pk.ctx.lib.czk_groth16_pk_vk(pk.h, vk1.ctypes.data_as(u64p), vk2.ctypes.data_as(u64p)))
    abc = np.zeros((pk.ninst, 12), np.uint64)
    pk.ctx._chk(pk.ctx.lib.czk_groth16_pk_gamma_abc(pk.h, abc.ctypes.data, pk.ninst))
    return dict(vk_g1=vk1.reshape(3, 12), vk_g2=vk2.reshape(3, 24), gamma_abc_g1=abc)

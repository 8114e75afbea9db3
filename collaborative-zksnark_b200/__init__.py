"""czk_b200 - B200-native hot path of collaborative-zksnark (MSM, NTT, share opening / Beaver
multiplication, Groth16 prover loop) behind a C ABI (include/czk.h).

This package is the Python host-side mirror used by the tests and the benchmark; the product is
libczk_b200.so.  There is no CPU fallback: creating a Context without a CUDA device raises.
"""
from .binding import (  # noqa: F401
    CzkError,
    Context,
    DeviceVec,
    Bases,
    load_library,
    library_path,
    domain_params,
    mixed_domain_params,
    SCHEME_PLAIN,
    SCHEME_ADDITIVE,
    SCHEME_SPDZ,
    SCHEME_GSZ,
    NTT_FFT,
    NTT_IFFT,
    NTT_COSET_FFT,
    NTT_COSET_IFFT,
    NTT_IFFT_COSET_FFT,
    ProvingKey,
    groth16_witness_map,
    groth16_prove,
    squaring_chain,
    king_share_batch,
    R1cs,
    groth16_pk_upload_r1cs,
    groth16_prove_r1cs,
    fr_serialize,
    fr_deserialize,
    point_serialize,
    point_deserialize,
    groth16_proof_serialize,
    groth16_proof_deserialize,
    pairing_product_is_one,
    groth16_verify,
    fixed_base_msm,
    groth16_setup,
    groth16_setup_r1cs,
    pk_verifying_key,
    plonk_prove_wiring,
    PlonkTranscript,
    PlonkWiringProof,
)
from .build import build as build_library  # noqa: F401

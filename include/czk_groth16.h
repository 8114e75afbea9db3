/* czk_groth16.h - the Groth16 prover loop of collaborative-zksnark on top of czk.h.
 *
 * Host-side orchestration in C++ (the reference's is Rust: mpc-snarks/src/groth/prover.rs:66-177,
 * mpc-snarks/src/groth/r1cs_to_qap.rs:47-112), specialised like the reference's benchmark to the
 * repeated-squaring circuit (mpc-snarks/src/proof.rs:304-344).  One call = one party's share of the
 * proof plus the revealed proof (`create_random_proof` + `pf.reveal()`, proof.rs:130-139): every NTT,
 * MSM, share product and opening runs on this party's GPU; the parties meet only in the NCCL
 * collectives of czk_net_*.
 */
#ifndef CZK_GROTH16_H
#define CZK_GROTH16_H
#include "czk.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct czk_pk czk_pk; /* device-resident ProvingKey (groth16 ProvingKey<E>: a/b_g1/b_g2/h/l queries + vk points) */

/* Upload a proving key for n_sq squarings.  Shapes (groth16/src/generator.rs:109-221): a_query, b_g1_query,
 * b_g2_query: n_sq + 2 points; h_query: D - 1 points with D = next_pow2(n_sq + 2); l_query: n_sq points;
 * vk_g1 = alpha_g1 | beta_g1 | delta_g1 ; vk_g2 = beta_g2 | gamma_g2 | delta_g2.  inf arrays may be NULL. */
CZK_API int czk_groth16_pk_upload(czk_ctx* ctx, size_t n_sq, const uint64_t* a_query, const uint8_t* a_inf,
                                  const uint64_t* b_g1_query, const uint8_t* b1_inf, const uint64_t* b_g2_query,
                                  const uint8_t* b2_inf, const uint64_t* h_query, const uint8_t* h_inf,
                                  const uint64_t* l_query, const uint8_t* l_inf, const uint64_t vk_g1[36],
                                  const uint64_t vk_g2[72], czk_pk** out);
/* A proving key of the same shapes filled with synthetic device-generated bases (benchmarks: the proof does not
 * verify, the work is identical).  Same seed => same key on every rank. */
CZK_API int czk_groth16_pk_synthetic(czk_ctx* ctx, size_t n_sq, uint64_t seed, czk_pk** out);
CZK_API void czk_groth16_pk_free(czk_ctx* ctx, czk_pk* pk);
CZK_API size_t czk_groth16_pk_domain_size(const czk_pk* pk);
/* which: 0 a_query, 1 b_g1_query, 2 b_g2_query, 3 h_query, 4 l_query */
CZK_API const czk_bases* czk_groth16_pk_query(const czk_pk* pk, int which);
CZK_API int czk_groth16_pk_vk(const czk_pk* pk, uint64_t vk_g1[36], uint64_t vk_g2[72]);

/* R1CStoQAP::witness_map on this party's shares: chain_sh = n_sq + 1 Montgomery Fr (shares of w_0..w_{n-1}, out)
 * in host memory; h_out = D Fr (this party's share of the quotient coefficients), host memory. */
CZK_API int czk_groth16_witness_map(czk_ctx* ctx, int scheme, size_t n_sq, const uint64_t* chain_sh, uint64_t* h_out);

/* create_proof on shares followed by reveal.  r_sh / s_sh: this party's shares of the prover randomness.
 * proof_sh: this party's share A | B | C as affine x|y (12 + 24 + 12 limbs), proof_sh_inf: 3 infinity bytes;
 * proof / proof_inf: the revealed proof (identical on every party).  scheme PLAIN requires a 1-party context. */
CZK_API int czk_groth16_prove(czk_ctx* ctx, int scheme, const czk_pk* pk, const uint64_t* chain_sh, const uint64_t r_sh[4],
                              const uint64_t s_sh[4], uint64_t proof_sh[48], uint8_t proof_sh_inf[3], uint64_t proof[48],
                              uint8_t proof_inf[3]);
/* scheme CZK_SCHEME_GSZ: chain_sh / r_sh / s_sh are the values every party holds (king_share_batch hands the plaintext to
 * every party at degree t, gsz20/mod.rs:202-212); the first group reveal runs the queued field and group product checks.
 * After such a proof: the values those checks opened (field x | y | z; group x, then y | z affine + 2 infinity bytes)
 * and counts = { king computations, opens } of the context. */
CZK_API int czk_groth16_gsz_last_checks(const czk_ctx* ctx, uint64_t field_xyz[12], uint64_t group_x[4], uint64_t group_yz[24],
                                        uint8_t group_inf[2], uint64_t counts[2]);
/* Same, with this party's chain shares already resident on the device (n_sq + 1 elements). */
CZK_API int czk_groth16_prove_vec(czk_ctx* ctx, int scheme, const czk_pk* pk, const czk_vec* chain_dev, const uint64_t r_sh[4],
                                  const uint64_t s_sh[4], uint64_t proof_sh[48], uint8_t proof_sh_inf[3], uint64_t proof[48],
                                  uint8_t proof_inf[3]);
/* ---- any circuit -----------------------------------------------------------------------------------------------------
 * The entry points above are specialised to the benchmark's squaring circuit (like the reference's `proof` binary); these
 * take the constraint matrices the reference's prover reads from its ConstraintSystem (`prover.to_matrices()`,
 * mpc-snarks/src/groth/r1cs_to_qap.rs:47-83): for m in {A, B, C}, row i of matrix m is entries row_ptr[m][i] ..
 * row_ptr[m][i+1] of (col[m], coeff[m]) with coeff in Montgomery form.  Variables: ninst instance variables (the first is the
 * constant one) followed by nwit witness variables.  The sparse evaluate_constraint pass (r1cs_to_qap.rs:12-41) runs on
 * the device, linear in the shares. */
typedef struct czk_r1cs czk_r1cs;
CZK_API int czk_r1cs_upload(czk_ctx* ctx, size_t ncons, size_t ninst, size_t nwit, const uint64_t* const row_ptr[3],
                            const uint32_t* const col[3], const uint64_t* const coeff[3], czk_r1cs** out);
CZK_API void czk_r1cs_free(czk_ctx* ctx, czk_r1cs* r);
/* Proving key of a general circuit (groth16/src/generator.rs:109-221): a_query, b_g1_query, b_g2_query: ninst + nwit points;
 * h_query: D - 1 points with D = next_pow2(ncons + ninst); l_query: nwit points. */
CZK_API int czk_groth16_pk_upload_r1cs(czk_ctx* ctx, size_t ncons, size_t ninst, size_t nwit, const uint64_t* a_query,
                                       const uint8_t* a_inf, const uint64_t* b_g1_query, const uint8_t* b1_inf,
                                       const uint64_t* b_g2_query, const uint8_t* b2_inf, const uint64_t* h_query, const uint8_t* h_inf,
                                       const uint64_t* l_query, const uint8_t* l_inf, const uint64_t vk_g1[36],
                                       const uint64_t vk_g2[72], czk_pk** out);
/* create_proof + reveal on this party's shares of the full assignment [instance, witness] (ninst + nwit Montgomery Fr, host
 * memory).  Public values are passed in their lowered share form, as the reference's MpcField does when a Public meets a
 * Shared: additive / SPDZ - the king holds the value, the others 0 (add.rs:88-92); GSZ - every party holds it. */
CZK_API int czk_groth16_prove_r1cs(czk_ctx* ctx, int scheme, const czk_pk* pk, const czk_r1cs* cs, const uint64_t* full_sh,
                                   const uint64_t r_sh[4], const uint64_t s_sh[4], uint64_t proof_sh[48], uint8_t proof_sh_inf[3],
                                   uint64_t proof[48], uint8_t proof_inf[3]);

/* Proof::serialize / deserialize (groth16/src/data_structures.rs: a: G1Affine | b: G2Affine | c: G1Affine, compressed:
 * 48 + 96 + 48 bytes) on the `proof` / `proof_inf` arrays the prover returns. */
CZK_API int czk_groth16_proof_serialize(const uint64_t proof[48], const uint8_t proof_inf[3], uint8_t out[192]);
CZK_API int czk_groth16_proof_deserialize(const uint8_t in[192], uint64_t proof[48], uint8_t proof_inf[3]);

/* ---- CRS generation: groth16/src/generator.rs:34-221 with caller-supplied toxic waste ----------------
 * toxic = alpha | beta | gamma | delta | tau | g1_scalar | g2_scalar (7 Montgomery Fr; the generators are g1_scalar * G1 and
 * g2_scalar * G2 of the curve's standard generators).  The five queries are computed by fixed-base multi-scalar
 * multiplications on the device (replacing FixedBaseMSM, algebra/ec/src/msm/fixed_base.rs:12-96) and stay resident in the
 * returned key, which also carries gamma_abc_g1 for the verifier.  czk_groth16_setup = the benchmark's squaring circuit. */
CZK_API int czk_groth16_setup(czk_ctx* ctx, size_t n_sq, const uint64_t toxic[28], czk_pk** out);
CZK_API int czk_groth16_setup_r1cs(czk_ctx* ctx, size_t ncons, size_t ninst, size_t nwit, const uint64_t* const row_ptr[3],
                                   const uint32_t* const col[3], const uint64_t* const coeff[3], const uint64_t toxic[28],
                                   czk_pk** out);
/* gamma_abc_g1 (ninst points) of a key generated by the two calls above. */
CZK_API int czk_groth16_pk_gamma_abc(const czk_pk* pk, uint64_t* out, size_t ninst);
/* out[i] = scalars[sc_off + i] * base for i < n as a resident base set, infinity where the scalar is zero
 * (FixedBaseMSM::multi_scalar_mul + batch normalisation).  curve: 1 = G1 (base 12 limbs), 2 = G2 (24 limbs). */
CZK_API int czk_fixed_base_msm(czk_ctx* ctx, int curve, const uint64_t* base_xy, const czk_vec* scalars, size_t sc_off, size_t n,
                               czk_bases** out);

/* ---- verifier, host-side, ctx-free ------------------------------------------------------------------
 * The acceptance test the reference runs after every benchmark proof (`verify_proof`, mpc-snarks/src/proof.rs:141,
 * groth16/src/verifier.rs; pairing: algebra/ec/src/models/bls12/mod.rs:59-200).
 * czk_pairing_product_is_one: result = 1 iff prod_i e(P_i, Q_i) == 1 (P_i in G1: n x 12 limbs, Q_i in G2: n x 24 limbs,
 * Montgomery affine; infinity flags may be NULL).
 * czk_groth16_verify: ok = 1 iff e(A, B) == e(alpha, beta) e(gamma_abc[0] + sum_i x_i gamma_abc[i], gamma) e(C, delta).
 * vk_g2 = beta_g2 | gamma_g2 | delta_g2 (as czk_groth16_pk_vk returns); gamma_abc_g1: ninst points; public_inputs:
 * ninst - 1 Montgomery Fr (the constant one is implicit). */
CZK_API int czk_pairing_product_is_one(const uint64_t* g1_xy, const uint8_t* g1_inf, const uint64_t* g2_xy, const uint8_t* g2_inf,
                                       size_t n, int* result);
CZK_API int czk_groth16_verify(const uint64_t alpha_g1[12], const uint64_t vk_g2[72], const uint64_t* gamma_abc_g1, size_t ninst,
                               const uint64_t* public_inputs, const uint64_t proof[48], const uint8_t proof_inf[3], int* ok);

/* Witness generation of the benchmark circuit (mpc-snarks/src/proof.rs:308-310): out[i] = start^(2^i), i <= n_sq.
 * Host-side, serial by nature, outside the reference's timed section. */
CZK_API int czk_squaring_chain(const uint64_t start[4], size_t n_sq, uint64_t* out);
/* Reveal::king_share_batch (share/add.rs:105-117, spdz.rs:150-162): additive shares of k values for n parties,
 * parties 0..n-2 uniform, party n-1 the remainder; out = n_parties * k Fr, party-major.  Host-side. */
CZK_API int czk_king_share_batch(const uint64_t* values, size_t k, int n_parties, uint64_t seed, uint64_t* out);
/* Per-phase device/host milliseconds of the last czk_groth16_prove on this context:
 * 0 upload, 1 witness map (NTTs + product), 2 MSM h, 3 MSM l, 4 MSM a, 5 MSM b_g1, 6 MSM b_g2, 7 group tail + reveal. */
CZK_API int czk_groth16_last_phases(const czk_ctx* ctx, double out_ms[8]);

#ifdef __cplusplus
}
#endif
#endif

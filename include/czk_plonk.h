/* czk_plonk.h - the data path of the collaborative Plonk prover's wiring argument on top of czk.h.
 *
 * Replaces the bodies of mpc-plonk/src/lib.rs:199-258 (prove_wiring), :110-197 (prove_unit_product) and :343-400
 * (eval, commit) when the field is MpcField over additive / SPDZ shares and the commitment scheme is KZG10 without
 * hiding (poly-commit/src/kzg10/mod.rs:141-262): every transform, share product / division / prefix product, commitment
 * MSM and opening MSM runs on this party's GPU, the parties meet in the NCCL exchanges of czk_net_*.
 *
 * The Fiat-Shamir transcript stays with the caller.  The reference derives its challenges from
 * FiatShamirRng<Blake2s> (mpc-plonk/src/util.rs:47-118); the prover here calls back into `czk_plonk_transcript` at the
 * same points, in the same order - absorb after every publicized commitment (lib.rs:393-396), challenge wherever the
 * reference calls fs_rng.gen() - so a Rust caller plugs in its own FiatShamirRng and gets the reference's challenges.
 * czk_plonk_standin_transcript is a documented stand-in (SplitMix64 over the absorbed limbs) for tests, the benchmark and
 * the CLI stand-in; it is NOT the reference's transcript.
 */
#ifndef CZK_PLONK_H
#define CZK_PLONK_H
#include "czk.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct czk_plonk_transcript {
    void* user;
    /* a publicized commitment: affine x | y (12 Montgomery limbs) and its infinity flag (ark to_bytes![commitment]) */
    void (*absorb_g1)(void* user, const uint64_t xy[12], int inf);
    /* fs_rng.gen::<F>(): a field element in Montgomery form */
    void (*challenge)(void* user, uint64_t out[4]);
} czk_plonk_transcript;

/* The stand-in transcript: state = one 64-bit word.  czk_plonk_standin_transcript(&state, seed, &t) fills `t`. */
CZK_API void czk_plonk_standin_transcript(uint64_t* state, uint64_t seed, czk_plonk_transcript* out);

/* WiringProof (mpc-plonk/src/data_structures.rs) in a flat layout.  Commitments: l1, t, q, l2_q.  Openings, in the order
 * the reference calls eval():  0 t(w r)  1 t(r)  2 t(w^(k-1))  3 l1(w r)  4 q(r)  5 l2_q(x)  6 w(x)  7 l1(x)  8 p(x);
 * each is the publicized value and the KZG10 opening proof (affine G1).  challenges: y, z, r, x as drawn. */
typedef struct czk_plonk_wiring_proof {
    uint64_t cmt_xy[4][12];
    uint8_t cmt_inf[4];
    uint64_t open_val[9][4];
    uint64_t open_pf_xy[9][12];
    uint8_t open_pf_inf[9];
    uint64_t challenges[4][4];
} czk_plonk_wiring_proof;

/* prove_wiring for the wire polynomial p (this party's shares of its 2^log_d coefficients; p_mac: the SPDZ MAC
 * component, NULL otherwise) against the public wiring polynomial w_pub, over the domain of size 2^log_d.
 * powers: the KZG10 committer key powers_of_g (>= 2^log_d points, resident; precompute its table for speed).
 * out_share: this party's shares of the nine opening proofs (values and commitments are already public);
 * out: the revealed proof, identical on every party.  phases_ms (optional, 4 doubles): transforms + share protocols,
 * commitments, openings, reveal.  Returns CZK_ERR_PROTOCOL on a failed MAC check or a zero denominator. */
CZK_API int czk_plonk_prove_wiring(czk_ctx* ctx, int scheme, const czk_bases* powers, unsigned log_d, const czk_vec* p_sh,
                                   const czk_vec* p_mac, const czk_vec* w_pub, const czk_plonk_transcript* transcript,
                                   czk_plonk_wiring_proof* out_share, czk_plonk_wiring_proof* out, double* phases_ms);
/* The same argument over the reference's own wire domain, MixedRadixEvaluationDomain::new(3 * n_gates) = 3 * 2^log_m points
 * (mpc-plonk/src/relations/flat.rs:282-300): every vector and the committer key hold >= 3 * 2^log_m elements. */
CZK_API int czk_plonk_prove_wiring_mixed(czk_ctx* ctx, int scheme, const czk_bases* powers, unsigned log_m, const czk_vec* p_sh,
                                         const czk_vec* p_mac, const czk_vec* w_pub, const czk_plonk_transcript* transcript,
                                         czk_plonk_wiring_proof* out_share, czk_plonk_wiring_proof* out, double* phases_ms);

#ifdef __cplusplus
}
#endif
#endif /* CZK_PLONK_H */

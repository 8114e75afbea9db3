/* czk.h - C ABI of libczk_b200.so: the B200 (sm_100a) hot path of collaborative-zksnark.
 *
 * Every entry point is what the reference's FFI for this path would bind (the Rust
 * `-sys` stub is in INTEGRATION.md).  Plain pointers and sizes only.  All field
 * elements cross the boundary in the reference's in-memory form: little-endian u64
 * limbs in Montgomery representation (Fr = 4 limbs, R = 2^256; Fq = 6 limbs,
 * R = 2^384; Fq2 = c0 | c1), see algebra/ff/src/fields/macros.rs:103-108.
 * Affine points are x | y (12 limbs G1, 24 limbs G2) plus a separate infinity byte
 * (the reference's GroupAffine is repr(Rust): x, y, infinity: bool -
 * algebra/ec/src/models/short_weierstrass_jacobian.rs:43-49).
 *
 * Error behaviour: the reference's functions on this path return values and panic on
 * misuse; here every function returns 0 on success and a non-zero czk_status
 * otherwise, czk_last_error() gives the message, and the binding panics on non-zero.
 * Calls are synchronous (they return when the result is in the caller's buffer),
 * matching the reference's blocking calls from a single main thread per party.
 */
#ifndef CZK_H
#define CZK_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define CZK_API __attribute__((visibility("default")))
#else
#define CZK_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef enum czk_status {
    CZK_OK = 0,
    CZK_ERR_CUDA = 1,       /* a CUDA runtime call or kernel failed */
    CZK_ERR_ARG = 2,        /* invalid argument (size, null pointer, unsupported domain) */
    CZK_ERR_NO_DEVICE = 3,  /* no usable CUDA device: the product path never falls back to the CPU */
    CZK_ERR_NCCL = 4,
    CZK_ERR_PROTOCOL = 5    /* MAC check / degree check / cross-party consistency failed (reference: assert!) */
} czk_status;

typedef struct czk_ctx czk_ctx;     /* one per party / GPU */
typedef struct czk_vec czk_vec;     /* device-resident vector of Fr */
typedef struct czk_bases czk_bases; /* device-resident affine bases (a CRS query) */

/* ---- context ---------------------------------------------------------------------------- */
CZK_API int czk_ctx_create(int device, czk_ctx** out);
CZK_API void czk_ctx_destroy(czk_ctx* ctx);
CZK_API const char* czk_last_error(const czk_ctx* ctx); /* ctx may be NULL: last error of the calling thread */
CZK_API int czk_ctx_sync(czk_ctx* ctx);
/* The CUDA stream every kernel of this context is launched on (a cudaStream_t). */
CZK_API void* czk_ctx_stream(czk_ctx* ctx);
/* Number of this library's kernel launches issued through ctx since creation. */
CZK_API uint64_t czk_ctx_launches(const czk_ctx* ctx);
CZK_API const char* czk_version(void);

/* ---- device vectors of Fr ----------------------------------------------------------------- */
CZK_API int czk_vec_alloc(czk_ctx* ctx, size_t n, czk_vec** out); /* zero filled */
CZK_API void czk_vec_free(czk_ctx* ctx, czk_vec* v);
CZK_API size_t czk_vec_len(const czk_vec* v);
CZK_API uint64_t* czk_vec_device_ptr(czk_vec* v);
CZK_API int czk_vec_upload(czk_ctx* ctx, czk_vec* v, size_t offset, const uint64_t* host, size_t n);
CZK_API int czk_vec_download(czk_ctx* ctx, const czk_vec* v, size_t offset, uint64_t* host, size_t n);
CZK_API int czk_vec_copy(czk_ctx* ctx, czk_vec* dst, size_t dst_off, const czk_vec* src, size_t src_off, size_t n);
CZK_API int czk_vec_zero(czk_ctx* ctx, czk_vec* v, size_t offset, size_t n);

/* ---- NTT: replaces Radix2EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place ------------
 * algebra/poly/src/domain/radix2/mod.rs:99-117, radix2/fft.rs:22-260, domain/mod.rs:93-142.
 * In place, natural order in and out, size 2^log_d (the caller zero-pads like mod.rs:100-101).
 * inverse: 0 fft, 1 ifft.  coset: 0 subgroup, 1 coset of the multiplicative generator 22.      */
CZK_API int czk_ntt_fr(czk_ctx* ctx, uint64_t* host_data, unsigned log_d, int inverse, int coset); /* host buffer in/out */
CZK_API int czk_ntt_fr_dev(czk_ctx* ctx, uint64_t* dev_data, unsigned log_d, int inverse, int coset);
CZK_API int czk_ntt_vec(czk_ctx* ctx, czk_vec* v, unsigned log_d, int inverse, int coset);
/* Batched form: the same transform over `count` device vectors (distinct, each 2^log_d elements), all of them in ONE grid
 * per pass - what the generic `T: DomainCoeff<F>` call amounts to for SPDZ (value and MAC vectors, spdz.rs:186-208) and
 * for the a / b / c vectors of the witness map (r1cs_to_qap.rs:85-110).
 * op: one of the four transforms, or CZK_NTT_IFFT_COSET_FFT = ifft_in_place followed by coset_fft_in_place on the same
 * vector (r1cs_to_qap.rs:85-90): same result as the two calls, computed as an inverse decimation-in-frequency pass
 * sequence feeding a forward decimation-in-time one, with the D^-1 g^i scaling fused in and no reordering pass. */
#define CZK_NTT_FFT 0
#define CZK_NTT_IFFT 1
#define CZK_NTT_COSET_FFT 2
#define CZK_NTT_COSET_IFFT 3
#define CZK_NTT_IFFT_COSET_FFT 4
CZK_API int czk_ntt_fr_batch(czk_ctx* ctx, uint64_t* const* dev_vecs, int count, unsigned log_d, int op);
CZK_API int czk_ntt_vec_batch(czk_ctx* ctx, czk_vec* const* vecs, int count, unsigned log_d, int op);
/* Domain constants as Radix2EvaluationDomain::new computes them (radix2/mod.rs:51-82). */
CZK_API int czk_domain_params(unsigned log_d, uint64_t group_gen[4], uint64_t group_gen_inv[4], uint64_t size_inv[4],
                      uint64_t generator_inv[4]);

/* ---- mixed-radix domains: 3 * 2^log_m points ------------------------------------------------------
 * Replaces MixedRadixEvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place (algebra/poly/src/domain/mixed_radix.rs:
 * 130-157, domain/mod.rs:139-142) for Fr's small subgroup base 3 - the wire domain of the Plonk prover
 * (mpc-plonk/src/relations/flat.rs:282-300).  Natural order in and out, out[j] = sum_i in[i] w^(ij) with
 * w = get_root_of_unity(3 * 2^log_m); every vector holds 3 * 2^log_m elements.  op as above (CZK_NTT_IFFT_COSET_FFT runs as
 * the two transforms in sequence). */
/* host vector of 3 * 2^log_m elements, in place (fft_in_place::<Fr> on plain field elements) */
CZK_API int czk_ntt_mixed_fr(czk_ctx* ctx, uint64_t* host_data, unsigned log_m, int inverse, int coset);
CZK_API int czk_ntt_mixed_fr_batch(czk_ctx* ctx, uint64_t* const* dev_vecs, int count, unsigned log_m, int op);
CZK_API int czk_ntt_mixed_vec_batch(czk_ctx* ctx, czk_vec* const* vecs, int count, unsigned log_m, int op);
/* Domain constants as MixedRadixEvaluationDomain::new computes them (mixed_radix.rs:64-105). */
CZK_API int czk_mixed_domain_params(unsigned log_m, uint64_t group_gen[4], uint64_t group_gen_inv[4], uint64_t size_inv[4],
                            uint64_t generator_inv[4]);

/* ---- pointwise Fr helpers used between the transforms -------------------------------------------
 * domain/mod.rs:93-126 (distribute_powers), :184-191 (divide_by_vanishing_poly_on_coset_in_place),
 * mpc-snarks/src/groth/r1cs_to_qap.rs:92,105-109.                                                */
CZK_API int czk_vec_add(czk_ctx* ctx, czk_vec* a, const czk_vec* b, size_t n);              /* a += b */
CZK_API int czk_vec_sub(czk_ctx* ctx, czk_vec* a, const czk_vec* b, size_t n);              /* a -= b */
CZK_API int czk_vec_mul(czk_ctx* ctx, czk_vec* a, const czk_vec* b, size_t n);              /* a *= b (plain field) */
CZK_API int czk_vec_scale(czk_ctx* ctx, czk_vec* a, const uint64_t c[4], size_t n);         /* a *= c */
CZK_API int czk_vec_distribute_powers(czk_ctx* ctx, czk_vec* a, const uint64_t g[4], const uint64_t c[4], size_t n); /* a[i] *= c g^i */
CZK_API int czk_vec_divide_by_vanishing_on_coset(czk_ctx* ctx, czk_vec* a, unsigned log_d);

/* ---- MSM: replaces VariableBaseMSM::multi_scalar_mul / AffineCurve::multi_scalar_mul / Msm::msm --
 * algebra/ec/src/msm/variable_base.rs:12-106, algebra/ec/src/lib.rs:302-311,
 * mpc-algebra/src/share/msm.rs:6-48.  Uses min(len) terms; infinity bases and zero scalars are
 * skipped.  scalars_montgomery = 1: Fr in Montgomery form (what AffineCurve::multi_scalar_mul
 * takes); 0: canonical BigInt256 (what VariableBaseMSM takes).  The result is written as a Jacobian
 * triple x | y | z that is already affine-normalised: (x, y, 1), or (1, 1, 0) for infinity
 * (short_weierstrass_jacobian.rs:440-448), i.e. `.into_affine()` is a no-op on it.               */
CZK_API int czk_msm_g1(czk_ctx* ctx, const uint64_t* bases_xy, const uint8_t* inf, const uint64_t* scalars,
               int scalars_montgomery, size_t n, uint64_t out_xyz[18]);
CZK_API int czk_msm_g2(czk_ctx* ctx, const uint64_t* bases_xy, const uint8_t* inf, const uint64_t* scalars,
               int scalars_montgomery, size_t n, uint64_t out_xyz[36]);
/* Device-resident bases (the reference re-passes the same pk.*_query every proof). curve: 1 = G1, 2 = G2. */
CZK_API int czk_bases_upload(czk_ctx* ctx, int curve, const uint64_t* bases_xy, const uint8_t* inf, size_t n, czk_bases** out);
CZK_API void czk_bases_free(czk_ctx* ctx, czk_bases* b);
CZK_API size_t czk_bases_len(const czk_bases* b);
/* Precompute the merged-window table 2^(c w) * P_i for a resident base set (c = 0: chosen from its length).  One-off
 * cost per CRS query; afterwards czk_msm_bases uses one bucket set for all windows.  Same results. */
CZK_API int czk_bases_precompute(czk_ctx* ctx, czk_bases* b, unsigned c);
/* Bytes of device memory held by the base set: { points + flags, merged-window table } (the per-key cost of the table). */
CZK_API int czk_bases_device_bytes(const czk_bases* b, uint64_t out[2]);
/* MSM of bases[base_off .. base_off+n) by the device scalars sc[sc_off .. sc_off+n). */
CZK_API int czk_msm_bases(czk_ctx* ctx, const czk_bases* b, size_t base_off, const czk_vec* sc, size_t sc_off,
                  int scalars_montgomery, size_t n, uint64_t* out_xyz);
/* Several MSMs over ONE scalar vector: out_xyz[k] = sum_i sc[sc_off + i] * b[k][base_off + i], k < count.  Groth16's a-,
 * b_g1- and b_g2-query MSMs all take the full assignment (groth16/src/prover.rs:80-89, mpc-snarks/src/groth/prover.rs:
 * 104-160): the digit decomposition and the bucket sort are done once for every base set that addresses a precomputed
 * table of the first set's shape and carries its infinity flags (checked on the device); any other set runs as an MSM
 * of its own.  Same results as `count` calls of czk_msm_bases.  ms_out (optional): wall ms per set. */
CZK_API int czk_msm_bases_multi(czk_ctx* ctx, const czk_bases* const* b, int count, size_t base_off, const czk_vec* sc, size_t sc_off,
                        int scalars_montgomery, size_t n, uint64_t* const* out_xyz, double* ms_out);
/* Synthetic bases for benchmarks: P_i = (k0 + i*k1 + i^2*k2) * G with seeded 248-bit k0, k1, k2 (distinct points of
 * the prime-order subgroup, generated on the device), every `inf_every`-th entry flagged infinity (0 = none). */
CZK_API int czk_bases_synthetic(czk_ctx* ctx, int curve, uint64_t seed, size_t n, size_t inf_every, czk_bases** out);
CZK_API int czk_bases_download(czk_ctx* ctx, const czk_bases* b, size_t off, size_t n, uint64_t* xy, uint8_t* inf);

/* ---- network: replaces mpc-net's MpcNet (mpc-net/src/lib.rs:28-70, multi.rs:145-242) ------------
 * One rank = one party = one GPU; rank 0 is the king.  nccl_unique_id is the 128-byte ncclUniqueId
 * produced by czk_net_unique_id on rank 0 and distributed by the launcher.                        */
CZK_API int czk_net_unique_id(uint8_t out[128]);
CZK_API int czk_net_init(czk_ctx* ctx, int rank, int nranks, const uint8_t nccl_unique_id[128]);
CZK_API int czk_net_init_single(czk_ctx* ctx); /* 1 party, no communicator */
CZK_API void czk_net_deinit(czk_ctx* ctx);
CZK_API int czk_net_party_id(const czk_ctx* ctx);
CZK_API int czk_net_n_parties(const czk_ctx* ctx);
/* broadcast_bytes: every party contributes `bytes` bytes, receives all (nranks * bytes) in rank order. */
CZK_API int czk_net_allgather_dev(czk_ctx* ctx, const void* dev_send, void* dev_recv, size_t bytes);
CZK_API int czk_net_allgather_host(czk_ctx* ctx, const void* host_send, void* host_recv, size_t bytes);
/* send_bytes_to_king / recv_bytes_from_king with equal slices for every party. */
CZK_API int czk_net_bcast_from_king_dev(czk_ctx* ctx, void* dev_buf, size_t bytes);
/* send_bytes_to_king (mpc-net/src/multi.rs:176-209): rank 0 receives every party's `bytes` bytes at
 * dev_recv_king[p * bytes] (its own slice is copied); dev_recv_king is ignored on the other ranks. */
CZK_API int czk_net_gather_to_king_dev(czk_ctx* ctx, const void* dev_send, void* dev_recv_king, size_t bytes);
/* Stats (mpc-net/src/lib.rs:8-26): bytes_sent, bytes_recv, broadcasts, to_king, from_king. */
CZK_API int czk_net_stats(const czk_ctx* ctx, uint64_t out[5]);
CZK_API void czk_net_reset_stats(czk_ctx* ctx);

/* ---- shares: replaces FieldShare::{batch_open,batch_mul} ------------------------------------------
 * mpc-algebra/src/share/{field.rs:97-127, add.rs:121-125, spdz.rs:166-185}.
 * scheme: 1 = additive (hbc), 2 = SPDZ, 3 = GSZ (see below).  For SPDZ a share vector is two czk_vec (sh, mac). */
#define CZK_SCHEME_PLAIN 0
#define CZK_SCHEME_ADDITIVE 1
#define CZK_SCHEME_SPDZ 2
CZK_API int czk_batch_open(czk_ctx* ctx, int scheme, const czk_vec* sh, const czk_vec* mac, czk_vec* out_pub, size_t n);
/* x *= y elementwise on shares with the reference's stub triple source (DummyFieldTripleSource,
 * mpc-algebra/src/wire/field.rs:41-77): Beaver multiplication, two opens.                         */
CZK_API int czk_beaver_batch_mul(czk_ctx* ctx, int scheme, czk_vec* x_sh, czk_vec* x_mac, const czk_vec* y_sh,
                         const czk_vec* y_mac, size_t n);

/* ---- Plonk / KZG10 leaves: replaces the MpcField hooks mpc-algebra/src/wire/field.rs:395-455 ---------
 * -> FieldShare::{batch_inv, batch_div, partial_products} (share/field.rs:135-182, stub inv pairs of wire/field.rs:62-77),
 * DensePolynomial::divide_with_q_and_r by (X - z) (poly/src/polynomial/univariate/mod.rs:133-174; on shares applied to every
 * share vector, share/add.rs:148-156) and KZG10::open without hiding (poly-commit/src/kzg10/mod.rs:196-262).
 * KZG10::commit without hiding (kzg10/mod.rs:141-193) is czk_msm_bases over the resident powers_of_g.               */
CZK_API int czk_vec_prefix_products(czk_ctx* ctx, czk_vec* v, size_t n);   /* v[i] *= v[i-1], plain values (scan) */
CZK_API int czk_vec_batch_inverse(czk_ctx* ctx, czk_vec* v, size_t n);     /* v[i] = 1 / v[i]; a zero -> CZK_ERR_PROTOCOL */
/* q_out (n - 1 coefficients, may be NULL) = p / (X - z), rem_out (host, may be NULL) = p(z). */
CZK_API int czk_poly_div_linear(czk_ctx* ctx, const czk_vec* p, size_t n, const uint64_t z[4], czk_vec* q_out, uint64_t rem_out[4]);
/* x <- shares of 1 / x.  schemes PLAIN, ADDITIVE, SPDZ (x_mac for SPDZ). */
CZK_API int czk_share_batch_inv(czk_ctx* ctx, int scheme, czk_vec* x_sh, czk_vec* x_mac, size_t n);
/* x <- shares of x / y; y returns holding the shares of 1 / y (batch_mul(xs, batch_inv(ys))). */
CZK_API int czk_share_batch_div(czk_ctx* ctx, int scheme, czk_vec* x_sh, czk_vec* x_mac, czk_vec* y_sh, czk_vec* y_mac, size_t n);
/* x[i] <- shares of x[0] * ... * x[i] (the masked prefix-product protocol). */
CZK_API int czk_share_partial_products(czk_ctx* ctx, int scheme, czk_vec* x_sh, czk_vec* x_mac, size_t n);
/* KZG10 open on one coefficient vector (a plain polynomial, or one share component): eval_out = p(z),
 * w_xyz = MSM(powers_of_g, coefficients of p / (X - z)) as an affine-normalised Jacobian triple like czk_msm_g1. */
CZK_API int czk_kzg_open(czk_ctx* ctx, const czk_bases* powers, const czk_vec* p, size_t n, const uint64_t z[4],
                         uint64_t w_xyz[18], uint64_t eval_out[4]);

/* ---- wire format: ark-serialize canonical encodings, host-side -------------------------------------
 * algebra/ff/src/fields/macros.rs:1-87 (Fp: canonical little-endian bytes, flags in the top bits of the last byte),
 * fields/models/quadratic_extension.rs:600-647 (Fq2: c0 | c1), algebra/serialize/src/flags.rs (SWFlags: bit 7 = y is the
 * greater of (y, -y), bit 6 = infinity), algebra/ec/src/models/short_weierstrass_jacobian.rs:792-895 (GroupAffine).
 * Sizes: Fr 32 bytes; G1 48 compressed / 96 uncompressed; G2 96 / 192.  These are what mpc-net messages and
 * `Proof::serialize` carry, so a GPU party can exchange bytes with a stock party.  No device needed, ctx-free.
 * Deserialisation returns CZK_ERR_ARG on what the reference rejects (bad flags, non-canonical element, no such point,
 * and - when check_subgroup != 0, as CanonicalDeserialize does - a point outside the prime-order subgroup).            */
CZK_API int czk_fr_serialize(const uint64_t* fr_mont, size_t n, uint8_t* out);
CZK_API int czk_fr_deserialize(const uint8_t* in, size_t n, uint64_t* fr_mont);
CZK_API int czk_g1_serialize(const uint64_t* xy, const uint8_t* inf, size_t n, int compressed, uint8_t* out);
CZK_API int czk_g2_serialize(const uint64_t* xy, const uint8_t* inf, size_t n, int compressed, uint8_t* out);
CZK_API int czk_g1_deserialize(const uint8_t* in, size_t n, int compressed, int check_subgroup, uint64_t* xy, uint8_t* inf);
CZK_API int czk_g2_deserialize(const uint8_t* in, size_t n, int compressed, int check_subgroup, uint64_t* xy, uint8_t* inf);

/* ---- GSZ20 honest-majority shares: replaces mpc-algebra/src/share/gsz20/mod.rs on the Groth16 path -------
 * n parties, t = (n-1)/2, party j holds p(w^j) over the mixed-radix share domain of size n (n = 2^a or 3*2^a;
 * :94-105).  A share vector is one czk_vec.  The reference's preprocessing stubs are kept (rand() = 1,
 * double_rand() = (1, 1), the king returns the opened value to every party: :378-410, :468-486).
 * A failed degree check or product check returns CZK_ERR_PROTOCOL (the reference assert!s).              */
#define CZK_SCHEME_GSZ 3
/* batch_open / open_degree_vec (:286-299, :434-459): all-gather, interpolate, degree <= `degree`, value at 0. */
CZK_API int czk_gsz_open(czk_ctx* ctx, const czk_vec* sh, unsigned degree, czk_vec* out_pub, size_t n);
/* batch_king_compute with f = identity (:488-524): v <- the value the king opens at `degree`, returned to everyone. */
CZK_API int czk_gsz_king_compute(czk_ctx* ctx, czk_vec* v, unsigned degree, size_t n);
/* batch_mult (:559-594): x <- x * y (local product + mask, king degree reduction 2t -> t); queue_check != 0 keeps
 * the triple for czk_gsz_check_products.  czk_beaver_batch_mul(scheme GSZ) is this with queue_check = 1.   */
CZK_API int czk_gsz_batch_mul(czk_ctx* ctx, czk_vec* x, const czk_vec* y, size_t n, int queue_check);
/* check_accumulated_field_products: hadamard_check -> ip_check over every queued triple (:412-431, :599-808).
 * final_xyz (may be NULL): the three values opened by the last step (x * y == z).                          */
CZK_API int czk_gsz_check_products(czk_ctx* ctx, uint64_t final_xyz[12]);
/* out = { king computations, opens } since context creation. */
CZK_API int czk_gsz_stats(const czk_ctx* ctx, uint64_t out[2]);

/* ---- diagnostics --------------------------------------------------------------------------------- */
/* Bytes this rank actually moved over NVLink { sent, received } since context creation.  czk_net_stats keeps mpc-net's own
 * accounting (what the reference's Stats line would show for the same protocol steps: an SPDZ open counts 3 broadcasts
 * of N-1 full messages, multi.rs:145-174 + channel.rs:50-75); the device path exchanges slices (see csrc/shares.cu), so
 * the two differ by design. */
CZK_API int czk_net_link_bytes(const czk_ctx* ctx, uint64_t out[2]);
/* How additive / SPDZ opens move their slices on this communicator: 1 = NVLink peer memory (every rank maps every other
 * rank's exchange buffers with CUDA IPC; the reduce kernel loads from and stores to the peers directly, NCCL only orders the
 * kernels), -1 = grouped ncclSend / ncclRecv + ncclAllGather (CZK_SHARE_TRANSPORT=nccl, or some peer not addressable),
 * 0 = not decided yet (no open has run) or a single party. */
CZK_API int czk_net_share_transport(const czk_ctx* ctx);
/* N-party additive / SPDZ protocol arithmetic on ONE GPU: party q's vectors are sh[q] (mac[q]); the same kernels, slice
 * geometry and per-party constants as czk_batch_open / czk_beaver_batch_mul run for every simulated party, with the
 * collectives replaced by direct addressing between the parties' buffers.  flags_out[q] != 0: party q's slice of the
 * SPDZ MAC check failed (the product entry points return CZK_ERR_PROTOCOL in that case).  parties <= 16.
 * Replaces nothing in the reference: test infrastructure for share/add.rs:121-125, spdz.rs:166-185, field.rs:97-127. */
CZK_API int czk_diag_sim_batch_open(czk_ctx* ctx, int scheme, int parties, const czk_vec* const* sh, const czk_vec* const* mac,
                                    czk_vec* const* out_pub, size_t n, uint32_t* flags_out);
CZK_API int czk_diag_sim_beaver_mul(czk_ctx* ctx, int scheme, int parties, czk_vec* const* x_sh, czk_vec* const* x_mac,
                                    const czk_vec* const* y_sh, const czk_vec* const* y_mac, size_t n, uint32_t* flags_out);
/* GSZ open_degree_vec (gsz20/mod.rs:440-459) on a caller-supplied party-major matrix of gathered shares (parties x k):
 * out_pub = the values at 0, *flag_out != 0 iff some polynomial has degree > `degree`. */
CZK_API int czk_diag_gsz_open_gathered(czk_ctx* ctx, const czk_vec* gathered, int parties, unsigned degree, size_t k, czk_vec* out_pub,
                                       uint32_t* flag_out);
/* Device timing (CUDA events on the context's stream) of the MSMs run so far on this context, per curve:
 * out = { bucket-accumulation kernel ms (sum), its launch count, terms processed (sum of n), whole-MSM device ms (sum),
 *         (point, window) pairs = upper bound on mixed additions (sum of n * windows) }. */
CZK_API int czk_msm_stats(czk_ctx* ctx, int curve, double out[5], int reset);
/* Bucket accumulation algorithm: 1 (default) = tree of batched affine additions with the XYZZ walk as the fallback for
 * inputs that need P + P / P + (-P), taken from ~2^19 terms up; 0 = the XYZZ walk only; 2 = the tree at every size (so
 * that small test inputs exercise it).  Same results; for A/B measurements and tests. */
CZK_API int czk_msm_set_batched(czk_ctx* ctx, int enabled);
/* out[i] = in[i]^-1 in Fq (n x 6 limbs, Montgomery; 0 -> 0) with the device's binary-Euclid inversion (the one inversion
 * per block of the batched-affine path; algebra/ff/src/fields/macros.rs:368-422).  Host buffers. */
CZK_API int czk_fq_inverse(czk_ctx* ctx, const uint64_t* in, uint64_t* out, size_t n);
/* Integer-pipe microbenchmarks; result = operations per second.  kind: 0 IMAD.WIDE.U32 chain,
 * 1 IMAD lo/hi pair, 2 Fr mul, 3 Fq mul, 4 G1 mixed add. */
CZK_API int czk_microbench(czk_ctx* ctx, int kind, int blocks_per_sm, int threads, int iters, double* ops_per_s, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* CZK_H */

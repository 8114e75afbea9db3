// Groth16 prover loop on shares (include/czk_groth16.h): host-side C++ orchestration of the device
// kernels.  Mirrors mpc-snarks/src/groth/prover.rs:66-177 (create_proof), :216-232 (calculate_coeff) and
// mpc-snarks/src/groth/r1cs_to_qap.rs:47-112 (witness_map) for the benchmark's squaring circuit
// (mpc-snarks/src/proof.rs:304-344), with the share semantics of mpc-algebra/src/share/{add,spdz}.rs and the
// Beaver-on-groups step of share/group.rs:70-109.  The O(1) group operations run on the host (host_field.hpp).
#include <chrono>
#include <cstdlib>
#include <functional>
#include <future>

#include "../../include/czk_groth16.h"
#include "ctx.hpp"
#include "fr_ops.cuh"
#include "launch_count.hpp"

struct czk_pk {
    size_t n_sq = 0, D = 0;                       // squaring circuit: n_sq squarings; any circuit: n_sq = 0
    size_t ncons = 0, ninst = 0, nwit = 0;       // constraints, instance variables (incl. the constant one), witness variables
    unsigned log_d = 0;
    czk_bases* q[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // a, b_g1, b_g2, h, l
    uint64_t vk_g1[36];
    uint64_t vk_g2[72];
    // first entries of the a / b queries (calculate_coeff adds query[0] on the host)
    uint64_t a0[12], b10[12], b20[24];
    uint8_t a0_inf = 0, b10_inf = 0, b20_inf = 0;
    std::vector<uint64_t> gamma_abc;  // ninst x 12 limbs when the key was generated here (czk_groth16_setup*), else empty
    // the a- and b_g1-query MSMs take the same scalars; when the two queries also agree on which bases are infinity (and on
    // table shape) the second one reuses the first one's digit sort (msm_run's reuse_plan)
    bool ab_share_plan = false;
};

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// A constraint system in CSR form on the device (the reference's ConstraintMatrices, relations/src/r1cs): for matrix
// m in {A, B, C}, row i holds entries row_ptr[m][i] .. row_ptr[m][i+1] of (col[m], coeff[m]).
struct czk_r1cs {
    size_t ncons = 0, ninst = 0, nwit = 0;
    uint64_t* row_ptr[3] = {nullptr, nullptr, nullptr};
    uint32_t* col[3] = {nullptr, nullptr, nullptr};
    uint32_t* coeff[3] = {nullptr, nullptr, nullptr};
    size_t nnz[3] = {0, 0, 0};
};

namespace czk {
// evaluate_constraint (mpc-snarks/src/groth/r1cs_to_qap.rs:12-41) for every row: out[i] = sum_k coeff[k] * assign[col[k]].
// One thread per row (rows of real circuits hold a handful of terms); the assignment is a share vector, the
// coefficients are public, so the map is linear in the shares.
__global__ void k_r1cs_eval(uint32_t* __restrict__ out, const uint64_t* __restrict__ row_ptr, const uint32_t* __restrict__ col,
                            const uint32_t* __restrict__ coeff, const uint32_t* __restrict__ assign, size_t ncons) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < ncons; i += (size_t)gridDim.x * blockDim.x) {
        Fr acc = Fr::zero();
        for (uint64_t k = row_ptr[i]; k < row_ptr[i + 1]; k++) {
            const uint4* cq = reinterpret_cast<const uint4*>(coeff) + 2 * k;
            const uint4* vq = reinterpret_cast<const uint4*>(assign) + 2 * (size_t)col[k];
            uint4 c0 = cq[0], c1 = cq[1], v0 = vq[0], v1 = vq[1];
            Fr c, v;
            c.l[0] = c0.x; c.l[1] = c0.y; c.l[2] = c0.z; c.l[3] = c0.w; c.l[4] = c1.x; c.l[5] = c1.y; c.l[6] = c1.z; c.l[7] = c1.w;
            v.l[0] = v0.x; v.l[1] = v0.y; v.l[2] = v0.z; v.l[3] = v0.w; v.l[4] = v1.x; v.l[5] = v1.y; v.l[6] = v1.z; v.l[7] = v1.w;
            acc = Fr::add(acc, Fr::mul(v, c));
        }
        uint4* o = reinterpret_cast<uint4*>(out) + 2 * i;
        o[0] = make_uint4(acc.l[0], acc.l[1], acc.l[2], acc.l[3]);
        o[1] = make_uint4(acc.l[4], acc.l[5], acc.l[6], acc.l[7]);
    }
}
}  // namespace czk

int czk_r1cs_upload(czk_ctx* ctx, size_t ncons, size_t ninst, size_t nwit, const uint64_t* const row_ptr[3],
                    const uint32_t* const col[3], const uint64_t* const coeff[3], czk_r1cs** out) {
    if (!ctx || !out || !row_ptr || !col || !coeff || !ncons || !ninst) return fail(ctx, CZK_ERR_ARG, "czk_r1cs_upload: argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    struct Guard {
        czk_ctx* ctx;
        czk_r1cs* r;
        ~Guard() {
            if (r) czk_r1cs_free(ctx, r);
        }
    } guard{ctx, new czk_r1cs()};
    czk_r1cs* r = guard.r;
    r->ncons = ncons;
    r->ninst = ninst;
    r->nwit = nwit;
    for (int m = 0; m < 3; m++) {
        if (!row_ptr[m] || row_ptr[m][0] != 0) return fail(ctx, CZK_ERR_ARG, "czk_r1cs_upload: row_ptr");
        size_t nnz = (size_t)row_ptr[m][ncons];
        for (size_t i = 0; i < ncons; i++)
            if (row_ptr[m][i] > row_ptr[m][i + 1]) return fail(ctx, CZK_ERR_ARG, "czk_r1cs_upload: row_ptr not monotone");
        for (size_t k = 0; k < nnz; k++)
            if (col[m][k] >= ninst + nwit) return fail(ctx, CZK_ERR_ARG, "czk_r1cs_upload: variable index out of range");
        r->nnz[m] = nnz;
        CUDA_TRY(ctx, cudaMalloc((void**)&r->row_ptr[m], (ncons + 1) * 8));
        CUDA_TRY(ctx, cudaMalloc((void**)&r->col[m], (nnz ? nnz : 1) * 4));
        CUDA_TRY(ctx, cudaMalloc((void**)&r->coeff[m], (nnz ? nnz : 1) * 32));
        CUDA_TRY(ctx, cudaMemcpyAsync(r->row_ptr[m], row_ptr[m], (ncons + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        if (nnz) {
            CUDA_TRY(ctx, cudaMemcpyAsync(r->col[m], col[m], nnz * 4, cudaMemcpyHostToDevice, ctx->stream));
            CUDA_TRY(ctx, cudaMemcpyAsync(r->coeff[m], coeff[m], nnz * 32, cudaMemcpyHostToDevice, ctx->stream));
        }
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    guard.r = nullptr;
    *out = r;
    return CZK_OK;
}
void czk_r1cs_free(czk_ctx* ctx, czk_r1cs* r) {
    if (!r) return;
    if (ctx) cudaSetDevice(ctx->device);
    for (int m = 0; m < 3; m++) {
        cudaFree(r->row_ptr[m]);
        cudaFree(r->col[m]);
        cudaFree(r->coeff[m]);
    }
    delete r;
}

static size_t domain_size_for(size_t n_sq, unsigned* log_d) {
    size_t need = n_sq + 2, d = 1;  // num_constraints + num_instance_variables (r1cs_to_qap.rs:63-64)
    unsigned l = 0;
    while (d < need) {
        d <<= 1;
        l++;
    }
    *log_d = l;
    return d;
}

static int pk_finish(czk_ctx* ctx, czk_pk* pk) {
    uint8_t inf = 0;
    // merged-window tables for the five queries (one-off per key); CZK_PRECOMPUTE=0 keeps the windowed MSM
    const char* env = getenv("CZK_PRECOMPUTE");
    if (!env || atoi(env) != 0)
        for (int i = 0; i < 5; i++) CZK_TRY(czk_bases_precompute(ctx, pk->q[i], 0));
    CZK_TRY(czk_bases_download(ctx, pk->q[0], 0, 1, pk->a0, &inf));
    pk->a0_inf = inf;
    CZK_TRY(czk_bases_download(ctx, pk->q[1], 0, 1, pk->b10, &inf));
    pk->b10_inf = inf;
    CZK_TRY(czk_bases_download(ctx, pk->q[2], 0, 1, pk->b20, &inf));
    pk->b20_inf = inf;
    const czk_bases *qa = pk->q[0], *qb = pk->q[1];
    const char* share_env = getenv("CZK_AB_SHARE_PLAN");
    pk->ab_share_plan = false;
    if (!(share_env && atoi(share_env) == 0) && qa->table && qb->table && qa->pre_c == qb->pre_c && qa->n == qb->n && qa->n > 1 &&
        (qa->inf != nullptr) == (qb->inf != nullptr)) {
        uint32_t differ = 0;
        if (qa->inf) {  // entry 0 is added on the host (calculate_coeff): the MSMs start at base 1
            CUDA_TRY(ctx, cudaMemsetAsync(ctx->flag, 0, 4, ctx->stream));
            CUDA_TRY(ctx, msm_flags_differ(qa->inf + 1, qb->inf + 1, qa->n - 1, ctx->flag, ctx->stream));
            CUDA_TRY(ctx, cudaMemcpyAsync(&differ, ctx->flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            if (differ) CUDA_TRY(ctx, cudaMemsetAsync(ctx->flag, 0, 4, ctx->stream));  // its other users expect it clear
        }
        pk->ab_share_plan = differ == 0;
    }
    return CZK_OK;
}

// a key under construction: freed (with whatever queries and tables it already holds) unless the constructor releases it
struct PkGuard {
    czk_ctx* ctx;
    czk_pk* pk;
    ~PkGuard() {
        if (pk) czk_groth16_pk_free(ctx, pk);
    }
    czk_pk* release() {
        czk_pk* p = pk;
        pk = nullptr;
        return p;
    }
};

int czk_groth16_pk_upload(czk_ctx* ctx, size_t n_sq, const uint64_t* a_query, const uint8_t* a_inf, const uint64_t* b_g1_query,
                          const uint8_t* b1_inf, const uint64_t* b_g2_query, const uint8_t* b2_inf, const uint64_t* h_query,
                          const uint8_t* h_inf, const uint64_t* l_query, const uint8_t* l_inf, const uint64_t vk_g1[36],
                          const uint64_t vk_g2[72], czk_pk** out) {
    if (!ctx || !out || !n_sq || !a_query || !b_g1_query || !b_g2_query || !h_query || !l_query || !vk_g1 || !vk_g2)
        return fail(ctx, CZK_ERR_ARG, "czk_groth16_pk_upload: null argument");
    czk_pk* pk = new czk_pk();
    PkGuard guard{ctx, pk};
    pk->n_sq = n_sq;
    pk->ncons = n_sq;
    pk->ninst = 2;
    pk->nwit = n_sq;
    pk->D = domain_size_for(n_sq, &pk->log_d);
    CZK_TRY(czk_bases_upload(ctx, 1, a_query, a_inf, n_sq + 2, &pk->q[0]));
    CZK_TRY(czk_bases_upload(ctx, 1, b_g1_query, b1_inf, n_sq + 2, &pk->q[1]));
    CZK_TRY(czk_bases_upload(ctx, 2, b_g2_query, b2_inf, n_sq + 2, &pk->q[2]));
    CZK_TRY(czk_bases_upload(ctx, 1, h_query, h_inf, pk->D - 1, &pk->q[3]));
    CZK_TRY(czk_bases_upload(ctx, 1, l_query, l_inf, n_sq, &pk->q[4]));
    std::memcpy(pk->vk_g1, vk_g1, sizeof pk->vk_g1);
    std::memcpy(pk->vk_g2, vk_g2, sizeof pk->vk_g2);
    CZK_TRY(pk_finish(ctx, pk));
    *out = guard.release();
    return CZK_OK;
}

int czk_groth16_pk_upload_r1cs(czk_ctx* ctx, size_t ncons, size_t ninst, size_t nwit, const uint64_t* a_query, const uint8_t* a_inf,
                               const uint64_t* b_g1_query, const uint8_t* b1_inf, const uint64_t* b_g2_query, const uint8_t* b2_inf,
                               const uint64_t* h_query, const uint8_t* h_inf, const uint64_t* l_query, const uint8_t* l_inf,
                               const uint64_t vk_g1[36], const uint64_t vk_g2[72], czk_pk** out) {
    if (!ctx || !out || !ncons || !ninst || !a_query || !b_g1_query || !b_g2_query || !h_query || (nwit && !l_query) || !vk_g1 || !vk_g2)
        return fail(ctx, CZK_ERR_ARG, "czk_groth16_pk_upload_r1cs: null argument");
    czk_pk* pk = new czk_pk();
    PkGuard guard{ctx, pk};
    pk->n_sq = 0;
    pk->ncons = ncons;
    pk->ninst = ninst;
    pk->nwit = nwit;
    size_t need = ncons + ninst;  // r1cs_to_qap.rs:63-64
    pk->D = 1;
    pk->log_d = 0;
    while (pk->D < need) pk->D <<= 1, pk->log_d++;
    const size_t nvar = ninst + nwit;
    CZK_TRY(czk_bases_upload(ctx, 1, a_query, a_inf, nvar, &pk->q[0]));
    CZK_TRY(czk_bases_upload(ctx, 1, b_g1_query, b1_inf, nvar, &pk->q[1]));
    CZK_TRY(czk_bases_upload(ctx, 2, b_g2_query, b2_inf, nvar, &pk->q[2]));
    CZK_TRY(czk_bases_upload(ctx, 1, h_query, h_inf, pk->D - 1, &pk->q[3]));
    CZK_TRY(czk_bases_upload(ctx, 1, l_query, l_inf, nwit, &pk->q[4]));
    std::memcpy(pk->vk_g1, vk_g1, sizeof pk->vk_g1);
    std::memcpy(pk->vk_g2, vk_g2, sizeof pk->vk_g2);
    CZK_TRY(pk_finish(ctx, pk));
    *out = guard.release();
    return CZK_OK;
}

// ------------------------------------------------------------------------------------------ CRS generation
// groth16/src/generator.rs:34-221 (generate_parameters) with the toxic waste supplied by the caller:
// toxic = alpha | beta | gamma | delta | tau | g1_scalar | g2_scalar (7 Montgomery Fr; the reference draws the two generators
// as random group elements, here they are g1_scalar * G1 and g2_scalar * G2 of the curve's standard generators).
// The O(D) scalar work (Lagrange coefficients at tau, the QAP polynomials at tau) runs on the host; the five fixed-base
// multi-scalar multiplications - where the reference spends its time - run on the device and leave the queries resident.
template <class HF, int LIMBS>
static void point_to_affine_limbs(const HPoint<HF>& p, uint64_t* xy) {  // infinity is written as (0, 1)
    HF ax, ay;
    if (!p.to_affine(ax, ay)) {
        ax = HF::zero();
        ay = HF::one();
    }
    ax.to_limbs(xy);
    ay.to_limbs(xy + LIMBS);
}

static int setup_core(czk_ctx* ctx, size_t ncons, size_t ninst, size_t nwit, const uint64_t* const row_ptr[3],
                      const uint32_t* const col[3], const uint64_t* const coeff[3], const uint64_t toxic[28], czk_pk* pk) {
    const HFr alpha = HFr::from_limbs(toxic), beta = HFr::from_limbs(toxic + 4), gamma = HFr::from_limbs(toxic + 8),
              delta = HFr::from_limbs(toxic + 12), tau = HFr::from_limbs(toxic + 16), s1 = HFr::from_limbs(toxic + 20),
              s2 = HFr::from_limbs(toxic + 24);
    if (gamma.is_zero() || delta.is_zero() || s1.is_zero() || s2.is_zero()) return fail(ctx, CZK_ERR_ARG, "groth16 setup: zero toxic value");
    const size_t nvar = ninst + nwit, D = pk->D;
    // Lagrange coefficients at tau (radix2/mod.rs:119-181): u_i = Z(tau) w^i / (D (tau - w^i)), one batched inversion
    HFr w = HFr::from_limbs(FrParams::ROOT_2_47_64);
    for (unsigned i = pk->log_d; i < FrParams::TWO_ADICITY; i++) w = HFr::sqr(w);
    const HFr zt = HFr::sub(HFr::pow_u64(tau, (uint64_t)D), HFr::one()), dfe = HFr::from_u64((uint64_t)D);
    std::vector<HFr> u(D), pre(D);
    {
        HFr wi = HFr::one(), acc = HFr::one();
        for (size_t i = 0; i < D; i++) {
            HFr den = HFr::mul(HFr::sub(tau, wi), dfe);
            if (den.is_zero()) return fail(ctx, CZK_ERR_ARG, "groth16 setup: tau lies in the evaluation domain");
            u[i] = den;
            pre[i] = acc;
            acc = HFr::mul(acc, den);
            wi = HFr::mul(wi, w);
        }
        HFr inv = HFr::inv(acc);
        // walk back: 1/den_i = inv * pre_i ; w^i recomputed from the top
        std::vector<HFr> wp(D);
        wi = HFr::one();
        for (size_t i = 0; i < D; i++) {
            wp[i] = wi;
            wi = HFr::mul(wi, w);
        }
        for (size_t i = D; i-- > 0;) {
            HFr di = HFr::mul(inv, pre[i]);
            inv = HFr::mul(inv, u[i]);
            u[i] = HFr::mul(HFr::mul(zt, wp[i]), di);
        }
    }
    pre.clear();
    pre.shrink_to_fit();
    // a_i(tau), b_i(tau), c_i(tau) per variable (generator.rs:109-121, r1cs_to_qap.rs:51-92)
    std::vector<HFr> qa(nvar, HFr::zero()), qb(nvar, HFr::zero()), qc(nvar, HFr::zero());
    for (size_t i = 0; i < ninst; i++) qa[i] = u[ncons + i];
    std::vector<HFr>* dst[3] = {&qa, &qb, &qc};
    for (int m = 0; m < 3; m++)
        for (size_t i = 0; i < ncons; i++)
            for (uint64_t k = row_ptr[m][i]; k < row_ptr[m][i + 1]; k++) {
                HFr& x = (*dst[m])[col[m][k]];
                x = HFr::add(x, HFr::mul(u[i], HFr::from_limbs(coeff[m] + 4 * k)));
            }
    const HFr gamma_inv = HFr::inv(gamma), delta_inv = HFr::inv(delta);
    std::vector<uint64_t> sa(nvar * 4), sb(nvar * 4), sl((nwit ? nwit : 1) * 4), sh((D - 1) * 4);
    std::vector<HFr> gabc(ninst);
    for (size_t i = 0; i < nvar; i++) {
        qa[i].to_limbs(sa.data() + 4 * i);
        qb[i].to_limbs(sb.data() + 4 * i);
        HFr x = HFr::add(HFr::add(HFr::mul(beta, qa[i]), HFr::mul(alpha, qb[i])), qc[i]);
        if (i < ninst) gabc[i] = HFr::mul(x, gamma_inv);
        else HFr::mul(x, delta_inv).to_limbs(sl.data() + 4 * (i - ninst));
    }
    {
        HFr f = HFr::mul(zt, delta_inv), pw = HFr::one();
        for (size_t i = 0; i + 1 < D; i++) {
            HFr::mul(f, pw).to_limbs(sh.data() + 4 * i);
            pw = HFr::mul(pw, tau);
        }
    }
    u.clear();
    u.shrink_to_fit();
    // generators and the verifying-key elements (host: O(ninst) scalar multiplications)
    auto canon = [](const HFr& x, uint64_t k[4]) { x.from_mont().to_limbs(k); };
    uint64_t k[4];
    HG1 g1 = HG1::from_affine(HFq::from_limbs(CurveConsts::G1_GEN), HFq::from_limbs(CurveConsts::G1_GEN + 6));
    HG2 g2 = HG2::from_affine(HFq2::from_limbs(CurveConsts::G2_GEN), HFq2::from_limbs(CurveConsts::G2_GEN + 12));
    canon(s1, k);
    g1 = HG1::mul(g1, k, 4);
    canon(s2, k);
    g2 = HG2::mul(g2, k, 4);
    uint64_t g1_xy[12], g2_xy[24];
    point_to_affine_limbs<HFq, 6>(g1, g1_xy);
    point_to_affine_limbs<HFq2, 12>(g2, g2_xy);
    const HFr v1[3] = {alpha, beta, delta}, v2[3] = {beta, gamma, delta};
    for (int i = 0; i < 3; i++) {
        canon(v1[i], k);
        point_to_affine_limbs<HFq, 6>(HG1::mul(g1, k, 4), pk->vk_g1 + 12 * i);
        canon(v2[i], k);
        point_to_affine_limbs<HFq2, 12>(HG2::mul(g2, k, 4), pk->vk_g2 + 24 * i);
    }
    pk->gamma_abc.assign(ninst * 12, 0);
    for (size_t i = 0; i < ninst; i++) {
        canon(gabc[i], k);
        point_to_affine_limbs<HFq, 6>(HG1::mul(g1, k, 4), pk->gamma_abc.data() + 12 * i);
    }
    // the five queries: fixed-base MSMs on the device, left resident
    czk_vec *va = nullptr, *vb = nullptr, *vl = nullptr, *vh = nullptr;
    auto cleanup = [&](int rc) {
        for (czk_vec* v : {va, vb, vl, vh}) czk_vec_free(ctx, v);
        return rc;
    };
    int rc;
    if ((rc = czk_vec_alloc(ctx, nvar, &va)) || (rc = czk_vec_alloc(ctx, nvar, &vb)) || (rc = czk_vec_alloc(ctx, nwit ? nwit : 1, &vl)) ||
        (rc = czk_vec_alloc(ctx, D - 1, &vh)))
        return cleanup(rc);
    if ((rc = czk_vec_upload(ctx, va, 0, sa.data(), nvar)) || (rc = czk_vec_upload(ctx, vb, 0, sb.data(), nvar)) ||
        (nwit && (rc = czk_vec_upload(ctx, vl, 0, sl.data(), nwit))) || (rc = czk_vec_upload(ctx, vh, 0, sh.data(), D - 1)))
        return cleanup(rc);
    if ((rc = czk_fixed_base_msm(ctx, 1, g1_xy, va, 0, nvar, &pk->q[0])) || (rc = czk_fixed_base_msm(ctx, 1, g1_xy, vb, 0, nvar, &pk->q[1])) ||
        (rc = czk_fixed_base_msm(ctx, 2, g2_xy, vb, 0, nvar, &pk->q[2])) || (rc = czk_fixed_base_msm(ctx, 1, g1_xy, vh, 0, D - 1, &pk->q[3])) ||
        (rc = czk_fixed_base_msm(ctx, 1, g1_xy, vl, 0, nwit, &pk->q[4])))
        return cleanup(rc);
    return cleanup(CZK_OK);
}

int czk_groth16_setup_r1cs(czk_ctx* ctx, size_t ncons, size_t ninst, size_t nwit, const uint64_t* const row_ptr[3],
                           const uint32_t* const col[3], const uint64_t* const coeff[3], const uint64_t toxic[28], czk_pk** out) {
    if (!ctx || !out || !row_ptr || !col || !coeff || !toxic || !ncons || !ninst)
        return fail(ctx, CZK_ERR_ARG, "czk_groth16_setup_r1cs: argument");
    for (int m = 0; m < 3; m++) {
        if (!row_ptr[m] || row_ptr[m][0] != 0) return fail(ctx, CZK_ERR_ARG, "czk_groth16_setup_r1cs: row_ptr");
        for (uint64_t k = 0; k < row_ptr[m][ncons]; k++)
            if (col[m][k] >= ninst + nwit) return fail(ctx, CZK_ERR_ARG, "czk_groth16_setup_r1cs: variable index out of range");
    }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    czk_pk* pk = new czk_pk();
    pk->n_sq = 0;
    pk->ncons = ncons;
    pk->ninst = ninst;
    pk->nwit = nwit;
    pk->D = 1;
    pk->log_d = 0;
    while (pk->D < ncons + ninst) pk->D <<= 1, pk->log_d++;
    int rc = setup_core(ctx, ncons, ninst, nwit, row_ptr, col, coeff, toxic, pk);
    if (rc == CZK_OK) rc = pk_finish(ctx, pk);
    if (rc != CZK_OK) {
        czk_groth16_pk_free(ctx, pk);
        return rc;
    }
    *out = pk;
    return CZK_OK;
}

// the benchmark circuit: variables [one, out, w_0 .. w_{n-1}], constraint i: w_i * w_i = w_{i+1} (the last: = out)
int czk_groth16_setup(czk_ctx* ctx, size_t n_sq, const uint64_t toxic[28], czk_pk** out) {
    if (!ctx || !out || !toxic || !n_sq) return fail(ctx, CZK_ERR_ARG, "czk_groth16_setup: argument");
    std::vector<uint64_t> rp(n_sq + 1), ones(n_sq * 4);
    std::vector<uint32_t> ab(n_sq), cc(n_sq);
    const HFr one = HFr::one();
    for (size_t i = 0; i <= n_sq; i++) rp[i] = i;
    for (size_t i = 0; i < n_sq; i++) {
        ab[i] = (uint32_t)(2 + i);
        cc[i] = (uint32_t)(i + 1 < n_sq ? 2 + i + 1 : 1);
        one.to_limbs(ones.data() + 4 * i);
    }
    const uint64_t* rps[3] = {rp.data(), rp.data(), rp.data()};
    const uint32_t* cols[3] = {ab.data(), ab.data(), cc.data()};
    const uint64_t* cfs[3] = {ones.data(), ones.data(), ones.data()};
    czk_pk* pk = nullptr;
    CZK_TRY(czk_groth16_setup_r1cs(ctx, n_sq, 2, n_sq, rps, cols, cfs, toxic, &pk));
    pk->n_sq = n_sq;  // the squaring entry points (czk_groth16_prove) accept this key
    *out = pk;
    return CZK_OK;
}

int czk_groth16_pk_gamma_abc(const czk_pk* pk, uint64_t* out, size_t ninst) {
    if (!pk || !out || pk->gamma_abc.size() != ninst * 12) return fail(nullptr, CZK_ERR_ARG, "czk_groth16_pk_gamma_abc: this key carries no gamma_abc of that size");
    std::memcpy(out, pk->gamma_abc.data(), ninst * 12 * 8);
    return CZK_OK;
}

int czk_groth16_pk_synthetic(czk_ctx* ctx, size_t n_sq, uint64_t seed, czk_pk** out) {
    if (!ctx || !out || !n_sq) return fail(ctx, CZK_ERR_ARG, "czk_groth16_pk_synthetic: argument");
    czk_pk* pk = new czk_pk();
    PkGuard guard{ctx, pk};
    pk->n_sq = n_sq;
    pk->ncons = n_sq;
    pk->ninst = 2;
    pk->nwit = n_sq;
    pk->D = domain_size_for(n_sq, &pk->log_d);
    // Groth16 queries hold the point at infinity for variables absent from a matrix: flag every 1024th entry
    CZK_TRY(czk_bases_synthetic(ctx, 1, seed * 8 + 1, n_sq + 2, 1024, &pk->q[0]));
    CZK_TRY(czk_bases_synthetic(ctx, 1, seed * 8 + 2, n_sq + 2, 1024, &pk->q[1]));
    CZK_TRY(czk_bases_synthetic(ctx, 2, seed * 8 + 3, n_sq + 2, 1024, &pk->q[2]));
    CZK_TRY(czk_bases_synthetic(ctx, 1, seed * 8 + 4, pk->D - 1, 0, &pk->q[3]));
    CZK_TRY(czk_bases_synthetic(ctx, 1, seed * 8 + 5, n_sq, 1024, &pk->q[4]));
    // vk points: further synthetic points (downloaded from two tiny synthetic sets)
    czk_bases *v1 = nullptr, *v2 = nullptr;
    CZK_TRY(czk_bases_synthetic(ctx, 1, seed * 8 + 6, 3, 0, &v1));
    CZK_TRY(czk_bases_synthetic(ctx, 2, seed * 8 + 7, 3, 0, &v2));
    CZK_TRY(czk_bases_download(ctx, v1, 0, 3, pk->vk_g1, nullptr));
    CZK_TRY(czk_bases_download(ctx, v2, 0, 3, pk->vk_g2, nullptr));
    czk_bases_free(ctx, v1);
    czk_bases_free(ctx, v2);
    CZK_TRY(pk_finish(ctx, pk));
    *out = guard.release();
    return CZK_OK;
}

void czk_groth16_pk_free(czk_ctx* ctx, czk_pk* pk) {
    if (!pk) return;
    for (int i = 0; i < 5; i++) czk_bases_free(ctx, pk->q[i]);
    delete pk;
}
size_t czk_groth16_pk_domain_size(const czk_pk* pk) { return pk ? pk->D : 0; }
const czk_bases* czk_groth16_pk_query(const czk_pk* pk, int which) { return (pk && which >= 0 && which < 5) ? pk->q[which] : nullptr; }
int czk_groth16_pk_vk(const czk_pk* pk, uint64_t vk_g1[36], uint64_t vk_g2[72]) {
    if (!pk) return CZK_ERR_ARG;
    std::memcpy(vk_g1, pk->vk_g1, sizeof pk->vk_g1);
    std::memcpy(vk_g2, pk->vk_g2, sizeof pk->vk_g2);
    return CZK_OK;
}
int czk_groth16_last_phases(const czk_ctx* ctx, double out_ms[8]) {
    if (!ctx || !out_ms) return CZK_ERR_ARG;
    for (int i = 0; i < 8; i++) out_ms[i] = ctx->phases[i];
    return CZK_OK;
}

// ------------------------------------------------------------------------------------------ witness map
struct ShareVecs {
    czk_vec *a = nullptr, *b = nullptr, *c = nullptr;       // value component
    czk_vec *am = nullptr, *bm = nullptr, *cm = nullptr;    // SPDZ MAC component
    czk_vec* chain = nullptr;                               // n_sq + 1 uploaded shares
    czk_vec* assign = nullptr;                              // [out, w_0 .. w_{n-1}]
    czk_vec* full = nullptr;                                // any circuit: [instance, witness] shares
};
static void free_share_vecs(czk_ctx* ctx, ShareVecs& v) {
    for (czk_vec* p : {v.a, v.b, v.c, v.am, v.bm, v.cm, v.chain, v.assign, v.full}) czk_vec_free(ctx, p);
    v = ShareVecs();
}

// r1cs_to_qap.rs:66-110 on this party's shares.  On return v.a (and v.am) hold h; v.chain / v.assign are filled.
// defer_check: leave the SPDZ MAC verdict of the Beaver product on the device and do not wait for the stream (the prover
// reads the verdict once, at the end of the proof, so the whole proof is enqueued without a host round trip).
static int witness_map_transforms(czk_ctx* ctx, int scheme, unsigned log_d, ShareVecs& v, bool defer_check = false);
// cs != nullptr: any circuit, full_sh = this party's shares of [instance, witness] (host); else the squaring chain.
// after_inputs (optional) runs once the share vectors the MSMs need (chain / assign / full) are enqueued, before the
// transforms: the prover uses it to put its witness-only MSMs in flight under the witness map.
static int witness_map_dev(czk_ctx* ctx, int scheme, size_t n_sq, unsigned log_d, const uint64_t* chain_sh,
                           const czk_vec* chain_dev, ShareVecs& v, const czk_r1cs* cs = nullptr, const uint64_t* full_sh = nullptr,
                           const std::function<int()>& after_inputs = nullptr, bool defer_check = false) {
    const size_t D = (size_t)1 << log_d;
    const bool spdz = scheme == CZK_SCHEME_SPDZ;
    double t0 = now_ms();
    CZK_TRY(czk_vec_alloc(ctx, D, &v.a));
    CZK_TRY(czk_vec_alloc(ctx, D, &v.b));
    CZK_TRY(czk_vec_alloc(ctx, D, &v.c));
    if (cs) {
        const size_t nvar = cs->ninst + cs->nwit;
        CZK_TRY(czk_vec_alloc(ctx, nvar, &v.full));
        CUDA_TRY(ctx, cudaMemcpyAsync(czk_vec_device_ptr(v.full), full_sh, nvar * 32, cudaMemcpyHostToDevice, ctx->stream));
        // a[i] = <A_i, z>, b[i] = <B_i, z>, c[i] = <C_i, z>;  a[ncons + i] = z[i] for the instance variables
        czk_vec* dst[3] = {v.a, v.b, v.c};
        size_t blocks = (cs->ncons + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        for (int m = 0; m < 3; m++) {
            czk::k_r1cs_eval<<<(unsigned)blocks, 256, 0, ctx->stream>>>((uint32_t*)dst[m]->d, cs->row_ptr[m], cs->col[m], cs->coeff[m],
                                                                       (const uint32_t*)v.full->d, cs->ncons);
            CZK_LAUNCHED();
        }
        CUDA_TRY(ctx, cudaGetLastError());
        CZK_TRY(czk_vec_copy(ctx, v.a, cs->ncons, v.full, 0, cs->ninst));
        if (spdz) {
            CZK_TRY(czk_vec_alloc(ctx, D, &v.am));
            CZK_TRY(czk_vec_alloc(ctx, D, &v.bm));
            CZK_TRY(czk_vec_alloc(ctx, D, &v.cm));
            CZK_TRY(czk_vec_copy(ctx, v.am, 0, v.a, 0, D));
            CZK_TRY(czk_vec_copy(ctx, v.bm, 0, v.b, 0, D));
            CZK_TRY(czk_vec_copy(ctx, v.cm, 0, v.c, 0, D));
        }
        ctx->phases[0] = now_ms() - t0;
        if (after_inputs) CZK_TRY(after_inputs());
        return witness_map_transforms(ctx, scheme, log_d, v, defer_check);
    }
    CZK_TRY(czk_vec_alloc(ctx, n_sq + 1, &v.chain));
    CZK_TRY(czk_vec_alloc(ctx, n_sq + 1, &v.assign));
    if (chain_dev) CZK_TRY(czk_vec_copy(ctx, v.chain, 0, chain_dev, 0, n_sq + 1));
    else CUDA_TRY(ctx, cudaMemcpyAsync(czk_vec_device_ptr(v.chain), chain_sh, (n_sq + 1) * 32, cudaMemcpyHostToDevice, ctx->stream));
    // full_assignment = [one, out] ++ witness;  assignment (prover.rs:118) = [out] ++ witness
    CZK_TRY(czk_vec_copy(ctx, v.assign, 0, v.chain, n_sq, 1));
    CZK_TRY(czk_vec_copy(ctx, v.assign, 1, v.chain, 0, n_sq));
    // a[i] = b[i] = w_i, c[i] = w_{i+1} (c[n-1] = out);  a[n] = one, a[n+1] = out   (r1cs_to_qap.rs:70-83)
    CZK_TRY(czk_vec_copy(ctx, v.a, 0, v.chain, 0, n_sq));
    CZK_TRY(czk_vec_copy(ctx, v.b, 0, v.chain, 0, n_sq));
    CZK_TRY(czk_vec_copy(ctx, v.c, 0, v.chain, 1, n_sq));
    CZK_TRY(czk_vec_copy(ctx, v.a, n_sq + 1, v.chain, n_sq, 1));
    // Public(1) lowered to share form: the king holds 1 (add.rs:88-92; SPDZ mac = 1 * mac_share, spdz.rs:132-137)
    // GSZ: shift adds the public value at EVERY party (gsz20/mod.rs:270-273)
    HFr one_val = (scheme == CZK_SCHEME_PLAIN || scheme == CZK_SCHEME_GSZ || ctx->rank == 0) ? HFr::one() : HFr::zero();
    CUDA_TRY(ctx, cudaMemcpyAsync(czk_vec_device_ptr(v.a) + 4 * n_sq, one_val.l, 32, cudaMemcpyHostToDevice, ctx->stream));
    if (spdz) {
        // from_add_shared: mac = share * mac() with the MAC key stubbed to 1 (spdz.rs:41-47,138-143); the MAC vectors
        // then go through every linear map separately, as in spdz.rs:186-208
        CZK_TRY(czk_vec_alloc(ctx, D, &v.am));
        CZK_TRY(czk_vec_alloc(ctx, D, &v.bm));
        CZK_TRY(czk_vec_alloc(ctx, D, &v.cm));
        CZK_TRY(czk_vec_copy(ctx, v.am, 0, v.a, 0, D));
        CZK_TRY(czk_vec_copy(ctx, v.bm, 0, v.b, 0, D));
        CZK_TRY(czk_vec_copy(ctx, v.cm, 0, v.c, 0, D));
    }
    ctx->phases[0] = now_ms() - t0;
    if (after_inputs) CZK_TRY(after_inputs());
    return witness_map_transforms(ctx, scheme, log_d, v, defer_check);
}

static int witness_map_transforms(czk_ctx* ctx, int scheme, unsigned log_d, ShareVecs& v, bool defer_check) {
    const size_t D = (size_t)1 << log_d;
    const bool spdz = scheme == CZK_SCHEME_SPDZ;
    double t0 = now_ms();
    if (ctx->ev_phase[0]) CUDA_TRY(ctx, cudaEventRecord(ctx->ev_phase[0], ctx->stream));
    czk_vec* comps[2][3] = {{v.a, v.b, v.c}, {v.am, v.bm, v.cm}};
    const int ncomp = spdz ? 2 : 1;
    // ifft_in_place + coset_fft_in_place on a, b and c (r1cs_to_qap.rs:85-90,95-98): none of them depends on the product,
    // so all (3, or 6 with the SPDZ MAC vectors, which go through every linear map separately: spdz.rs:186-208) run as
    // ONE batched transform pair - an inverse DIF feeding a forward DIT, no reordering pass
    {
        czk_vec* all[6];
        int cnt = 0;
        for (int k = 0; k < ncomp; k++)
            for (int j = 0; j < 3; j++) all[cnt++] = comps[k][j];
        CZK_TRY(czk_ntt_vec_batch(ctx, all, cnt, log_d, CZK_NTT_IFFT_COSET_FFT));
    }
    // F::batch_product_in_place(&mut ab, &b)
    if (defer_check && (scheme == CZK_SCHEME_SPDZ || scheme == CZK_SCHEME_ADDITIVE)) CZK_TRY(sh_beaver_mul_enqueue(ctx, scheme, v.a, v.am, v.b, v.bm, D));
    else CZK_TRY(czk_beaver_batch_mul(ctx, scheme, v.a, v.am, v.b, v.bm, D));
    for (int k = 0; k < ncomp; k++) {
        CZK_TRY(czk_vec_sub(ctx, comps[k][0], comps[k][2], D));                 // ab -= c
        CZK_TRY(czk_vec_divide_by_vanishing_on_coset(ctx, comps[k][0], log_d));  // /= Z_H(g)
    }
    {
        czk_vec* ab[2] = {comps[0][0], comps[1][0]};
        CZK_TRY(czk_ntt_vec_batch(ctx, ab, ncomp, log_d, CZK_NTT_COSET_IFFT));  // coset_ifft_in_place
    }
    if (ctx->ev_phase[1]) CUDA_TRY(ctx, cudaEventRecord(ctx->ev_phase[1], ctx->stream));
    if (!defer_check) {
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->phases[1] = now_ms() - t0;
    }
    return CZK_OK;
}

int czk_groth16_witness_map(czk_ctx* ctx, int scheme, size_t n_sq, const uint64_t* chain_sh, uint64_t* h_out) {
    if (!ctx || !chain_sh || !h_out || !n_sq) return fail(ctx, CZK_ERR_ARG, "czk_groth16_witness_map: argument");
    if (scheme == CZK_SCHEME_PLAIN && ctx->nranks != 1) return fail(ctx, CZK_ERR_ARG, "plain scheme needs a 1-party context");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    unsigned log_d;
    size_t D = domain_size_for(n_sq, &log_d);
    ShareVecs v;
    int rc = witness_map_dev(ctx, scheme, n_sq, log_d, chain_sh, nullptr, v);
    if (rc == CZK_OK) rc = czk_vec_download(ctx, v.a, 0, h_out, D);
    free_share_vecs(ctx, v);
    return rc;
}

#include "group_shares.hpp"

// GroupShare::scale (share/group.rs:70-109) with DummyGroupTripleSource (wire/group.rs:37-75): x = 0, y = 1@king, z = 0
template <class HF, int LIMBS>
static int group_scale_shared(czk_ctx* ctx, int scheme, GShare<HF, LIMBS>& self, const HFr& o_sh, const HFr& o_mac) {
    typedef HPoint<HF> P;
    if (scheme == CZK_SCHEME_PLAIN) {
        uint64_t k[4];
        o_sh.from_mont().to_limbs(k);
        self.sh = P::mul(self.sh, k, 4);
        self.mac = self.sh;
        return CZK_OK;
    }
    P sx;
    CZK_TRY((group_open<HF, LIMBS>(ctx, scheme, self, &sx)));
    HFr y = ctx->rank == 0 ? HFr::one() : HFr::zero();
    HFr oy;
    CZK_TRY(field_open1(ctx, scheme, HFr::add(o_sh, y), HFr::add(o_mac, y), &oy));
    // out = z - scale_pub_group(sx, y) - x.scale_pub_scalar(oy) ; out.shift(sx * oy)
    GShare<HF, LIMBS> out;
    out.sh = P::infinity();
    out.mac = P::infinity();
    if (ctx->rank == 0) {  // y = (1, mac 1) at the king, (0, 0) elsewhere: sx * 0 is the identity
        P neg = sx;
        neg.negate();
        out.sh.add(neg);
        out.mac.add(neg);
        uint64_t k[4];
        oy.from_mont().to_limbs(k);
        P sxoy = P::mul(sx, k, 4);
        out.sh.add(sxoy);   // AdditiveGroupShare::shift: king only
        out.mac.add(sxoy);  // mac += other * mac_share (1 at the king)
    }
    self = out;
    return CZK_OK;
}

// The local half of GroupShare::scale once sx = open(self) and oy = open(other + y) are known (see group_scale_shared):
// out = z - scale_pub_group(sx, y) - x.scale_pub_scalar(oy), then out.shift(sx * oy), with x = 0, y = 1@king, z = 0
template <class HF, int LIMBS>
static GShare<HF, LIMBS> group_scale_finish(const czk_ctx* ctx, const HPoint<HF>& sx, const HFr& oy) {
    typedef HPoint<HF> P;
    GShare<HF, LIMBS> out;
    out.sh = P::infinity();
    out.mac = P::infinity();
    if (ctx->rank == 0) {
        P neg = sx;
        neg.negate();
        out.sh.add(neg);
        out.mac.add(neg);
        uint64_t k[4];
        oy.from_mont().to_limbs(k);
        P sxoy = P::mul(sx, k, 4);
        out.sh.add(sxoy);
        out.mac.add(sxoy);
    }
    return out;
}

template <class HF, int LIMBS>
static void shift_pub(const czk_ctx* ctx, int scheme, GShare<HF, LIMBS>& s, const HPoint<HF>& el) {
    if (scheme == CZK_SCHEME_PLAIN || ctx->rank == 0) {
        s.sh.add(el);
        s.mac.add(el);
    }
}

static int prove_impl(czk_ctx* ctx, int scheme, const czk_pk* pk, const uint64_t* chain_sh, const czk_vec* chain_dev,
                      const uint64_t r_sh[4], const uint64_t s_sh[4], uint64_t proof_sh[48], uint8_t proof_sh_inf[3],
                      uint64_t proof[48], uint8_t proof_inf[3], const czk_r1cs* cs = nullptr, const uint64_t* full_sh = nullptr);

int czk_groth16_prove(czk_ctx* ctx, int scheme, const czk_pk* pk, const uint64_t* chain_sh, const uint64_t r_sh[4],
                      const uint64_t s_sh[4], uint64_t proof_sh[48], uint8_t proof_sh_inf[3], uint64_t proof[48],
                      uint8_t proof_inf[3]) {
    if (!chain_sh) return fail(ctx, CZK_ERR_ARG, "czk_groth16_prove: null argument");
    return prove_impl(ctx, scheme, pk, chain_sh, nullptr, r_sh, s_sh, proof_sh, proof_sh_inf, proof, proof_inf);
}
int czk_groth16_prove_vec(czk_ctx* ctx, int scheme, const czk_pk* pk, const czk_vec* chain_dev, const uint64_t r_sh[4],
                          const uint64_t s_sh[4], uint64_t proof_sh[48], uint8_t proof_sh_inf[3], uint64_t proof[48],
                          uint8_t proof_inf[3]) {
    if (!chain_dev || !pk || chain_dev->n < pk->n_sq + 1) return fail(ctx, CZK_ERR_ARG, "czk_groth16_prove_vec: chain vector");
    return prove_impl(ctx, scheme, pk, nullptr, chain_dev, r_sh, s_sh, proof_sh, proof_sh_inf, proof, proof_inf);
}

int czk_groth16_prove_r1cs(czk_ctx* ctx, int scheme, const czk_pk* pk, const czk_r1cs* cs, const uint64_t* full_sh,
                           const uint64_t r_sh[4], const uint64_t s_sh[4], uint64_t proof_sh[48], uint8_t proof_sh_inf[3],
                           uint64_t proof[48], uint8_t proof_inf[3]) {
    if (!cs || !full_sh) return fail(ctx, CZK_ERR_ARG, "czk_groth16_prove_r1cs: null argument");
    return prove_impl(ctx, scheme, pk, nullptr, nullptr, r_sh, s_sh, proof_sh, proof_sh_inf, proof, proof_inf, cs, full_sh);
}

int czk_squaring_chain(const uint64_t start[4], size_t n_sq, uint64_t* out) {
    if (!start || !out) return fail(nullptr, CZK_ERR_ARG, "czk_squaring_chain: null argument");
    HFr x = HFr::from_limbs(start);
    x.to_limbs(out);
    for (size_t i = 1; i <= n_sq; i++) {
        x = HFr::sqr(x);
        x.to_limbs(out + 4 * i);
    }
    return CZK_OK;
}

int czk_king_share_batch(const uint64_t* values, size_t k, int n_parties, uint64_t seed, uint64_t* out) {
    if (!values || !out || n_parties < 1) return fail(nullptr, CZK_ERR_ARG, "czk_king_share_batch: argument");
    uint64_t st = seed ^ 0x5eedc0deull;
    auto next = [&]() {
        st += 0x9E3779B97F4A7C15ull;
        uint64_t z = st;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    };
    for (size_t i = 0; i < k; i++) {
        HFr rest = HFr::from_limbs(values + 4 * i);
        for (int p = 0; p + 1 < n_parties; p++) {
            HFr r;
            do {  // uniform in [0, r): 253-bit candidates, rejection
                for (int j = 0; j < 4; j++) r.l[j] = next();
                r.l[3] &= (1ull << 61) - 1;
            } while (HFr::geq_mod(r.l));
            r.to_limbs(out + ((size_t)p * k + i) * 4);
            rest = HFr::sub(rest, r);
        }
        rest.to_limbs(out + ((size_t)(n_parties - 1) * k + i) * 4);
    }
    return CZK_OK;
}

// ------------------------------------------------------------------------------------------ GSZ20 group shares (host)
// gsz20/mod.rs:942-1000, :1043-1130 (scale_pub_group, shift, open, king_compute, mult) and :1135-1400 (the group
// ip_compute / ip_compress / ip_check / hadamard_check), for the three shared-scalar x shared-point products of
// create_proof.  O(1) points: host arithmetic, the exchanges are host-staged NCCL all-gathers of single points.
int gsz_coin(czk_ctx* ctx, HFr* out);
int gsz_mult1(czk_ctx* ctx, const HFr& x, const HFr& y, HFr* out);
int czk_gsz_open_scalar_internal(czk_ctx* ctx, const HFr& v, HFr* out);
int czk_gsz_prepare_internal(czk_ctx* ctx);


template <class HF>
static HPoint<HF> pt_scale(const HPoint<HF>& p, const HFr& s) {
    uint64_t k[4];
    s.from_mont().to_limbs(k);
    return HPoint<HF>::mul(p, k, 4);
}
template <class HF>
static void pt_sub(HPoint<HF>& a, const HPoint<HF>& b) {
    HPoint<HF> t = b;
    t.negate();
    a.add(t);
}
// open_degree_vec on group shares with the reference's shares[i] indexing (:1056-1068):
//   coeff_i = sum_j shares[i] * (w^-ij / n) = shares[i] * (n^-1 sum_j w^-ij)
// so the value is the king's share and every higher coefficient is the identity.  Used by open (all-gather, every party
// interpolates) and king_compute (gather to the king, king returns the value): the same points move either way.
template <class HF, int LIMBS>
static int gsz_group_open(czk_ctx* ctx, const HPoint<HF>& share, int degree, HPoint<HF>* out) {
    typedef GShare<HF, LIMBS> GS;
    typedef HPoint<HF> P;
    const int n = ctx->nranks;
    const size_t rec = (2 * LIMBS + 1) * 8;
    std::vector<uint64_t> send(2 * LIMBS + 1), recv((size_t)n * (2 * LIMBS + 1));
    send[2 * LIMBS] = (uint64_t)GS::to_affine_limbs(share, send.data());
    CZK_TRY(czk_net_allgather_host(ctx, send.data(), recv.data(), rec));
    const HFr n_inv = ctx->gsz.n_inv;
    const std::vector<HFr>& winv = ctx->gsz.winv_host;  // w^-k of the share domain (built with the device table)
    for (int i = 0; i < n; i++) {
        HFr sc = HFr::zero();
        for (int j = 0; j < n; j++) sc = HFr::add(sc, winv[(size_t)((i * j) % n)]);
        sc = HFr::mul(sc, n_inv);
        const uint64_t* r = recv.data() + (size_t)i * (2 * LIMBS + 1);
        P sh = GS::from_affine_limbs(r, (int)r[2 * LIMBS]);
        P coeff = sc.is_zero() ? P::infinity() : (sc == HFr::one() ? sh : pt_scale(sh, sc));
        if (i == 0) *out = coeff;
        else if (i > degree && !coeff.is_inf())
            return fail(ctx, CZK_ERR_PROTOCOL, "GSZ group degree check failed (gsz20/mod.rs:1070 assert)");
    }
    return CZK_OK;
}
// mult (:1112-1130): z = king_compute(y * x + r2) - r with r = r2 = identity
static int gsz_g1_mult(czk_ctx* ctx, const HFr& x, const HG1& y, HG1* z) {
    HG1 t = pt_scale(y, x);
    ctx->gsz.king_computes++;
    return gsz_group_open<HFq, 6>(ctx, t, 2 * ctx->gsz.t, z);
}
static int gsz_g1_ip_compute(czk_ctx* ctx, const HFr* xs, const HG1* ys, size_t k, HG1* out) {
    HG1 acc = HG1::infinity();
    for (size_t i = 0; i < k; i++) acc.add(pt_scale(ys[i], xs[i]));
    ctx->gsz.king_computes++;
    return gsz_group_open<HFq, 6>(ctx, acc, 2 * ctx->gsz.t, out);
}
// hadamard_check (:1330-1349) + ip_check (:1262-1328) on (field, group) triples
static int gsz_g1_product_check(czk_ctx* ctx, std::vector<HFr> x, std::vector<HG1> y, std::vector<HG1> z) {
    const size_t k = x.size();
    if (!k) return CZK_OK;
    HFr r, r_i = HFr::one();
    CZK_TRY(gsz_coin(ctx, &r));
    HG1 ip = HG1::infinity();
    for (size_t i = 0; i < k; i++) {
        x[i] = HFr::mul(x[i], r_i);
        z[i] = pt_scale(z[i], r_i);
        ip.add(z[i]);
        r_i = HFr::mul(r_i, r);
    }
    size_t len = k;
    std::vector<HFr> xm(k + 2), x3(k + 2);
    std::vector<HG1> ym(k + 2), y3(k + 2);
    x.resize(k + 2);
    y.resize(k + 2);
    while (len > 1) {
        if (len & 1) {
            x[len] = HFr::zero();
            y[len] = HG1::infinity();
            len++;
        }
        const size_t h = len / 2;
        HG1 ip_l, ip_r, ip3;
        CZK_TRY(gsz_g1_ip_compute(ctx, x.data(), y.data(), h, &ip_l));
        ip_r = ip;
        pt_sub(ip_r, ip_l);
        for (size_t i = 0; i < h; i++) {
            xm[i] = HFr::sub(x[h + i], x[i]);
            x3[i] = HFr::add(x[h + i], xm[i]);
            ym[i] = y[h + i];
            pt_sub(ym[i], y[i]);
            y3[i] = y[h + i];
            y3[i].add(ym[i]);
        }
        CZK_TRY(gsz_g1_ip_compute(ctx, x3.data(), y3.data(), h, &ip3));
        HFr rr;
        CZK_TRY(gsz_coin(ctx, &rr));
        for (size_t i = 0; i < h; i++) {
            x[i] = HFr::add(HFr::mul(xm[i], rr), HFr::sub(x[i], xm[i]));
            HG1 yb = y[i];
            pt_sub(yb, ym[i]);
            HG1 ymr = pt_scale(ym[i], rr);
            ymr.add(yb);
            y[i] = ymr;
        }
        {
            HFr one = HFr::one(), two = HFr::from_u64(2), three = HFr::from_u64(3), inv2 = HFr::inv(two);
            HFr a = HFr::sub(rr, two), b = HFr::sub(rr, three), c = HFr::sub(rr, one);
            HFr f1 = HFr::mul(HFr::mul(a, b), inv2), f2 = HFr::neg(HFr::mul(c, b)), f3 = HFr::mul(HFr::mul(c, a), inv2);
            HG1 s1 = pt_scale(ip_l, f1);
            s1.add(pt_scale(ip_r, f2));
            s1.add(pt_scale(ip3, f3));
            ip = s1;
        }
        len = h;
    }
    HFr one = HFr::one(), ipr, xb, fx;
    CZK_TRY(gsz_mult1(ctx, one, one, &ipr));
    CZK_TRY(gsz_mult1(ctx, x[0], one, &xb));
    HG1 yb, ib, fy, fz;
    CZK_TRY(gsz_g1_mult(ctx, one, y[0], &yb));
    CZK_TRY(gsz_g1_mult(ctx, ipr, ip, &ib));
    CZK_TRY(czk_gsz_open_scalar_internal(ctx, xb, &fx));
    ctx->gsz.opens += 2;
    CZK_TRY((gsz_group_open<HFq, 6>(ctx, yb, ctx->gsz.t, &fy)));
    CZK_TRY((gsz_group_open<HFq, 6>(ctx, ib, ctx->gsz.t, &fz)));
    fx.to_limbs(ctx->gsz_check.group_x);
    ctx->gsz_check.group_inf[0] = (uint8_t)GShare<HFq, 6>::to_affine_limbs(fy, ctx->gsz_check.group_yz);
    ctx->gsz_check.group_inf[1] = (uint8_t)GShare<HFq, 6>::to_affine_limbs(fz, ctx->gsz_check.group_yz + 12);
    uint64_t chk[12];
    int chk_inf = GShare<HFq, 6>::to_affine_limbs(pt_scale(fy, fx), chk);
    if (chk_inf != (int)ctx->gsz_check.group_inf[1] || std::memcmp(chk, ctx->gsz_check.group_yz + 12, sizeof chk) != 0)
        return fail(ctx, CZK_ERR_PROTOCOL, "GSZ group product check failed (gsz20/mod.rs:1326 assert_eq!)");
    return CZK_OK;
}

// r * delta_g1, s * delta_g1 and s * delta_g2 (prover.rs:118-160) depend on the key and on the blinding shares only: a host
// thread computes them while the device runs the witness map and the MSMs (serial, they were half of the 3 ms tail).
struct TailPre {
    HG1 r_delta, s_delta;
    HG2 s_delta_g2;
};
static TailPre tail_precompute(const czk_pk* pk, const uint64_t r_sh[4], const uint64_t s_sh[4]) {
    const HG1 delta_g1 = HG1::from_affine(HFq::from_limbs(pk->vk_g1 + 24), HFq::from_limbs(pk->vk_g1 + 30));
    const HG2 delta_g2 = HG2::from_affine(HFq2::from_limbs(pk->vk_g2 + 48), HFq2::from_limbs(pk->vk_g2 + 60));
    uint64_t rk[4], sk[4];
    fr_canonical(r_sh, rk);
    fr_canonical(s_sh, sk);
    TailPre t;
    t.r_delta = HG1::mul(delta_g1, rk, 4);
    t.s_delta = HG1::mul(delta_g1, sk, 4);
    t.s_delta_g2 = HG2::mul(delta_g2, sk, 4);
    return t;
}

// create_proof's group arithmetic on GSZ shares + pf.reveal(); the first group reveal runs every queued product check
// (:900-905, :1700-1711).  r, s: the value every party holds (the reference's rand() stub gives 1).
static int prove_tail_gsz(czk_ctx* ctx, const czk_pk* pk, const TailPre& pre, const uint64_t r_sh[4], const uint64_t s_sh[4], const HG1& h_acc,
                          const HG1& l_acc, const HG1& a_acc, const HG1& b1_acc, const HG2& b2_acc, uint64_t proof_sh[48],
                          uint8_t proof_sh_inf[3], uint64_t proof[48], uint8_t proof_inf[3]) {
    typedef GShare<HFq, 6> S1;
    typedef GShare<HFq2, 12> S2;
    double t0 = now_ms();
    CZK_TRY(czk_gsz_prepare_internal(ctx));
    HG1 alpha_g1 = HG1::from_affine(HFq::from_limbs(pk->vk_g1), HFq::from_limbs(pk->vk_g1 + 6));
    HG1 beta_g1 = HG1::from_affine(HFq::from_limbs(pk->vk_g1 + 12), HFq::from_limbs(pk->vk_g1 + 18));
    HG2 beta_g2 = HG2::from_affine(HFq2::from_limbs(pk->vk_g2), HFq2::from_limbs(pk->vk_g2 + 12));
    const HFr r = HFr::from_limbs(r_sh), s = HFr::from_limbs(s_sh);
    std::vector<HFr> gx;
    std::vector<HG1> gy, gz;
    // r_s_delta_g1 = (delta * r) * s
    HG1 rsd = pre.r_delta, t;
    gx.push_back(s);
    gy.push_back(rsd);
    CZK_TRY(gsz_g1_mult(ctx, s, rsd, &t));
    rsd = t;
    gz.push_back(rsd);
    // g_a = r*delta + a_query[0] + MSM + alpha   (shift adds public points at every party, :971-974)
    HG1 g_a = pre.r_delta;
    g_a.add(S1::from_affine_limbs(pk->a0, pk->a0_inf));
    g_a.add(a_acc);
    g_a.add(alpha_g1);
    HG1 s_g_a;
    gx.push_back(s);
    gy.push_back(g_a);
    CZK_TRY(gsz_g1_mult(ctx, s, g_a, &s_g_a));
    gz.push_back(s_g_a);
    HG1 g1_b = pre.s_delta;
    g1_b.add(S1::from_affine_limbs(pk->b10, pk->b10_inf));
    g1_b.add(b1_acc);
    g1_b.add(beta_g1);
    HG2 g2_b = pre.s_delta_g2;
    g2_b.add(S2::from_affine_limbs(pk->b20, pk->b20_inf));
    g2_b.add(b2_acc);
    g2_b.add(beta_g2);
    HG1 r_g1_b;
    gx.push_back(r);
    gy.push_back(g1_b);
    CZK_TRY(gsz_g1_mult(ctx, r, g1_b, &r_g1_b));
    gz.push_back(r_g1_b);
    HG1 g_c = s_g_a;
    g_c.add(r_g1_b);
    pt_sub(g_c, rsd);
    g_c.add(l_acc);
    g_c.add(h_acc);
    proof_sh_inf[0] = (uint8_t)S1::to_affine_limbs(g_a, proof_sh);
    proof_sh_inf[1] = (uint8_t)S2::to_affine_limbs(g2_b, proof_sh + 12);
    proof_sh_inf[2] = (uint8_t)S1::to_affine_limbs(g_c, proof_sh + 36);
    // pf.reveal(): queued field products first, then the group products, then the three opens
    CZK_TRY(czk_gsz_check_products(ctx, nullptr));
    CZK_TRY(gsz_g1_product_check(ctx, gx, gy, gz));
    HG1 A, Cc;
    HG2 B;
    ctx->gsz.opens += 3;
    CZK_TRY((gsz_group_open<HFq, 6>(ctx, g_a, ctx->gsz.t, &A)));
    CZK_TRY((gsz_group_open<HFq2, 12>(ctx, g2_b, ctx->gsz.t, &B)));
    CZK_TRY((gsz_group_open<HFq, 6>(ctx, g_c, ctx->gsz.t, &Cc)));
    proof_inf[0] = (uint8_t)S1::to_affine_limbs(A, proof);
    proof_inf[1] = (uint8_t)S2::to_affine_limbs(B, proof + 12);
    proof_inf[2] = (uint8_t)S1::to_affine_limbs(Cc, proof + 36);
    ctx->phases[7] = now_ms() - t0;
    return CZK_OK;
}

int czk_groth16_gsz_last_checks(const czk_ctx* ctx, uint64_t field_xyz[12], uint64_t group_x[4], uint64_t group_yz[24],
                                uint8_t group_inf[2], uint64_t counts[2]) {
    if (!ctx) return CZK_ERR_ARG;
    if (field_xyz) std::memcpy(field_xyz, ctx->gsz.last_check, sizeof ctx->gsz.last_check);
    if (group_x) std::memcpy(group_x, ctx->gsz_check.group_x, sizeof ctx->gsz_check.group_x);
    if (group_yz) std::memcpy(group_yz, ctx->gsz_check.group_yz, sizeof ctx->gsz_check.group_yz);
    if (group_inf) std::memcpy(group_inf, ctx->gsz_check.group_inf, 2);
    if (counts) {
        counts[0] = ctx->gsz.king_computes;
        counts[1] = ctx->gsz.opens;
    }
    return CZK_OK;
}

static int prove_impl(czk_ctx* ctx, int scheme, const czk_pk* pk, const uint64_t* chain_sh, const czk_vec* chain_dev,
                      const uint64_t r_sh[4], const uint64_t s_sh[4], uint64_t proof_sh[48], uint8_t proof_sh_inf[3],
                      uint64_t proof[48], uint8_t proof_inf[3], const czk_r1cs* cs, const uint64_t* full_sh) {
    if (!ctx || !pk || !r_sh || !s_sh || !proof_sh || !proof_sh_inf || !proof || !proof_inf)
        return fail(ctx, CZK_ERR_ARG, "czk_groth16_prove: null argument");
    if (scheme == CZK_SCHEME_PLAIN && ctx->nranks != 1) return fail(ctx, CZK_ERR_ARG, "plain scheme needs a 1-party context");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    typedef GShare<HFq, 6> S1;
    typedef GShare<HFq2, 12> S2;
    const size_t n_sq = pk->n_sq, D = pk->D;
    if (!cs && !n_sq) return fail(ctx, CZK_ERR_ARG, "czk_groth16_prove: this key was uploaded for a general circuit - use czk_groth16_prove_r1cs");
    for (int i = 0; i < 8; i++) ctx->phases[i] = 0;
    std::future<TailPre> tail_pre;
    try {
        tail_pre = std::async(std::launch::async, tail_precompute, pk, r_sh, s_sh);
    } catch (const std::exception&) {  // no thread to be had: the tail computes them itself
    }
    auto tail_pre_get = [&] { return tail_pre.valid() ? tail_pre.get() : tail_precompute(pk, r_sh, s_sh); };
    ShareVecs v;
    if (cs && (cs->ncons != pk->ncons || cs->ninst != pk->ninst || cs->nwit != pk->nwit))
        return fail(ctx, CZK_ERR_ARG, "czk_groth16_prove_r1cs: the proving key was made for a circuit of another shape");
    // ---- the five MSMs (prover.rs:104,108,132,143,155); share-local, no communication.
    // SPDZ computes sh and mac as the same MSM of the value shares (spdz.rs:440-446): done once, used twice.
    // Four of them take only the assignment, so they are enqueued BEFORE the witness map, on the context's two MSM lanes
    // (lane 0: a, b_g1 on a's digit sort, then h; lane 1: b_g2, l): the transforms run on the context stream under their accumulation rounds,
    // and the serial stretches of one MSM (digit sort, finish walk, bucket reduction, host tail) run under another's rounds.
    // Nothing is read back between the kernels of an MSM, so the host only waits when it collects.
    MsmJob job_h, job_l, job_a, job_b1, job_b2;
    auto drain = [&] {  // error path: nothing of this proof may still be in flight when the share vectors are freed
        czk_ctx_sync(ctx);
        for (MsmLane& l : ctx->lanes) l.collected = l.enqueued;
        cudaMemsetAsync(ctx->flag, 0, 4, ctx->stream);  // a verdict nobody read must not leak into the next call
    };
    auto enqueue_witness_msms = [&]() -> int {
        // scalar vectors of the l-query MSM (the witness) and of the a / b-query MSMs (instance[1..] ++ witness)
        const czk_vec* wit_vec = cs ? v.full : v.chain;
        const size_t wit_off = cs ? pk->ninst : 0, n_wit = pk->nwit;
        const czk_vec* asg_vec = cs ? v.full : v.assign;
        const size_t asg_off = cs ? 1 : 0, n_asg = pk->ninst + pk->nwit - 1;
        // a and b_g1 back to back on one lane: same scalars, so b_g1 runs on a's digit sort when the key allows it
        CZK_TRY(msm_bases_enqueue(ctx, 0, pk->q[0], 1, asg_vec, asg_off, 1, n_asg, &job_a));
        CZK_TRY(msm_bases_enqueue(ctx, 1, pk->q[2], 1, asg_vec, asg_off, 1, n_asg, &job_b2));
        CZK_TRY(msm_bases_enqueue(ctx, 0, pk->q[1], 1, asg_vec, asg_off, 1, n_asg, &job_b1, pk->ab_share_plan));
        CZK_TRY(msm_bases_enqueue(ctx, 1, pk->q[4], 0, wit_vec, wit_off, 1, n_wit, &job_l));
        return CZK_OK;
    };
    int rc = witness_map_dev(ctx, scheme, n_sq, pk->log_d, chain_sh, chain_dev, v, cs, full_sh, enqueue_witness_msms, true);
    auto fin = [&](int c) {
        drain();
        free_share_vecs(ctx, v);
        return c;
    };
    if (rc != CZK_OK) return fin(rc);
    if ((rc = msm_bases_enqueue(ctx, 0, pk->q[3], 0, v.a, 0, 1, D - 1, &job_h)) != CZK_OK) return fin(rc);
    uint64_t o1[18], o2[36];
    S1 h_acc, l_acc, a_acc, b1_acc;
    S2 b2_acc;
    // collect in each lane's enqueue order; the phase figures are each job's device time on its own stream (the jobs overlap,
    // so they do not add up to the proof time)
    if ((rc = msm_collect(ctx, &job_a, o1, &ctx->phases[4])) != CZK_OK) return fin(rc);
    a_acc.sh = a_acc.mac = S1::from_jac_out(o1);
    if ((rc = msm_collect(ctx, &job_b2, o2, &ctx->phases[6])) != CZK_OK) return fin(rc);
    b2_acc.sh = b2_acc.mac = S2::from_jac_out(o2);
    if ((rc = msm_collect(ctx, &job_b1, o1, &ctx->phases[5])) != CZK_OK) return fin(rc);
    b1_acc.sh = b1_acc.mac = S1::from_jac_out(o1);
    if ((rc = msm_collect(ctx, &job_l, o1, &ctx->phases[3])) != CZK_OK) return fin(rc);
    l_acc.sh = l_acc.mac = S1::from_jac_out(o1);
    if ((rc = msm_collect(ctx, &job_h, o1, &ctx->phases[2])) != CZK_OK) return fin(rc);
    h_acc.sh = h_acc.mac = S1::from_jac_out(o1);
    // the h MSM waited for the witness map, so the context stream is idle now: the SPDZ MAC verdict of the Beaver product
    // (spdz.rs:182) is read here, once per proof
    if (scheme == CZK_SCHEME_SPDZ && (rc = sh_collect_flags(ctx, "czk_groth16_prove (witness-map product)")) != CZK_OK) return fin(rc);
    {
        float wm = 0;
        if (ctx->ev_phase[0] && cudaEventElapsedTime(&wm, ctx->ev_phase[0], ctx->ev_phase[1]) == cudaSuccess) ctx->phases[1] = wm;
    }
    free_share_vecs(ctx, v);

    if (scheme == CZK_SCHEME_GSZ)
        return prove_tail_gsz(ctx, pk, tail_pre_get(), r_sh, s_sh, h_acc.sh, l_acc.sh, a_acc.sh, b1_acc.sh, b2_acc.sh, proof_sh, proof_sh_inf, proof, proof_inf);
    // ---- O(1) group arithmetic on shares (prover.rs:110-177)
    double t0 = now_ms();
    HG1 alpha_g1 = HG1::from_affine(HFq::from_limbs(pk->vk_g1), HFq::from_limbs(pk->vk_g1 + 6));
    HG1 beta_g1 = HG1::from_affine(HFq::from_limbs(pk->vk_g1 + 12), HFq::from_limbs(pk->vk_g1 + 18));
    HG2 beta_g2 = HG2::from_affine(HFq2::from_limbs(pk->vk_g2), HFq2::from_limbs(pk->vk_g2 + 12));
    HFr r = HFr::from_limbs(r_sh), s = HFr::from_limbs(s_sh);
    const TailPre pre = tail_pre_get();
    // from_add_shared scalars: mac = share (key 1), so scale_pub_group gives sh == mac (spdz.rs:419-423)
    S1 rsd;
    rsd.sh = rsd.mac = pre.r_delta;  // delta_g1 * r
    // A = r*delta + a_query[0] + MSM + alpha   (calculate_coeff, prover.rs:216-232)
    S1 g_a;
    g_a.sh = g_a.mac = pre.r_delta;
    shift_pub(ctx, scheme, g_a, S1::from_affine_limbs(pk->a0, pk->a0_inf));
    g_a.sh.add(a_acc.sh);
    g_a.mac.add(a_acc.mac);
    shift_pub(ctx, scheme, g_a, alpha_g1);
    S1 g1_b;
    g1_b.sh = g1_b.mac = pre.s_delta;
    shift_pub(ctx, scheme, g1_b, S1::from_affine_limbs(pk->b10, pk->b10_inf));
    g1_b.sh.add(b1_acc.sh);
    g1_b.mac.add(b1_acc.mac);
    shift_pub(ctx, scheme, g1_b, beta_g1);
    S2 g2_b;
    g2_b.sh = g2_b.mac = pre.s_delta_g2;
    shift_pub(ctx, scheme, g2_b, S2::from_affine_limbs(pk->b20, pk->b20_inf));
    g2_b.sh.add(b2_acc.sh);
    g2_b.mac.add(b2_acc.mac);
    shift_pub(ctx, scheme, g2_b, beta_g2);
    // The three shared-scalar x shared-point products (r s delta, s A, r B1: GroupShare::scale, share/group.rs:70-109) are
    // independent: their six openings (the point and the masked scalar of each) travel in one exchange, their MAC checks in
    // another (group_shares.hpp, open_many) - the reference's 12 broadcasts, counted as such, in 2 all-gathers.
    S1 s_g_a, r_g1_b;
    if (scheme == CZK_SCHEME_PLAIN) {
        s_g_a = g_a;
        r_g1_b = g1_b;
        CZK_TRY((group_scale_shared<HFq, 6>(ctx, scheme, rsd, s, s)));
        CZK_TRY((group_scale_shared<HFq, 6>(ctx, scheme, s_g_a, s, s)));
        CZK_TRY((group_scale_shared<HFq, 6>(ctx, scheme, r_g1_b, r, r)));
    } else {
        const HFr y = ctx->rank == 0 ? HFr::one() : HFr::zero();
        std::vector<OpenItem> it = {OpenItem::point(rsd),  OpenItem::field(HFr::add(s, y), HFr::add(s, y)),
                                    OpenItem::point(g_a),  OpenItem::field(HFr::add(s, y), HFr::add(s, y)),
                                    OpenItem::point(g1_b), OpenItem::field(HFr::add(r, y), HFr::add(r, y))};
        CZK_TRY(open_many(ctx, scheme, it));
        rsd = group_scale_finish<HFq, 6>(ctx, it[0].g1_out, it[1].f_out);
        s_g_a = group_scale_finish<HFq, 6>(ctx, it[2].g1_out, it[3].f_out);
        r_g1_b = group_scale_finish<HFq, 6>(ctx, it[4].g1_out, it[5].f_out);
    }
    // C = s*A + r*B1 - r*s*delta + L + H
    S1 g_c = s_g_a;
    g_c.sh.add(r_g1_b.sh);
    g_c.mac.add(r_g1_b.mac);
    {
        HG1 t = rsd.sh, u = rsd.mac;
        t.negate();
        u.negate();
        g_c.sh.add(t);
        g_c.mac.add(u);
    }
    g_c.sh.add(l_acc.sh);
    g_c.mac.add(l_acc.mac);
    g_c.sh.add(h_acc.sh);
    g_c.mac.add(h_acc.mac);
    proof_sh_inf[0] = (uint8_t)S1::to_affine_limbs(g_a.sh, proof_sh);
    proof_sh_inf[1] = (uint8_t)S2::to_affine_limbs(g2_b.sh, proof_sh + 12);
    proof_sh_inf[2] = (uint8_t)S1::to_affine_limbs(g_c.sh, proof_sh + 36);
    // pf.reveal() (groth16/src/reveal.rs:7-12)
    HG1 A, Cc;
    HG2 B;
    {
        std::vector<OpenItem> it = {OpenItem::point(g_a), OpenItem::point(g2_b), OpenItem::point(g_c)};
        CZK_TRY(open_many(ctx, scheme, it));  // the three reveals in one exchange (+ one for the MAC checks)
        A = it[0].g1_out;
        B = it[1].g2_out;
        Cc = it[2].g1_out;
    }
    proof_inf[0] = (uint8_t)S1::to_affine_limbs(A, proof);
    proof_inf[1] = (uint8_t)S2::to_affine_limbs(B, proof + 12);
    proof_inf[2] = (uint8_t)S1::to_affine_limbs(Cc, proof + 36);
    ctx->phases[7] = now_ms() - t0;
    return CZK_OK;
}

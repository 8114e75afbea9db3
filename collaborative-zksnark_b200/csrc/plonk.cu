// The wiring argument of the collaborative Plonk prover on the device (include/czk_plonk.h).
//
// Host-side orchestration in C++ of mpc-plonk/src/lib.rs:199-258 (prove_wiring), :110-197 (prove_unit_product) and
// :343-400 (eval, commit) over additive / SPDZ shares with KZG10 commitments: the same sequence of domain transforms,
// share protocols (batch division, prefix products, Beaver products), commitment MSMs and opening MSMs, each one a call
// into the leaves of czk.h.  What differs from the reference is where things run and how they overlap:
//   * every linear map on a shared polynomial is ONE batched grid over its value and MAC vectors (and over the
//     independent polynomials of a step), czk_ntt_vec_batch;
//   * the nine opening MSMs do not feed the transcript, so they are enqueued on the context's two MSM lanes and collected
//     at the end, under the remaining transforms; the four commitment MSMs do feed it and are waited for.
// The transcript is the caller's (czk_plonk_transcript): absorb / challenge are called exactly where the reference calls
// fs_rng.absorb / fs_rng.gen.
#include <chrono>
#include <vector>

#include "../../include/czk_plonk.h"
#include "fr_ops.cuh"
#include "group_shares.hpp"

namespace {

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// ------------------------------------------------------------------------------------------ the stand-in transcript
uint64_t standin_mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
void standin_absorb(void* user, const uint64_t xy[12], int inf) {
    uint64_t& s = *static_cast<uint64_t*>(user);
    for (int i = 0; i < 12; i++) s = standin_mix(s ^ xy[i]);
    s = standin_mix(s ^ (uint64_t)(inf != 0));
}
void standin_challenge(void* user, uint64_t out[4]) {
    uint64_t& s = *static_cast<uint64_t*>(user);
    for (int j = 0; j < 4; j++) {
        s = standin_mix(s + (uint64_t)j + 1);
        out[j] = s;
    }
    out[3] &= (1ull << 60) - 1;  // < 2^252 < r: any such limbs are a valid Montgomery element
}

// ------------------------------------------------------------------------------------------ shared vectors
struct SVec {  // one shared polynomial / evaluation vector: value shares and, under SPDZ, MAC shares
    czk_vec* sh = nullptr;
    czk_vec* mac = nullptr;
};
struct Prover {
    czk_ctx* ctx;
    int scheme;
    bool spdz, shared;
    const czk_bases* powers;
    unsigned log_d;  // D = 2^log_d, or 3 * 2^log_d when mixed
    bool mixed;
    size_t D;
    const czk_plonk_transcript* tr;
    std::vector<czk_vec*> owned;
    struct Pending {
        MsmJob job;
        int slot;
        bool is_public;
        HFr y_sh, y_mac;  // this party's share of the evaluation (publicized with the proofs, in finish_evals)
    };
    std::vector<Pending> pending;
    typedef GShare<HFq, 6> S1;

    ~Prover() {
        bool in_flight = false;
        for (Pending& pd : pending) in_flight |= pd.job.lane >= 0;
        for (PendingCommit& pc : commits) in_flight |= pc.job.lane >= 0;
        if (in_flight) {  // error path: nothing of this call may still be running when its vectors are freed
            czk_ctx_sync(ctx);
            for (MsmLane& l : ctx->lanes) l.collected = l.enqueued;
        }
        for (czk_vec* v : owned) czk_vec_free(ctx, v);
    }
    int vec(czk_vec** out) {
        CZK_TRY(czk_vec_alloc(ctx, D, out));
        owned.push_back(*out);
        return CZK_OK;
    }
    int alloc(SVec& v) {
        CZK_TRY(vec(&v.sh));
        if (spdz) CZK_TRY(vec(&v.mac));
        return CZK_OK;
    }
    int copy(SVec& dst, const SVec& src) {
        if (!dst.sh) CZK_TRY(alloc(dst));
        CZK_TRY(czk_vec_copy(ctx, dst.sh, 0, src.sh, 0, D));
        if (spdz) CZK_TRY(czk_vec_copy(ctx, dst.mac, 0, src.mac, 0, D));
        return CZK_OK;
    }
    // one batched transform over every component of the listed shared vectors (+ optional public vectors)
    int ntt(std::initializer_list<SVec*> vs, int op, std::initializer_list<czk_vec*> pubs = {}) {
        std::vector<czk_vec*> all;
        for (SVec* v : vs) {
            all.push_back(v->sh);
            if (spdz) all.push_back(v->mac);
        }
        for (czk_vec* p : pubs) all.push_back(p);
        if (mixed) return czk_ntt_mixed_vec_batch(ctx, all.data(), (int)all.size(), log_d, op);
        return czk_ntt_vec_batch(ctx, all.data(), (int)all.size(), log_d, op);
    }
    int distribute_powers(SVec& v, const HFr& g) {
        HFr one = HFr::one();
        CZK_TRY(czk_vec_distribute_powers(ctx, v.sh, g.l, one.l, D));
        if (spdz) CZK_TRY(czk_vec_distribute_powers(ctx, v.mac, g.l, one.l, D));
        return CZK_OK;
    }
    // v += public vector: AdditiveFieldShare::shift adds at the king; the SPDZ MAC gets public * mac_share (1 at the king)
    int shift_pub(SVec& v, const czk_vec* pub) {
        if (shared && ctx->rank != 0) return CZK_OK;
        CZK_TRY(czk_vec_add(ctx, v.sh, pub, D));
        if (spdz) CZK_TRY(czk_vec_add(ctx, v.mac, pub, D));
        return CZK_OK;
    }
    int sub(SVec& a, const SVec& b) {
        CZK_TRY(czk_vec_sub(ctx, a.sh, b.sh, D));
        if (spdz) CZK_TRY(czk_vec_sub(ctx, a.mac, b.mac, D));
        return CZK_OK;
    }
    int div_vanishing(SVec& a) {  // evals *= 1 / (g^D - 1)   (domain/mod.rs:184-191, any multiplicative-subgroup domain)
        const HFr zi = HFr::inv(HFr::sub(HFr::pow_u64(HFr::from_limbs(FrParams::GENERATOR_64), (uint64_t)D), HFr::one()));
        CZK_TRY(czk_vec_scale(ctx, a.sh, zi.l, D));
        if (spdz) CZK_TRY(czk_vec_scale(ctx, a.mac, zi.l, D));
        return CZK_OK;
    }
    int mul(SVec& a, const SVec& b) { return czk_beaver_batch_mul(ctx, scheme, a.sh, a.mac, b.sh, b.mac, D); }

    // commit (lib.rs:367-400): MSM of this party's coefficient shares, publicize (group open), absorb.  Split in two so that
    // a commitment nothing waits for yet (l1, t: no challenge is drawn before the q commitment) runs on an MSM lane under
    // the transforms and share protocols that follow; commit_finish keeps the reference's absorb order.
    struct PendingCommit {
        MsmJob job;
        uint64_t* cmt_xy;
        uint8_t* cmt_inf;
    };
    std::vector<PendingCommit> commits;
    int commit_enqueue(const SVec& poly, int lane, uint64_t cmt_xy[12], uint8_t* cmt_inf) {
        PendingCommit pc;
        pc.cmt_xy = cmt_xy;
        pc.cmt_inf = cmt_inf;
        pc.job.lane = -1;
        if (ctx->lanes[lane].enqueued - ctx->lanes[lane].collected >= CZK_MSM_SLOTS) return fail(ctx, CZK_ERR_ARG, "plonk: commitment queue full");
        CZK_TRY(msm_bases_enqueue(ctx, lane, powers, 0, poly.sh, 0, 1, D, &pc.job));
        commits.push_back(pc);
        return CZK_OK;
    }
    int commit_finish() {  // in enqueue order = the order the reference commits in
        for (PendingCommit& pc : commits) {
            uint64_t o[18];
            CZK_TRY(msm_collect(ctx, &pc.job, o, nullptr));
            pc.job.lane = -1;
            S1 s;
            s.sh = s.mac = S1::from_jac_out(o);  // spdz.rs:440-446: both MSMs run on the value shares
            HG1 opened;
            CZK_TRY((group_open<HFq, 6>(ctx, scheme, s, &opened)));
            *pc.cmt_inf = (uint8_t)S1::to_affine_limbs(opened, pc.cmt_xy);
            tr->absorb_g1(tr->user, pc.cmt_xy, *pc.cmt_inf);
        }
        commits.clear();
        return CZK_OK;
    }
    int commit(const SVec& poly, uint64_t cmt_xy[12], uint8_t* cmt_inf) {
        CZK_TRY(commit_enqueue(poly, 0, cmt_xy, cmt_inf));
        return commit_finish();
    }

    // eval (lib.rs:343-365): KZG10 open at x.  The witness polynomial's MSM is only ENQUEUED (lane = slot & 1); the value is
    // publicized now.  finish_evals() collects the proofs and reveals them (reveal.rs).
    int eval(const SVec& poly, bool is_public, const HFr& x, int slot, czk_plonk_wiring_proof* out) {
        czk_vec* q;
        CZK_TRY(vec(&q));
        uint64_t ev[4], evm[4] = {0, 0, 0, 0};
        CZK_TRY(czk_poly_div_linear(ctx, poly.sh, D, x.l, q, ev));
        if (spdz && !is_public) CZK_TRY(czk_poly_div_linear(ctx, poly.mac, D, x.l, nullptr, evm));
        Pending pd;
        pd.slot = slot;
        pd.is_public = is_public;
        // at most CZK_MSM_SLOTS jobs wait on a lane: 9 openings over 2 lanes = 5 + 4; drain the older half when full
        const int lane = slot & 1;
        if (ctx->lanes[lane].enqueued - ctx->lanes[lane].collected >= CZK_MSM_SLOTS) return fail(ctx, CZK_ERR_ARG, "plonk: opening queue full");
        CZK_TRY(msm_bases_enqueue(ctx, lane, powers, 0, q, 0, 1, D > 1 ? D - 1 : 0, &pd.job));
        pd.y_sh = HFr::from_limbs(ev);
        pd.y_mac = spdz ? HFr::from_limbs(evm) : pd.y_sh;
        pending.push_back(pd);
        if (is_public) pd.y_sh.to_limbs(out->open_val[slot]);
        return CZK_OK;
    }
    // Collect the opening MSMs enqueued since `first`, then publicize the evaluations (y.publicize(), lib.rs:360-362) and
    // reveal the proofs (reveal.rs) of all of them in ONE exchange (+ one for the SPDZ MAC checks): group_shares.hpp, open_many.
    int finish_evals(czk_plonk_wiring_proof* share, czk_plonk_wiring_proof* out, size_t first) {
        std::vector<OpenItem> items;
        std::vector<size_t> who;
        for (size_t i = first; i < pending.size(); i++) {
            Pending& pd = pending[i];
            uint64_t o[18];
            CZK_TRY(msm_collect(ctx, &pd.job, o, nullptr));
            pd.job.lane = -1;
            S1 s;
            s.sh = s.mac = S1::from_jac_out(o);
            share->open_pf_inf[pd.slot] = (uint8_t)S1::to_affine_limbs(s.sh, share->open_pf_xy[pd.slot]);
            if (pd.is_public) {
                out->open_pf_inf[pd.slot] = share->open_pf_inf[pd.slot];
                std::memcpy(out->open_pf_xy[pd.slot], share->open_pf_xy[pd.slot], sizeof out->open_pf_xy[pd.slot]);
                continue;
            }
            items.push_back(OpenItem::field(pd.y_sh, pd.y_mac));
            items.push_back(OpenItem::point(s));
            who.push_back(i);
        }
        CZK_TRY(open_many(ctx, scheme, items));
        for (size_t k = 0; k < who.size(); k++) {
            const Pending& pd = pending[who[k]];
            items[2 * k].f_out.to_limbs(out->open_val[pd.slot]);
            out->open_pf_inf[pd.slot] = (uint8_t)S1::to_affine_limbs(items[2 * k + 1].g1_out, out->open_pf_xy[pd.slot]);
        }
        return CZK_OK;
    }
};

}  // namespace

void czk_plonk_standin_transcript(uint64_t* state, uint64_t seed, czk_plonk_transcript* out) {
    *state = standin_mix(seed);
    out->user = state;
    out->absorb_g1 = standin_absorb;
    out->challenge = standin_challenge;
}

static int prove_wiring_impl(czk_ctx* ctx, int scheme, const czk_bases* powers, unsigned log_d, bool mixed, const czk_vec* p_sh,
                             const czk_vec* p_mac, const czk_vec* w_pub, const czk_plonk_transcript* transcript,
                             czk_plonk_wiring_proof* out_share, czk_plonk_wiring_proof* out, double* phases_ms) {
    if (!ctx || !powers || !p_sh || !w_pub || !transcript || !transcript->absorb_g1 || !transcript->challenge || !out_share || !out)
        return fail(ctx, CZK_ERR_ARG, "czk_plonk_prove_wiring: null argument");
    if (log_d > (mixed ? 26u : 28u)) return fail(ctx, CZK_ERR_ARG, "czk_plonk_prove_wiring: domain too large");
    const size_t D = (size_t)(mixed ? 3 : 1) << log_d;
    const bool spdz = scheme == CZK_SCHEME_SPDZ;
    if (scheme != CZK_SCHEME_PLAIN && scheme != CZK_SCHEME_ADDITIVE && !spdz) return fail(ctx, CZK_ERR_ARG, "czk_plonk_prove_wiring: scheme");
    if (scheme == CZK_SCHEME_PLAIN && ctx->nranks != 1) return fail(ctx, CZK_ERR_ARG, "plain scheme needs a 1-party context");
    if (czk_vec_len(p_sh) < D || czk_vec_len(w_pub) < D || (spdz && (!p_mac || czk_vec_len(p_mac) < D)) || czk_bases_len(powers) < D)
        return fail(ctx, CZK_ERR_ARG, "czk_plonk_prove_wiring: vector or committer key shorter than the domain");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    std::memset(out, 0, sizeof *out);
    std::memset(out_share, 0, sizeof *out_share);
    Prover P{ctx, scheme, spdz, scheme != CZK_SCHEME_PLAIN, powers, log_d, mixed, D, transcript, {}, {}};
    uint64_t dom[4][4];
    if (mixed) CZK_TRY(czk_mixed_domain_params(log_d, dom[0], dom[1], dom[2], dom[3]));
    else CZK_TRY(czk_domain_params(log_d, dom[0], dom[1], dom[2], dom[3]));
    const HFr w = HFr::from_limbs(dom[0]), w_inv = HFr::from_limbs(dom[1]);
    double t_commit = 0, t_open = 0, t_reveal = 0, t_all = now_ms();
    auto timed = [&](double& acc, int rc, double t0) {
        acc += now_ms() - t0;
        return rc;
    };
    auto draw = [&](int which) {
        transcript->challenge(transcript->user, out->challenges[which]);
        return HFr::from_limbs(out->challenges[which]);
    };
    const HFr y = draw(0), z = draw(1);
    SVec p{const_cast<czk_vec*>(p_sh), const_cast<czk_vec*>(p_mac)};  // read only
    SVec p_evals, num, den, l1, l1e, t, tc, fe, tw, nv, dv, dtmp;
    czk_vec *w_evals, *yxz;
    CZK_TRY(P.vec(&w_evals));
    CZK_TRY(P.vec(&yxz));
    // p_evals, w_evals and (z + y X) over the domain: one batched forward transform
    CZK_TRY(P.copy(p_evals, p));
    CZK_TRY(czk_vec_copy(ctx, w_evals, 0, w_pub, 0, D));
    {
        uint64_t c01[8];
        HFr c0 = z, c1 = y;
        if (D == 1) c0 = HFr::add(z, y);  // one-point domain: X = 1
        c0.to_limbs(c01);
        c1.to_limbs(c01 + 4);
        CZK_TRY(czk_vec_upload(ctx, yxz, 0, c01, D > 1 ? 2 : 1));
    }
    CZK_TRY(P.ntt({&p_evals}, CZK_NTT_FFT, {w_evals, yxz}));
    // num = p_evals + w_evals * y + z ; den = p_evals + yx_z_evals   (public parts added at the king)
    CZK_TRY(czk_vec_scale(ctx, w_evals, y.l, D));
    CUDA_TRY(ctx, fr_add_const((uint32_t*)w_evals->d, (const uint32_t*)w_evals->d, z.l, D, ctx->stream));
    CZK_TRY(P.copy(num, p_evals));
    CZK_TRY(P.copy(den, p_evals));
    CZK_TRY(P.shift_pub(num, w_evals));
    CZK_TRY(P.shift_pub(den, yxz));
    // l1_evals = num / den (batch_division_in_place consumes its divisor) ; l1 = interpolate
    CZK_TRY(P.copy(l1e, num));
    CZK_TRY(P.copy(dtmp, den));
    CZK_TRY(czk_share_batch_div(ctx, scheme, l1e.sh, l1e.mac, dtmp.sh, dtmp.mac, D));
    CZK_TRY(P.copy(l1, l1e));
    CZK_TRY(P.ntt({&l1}, CZK_NTT_IFFT));
    double t0 = now_ms();
    CZK_TRY(timed(t_commit, P.commit_enqueue(l1, 0, out->cmt_xy[0], &out->cmt_inf[0]), t0));  // absorbed before t and q, below
    // ---- prove_unit_product(f = l1)
    CZK_TRY(P.copy(t, l1));
    CZK_TRY(P.ntt({&t}, CZK_NTT_FFT));  // f.evaluate_over_domain_by_ref
    CZK_TRY(czk_share_partial_products(ctx, scheme, t.sh, t.mac, D));
    CZK_TRY(P.ntt({&t}, CZK_NTT_IFFT));
    t0 = now_ms();
    CZK_TRY(timed(t_commit, P.commit_enqueue(t, 1, out->cmt_xy[1], &out->cmt_inf[1]), t0));
    CZK_TRY(P.copy(fe, l1));
    CZK_TRY(P.distribute_powers(fe, w));
    CZK_TRY(P.copy(tc, t));
    CZK_TRY(P.copy(tw, t));
    CZK_TRY(P.distribute_powers(tw, w));
    CZK_TRY(P.ntt({&fe, &tc, &tw}, CZK_NTT_COSET_FFT));  // f(wX), t(X), t(wX) over the coset
    CZK_TRY(P.mul(fe, tc));                              // f(wX) t(X)
    CZK_TRY(P.sub(tw, fe));
    CZK_TRY(P.div_vanishing(tw));
    CZK_TRY(P.ntt({&tw}, CZK_NTT_COSET_IFFT));  // q
    t0 = now_ms();
    CZK_TRY(timed(t_commit, P.commit(tw, out->cmt_xy[2], &out->cmt_inf[2]), t0));  // collects l1, t, q in that order
    const HFr r = draw(2);
    const HFr wr = HFr::mul(w, r);
    t0 = now_ms();
    CZK_TRY(P.eval(t, false, wr, 0, out));
    CZK_TRY(P.eval(t, false, r, 1, out));
    CZK_TRY(P.eval(t, false, w_inv, 2, out));  // domain.element(k - 1) = w^(k-1)
    CZK_TRY(P.eval(l1, false, wr, 3, out));
    CZK_TRY(P.eval(tw, false, r, 4, out));
    t_open += now_ms() - t0;
    // ---- l2_q: (l1 * den - num) / Z_H on the coset
    CZK_TRY(P.copy(fe, l1));
    CZK_TRY(P.copy(nv, num));
    CZK_TRY(P.copy(dv, den));
    CZK_TRY(P.ntt({&nv, &dv}, CZK_NTT_IFFT_COSET_FFT));  // interpolate, then over the coset
    CZK_TRY(P.ntt({&fe}, CZK_NTT_COSET_FFT));
    // the five openings of the unit-product argument are still in flight on the MSM lanes: collect them before the lanes
    // are needed again (they ran under the transforms above)
    t0 = now_ms();
    CZK_TRY(timed(t_reveal, P.finish_evals(out_share, out, 0), t0));
    const size_t done = P.pending.size();
    CZK_TRY(P.mul(fe, dv));
    CZK_TRY(P.sub(fe, nv));
    CZK_TRY(P.div_vanishing(fe));
    CZK_TRY(P.ntt({&fe}, CZK_NTT_COSET_IFFT));  // l2_q
    t0 = now_ms();
    CZK_TRY(timed(t_commit, P.commit(fe, out->cmt_xy[3], &out->cmt_inf[3]), t0));
    const HFr x = draw(3);
    t0 = now_ms();
    SVec wpub{const_cast<czk_vec*>(w_pub), const_cast<czk_vec*>(w_pub)};
    CZK_TRY(P.eval(fe, false, x, 5, out));
    CZK_TRY(P.eval(wpub, true, x, 6, out));
    CZK_TRY(P.eval(l1, false, x, 7, out));
    CZK_TRY(P.eval(p, false, x, 8, out));
    t_open += now_ms() - t0;
    t0 = now_ms();
    CZK_TRY(timed(t_reveal, P.finish_evals(out_share, out, done), t0));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    std::memcpy(out_share->cmt_xy, out->cmt_xy, sizeof out->cmt_xy);
    std::memcpy(out_share->cmt_inf, out->cmt_inf, sizeof out->cmt_inf);
    std::memcpy(out_share->open_val, out->open_val, sizeof out->open_val);
    std::memcpy(out_share->challenges, out->challenges, sizeof out->challenges);
    if (phases_ms) {
        const double total = now_ms() - t_all;
        phases_ms[0] = total - t_commit - t_open - t_reveal;
        phases_ms[1] = t_commit;
        phases_ms[2] = t_open;
        phases_ms[3] = t_reveal;
    }
    return CZK_OK;
}

int czk_plonk_prove_wiring(czk_ctx* ctx, int scheme, const czk_bases* powers, unsigned log_d, const czk_vec* p_sh, const czk_vec* p_mac,
                           const czk_vec* w_pub, const czk_plonk_transcript* transcript, czk_plonk_wiring_proof* out_share,
                           czk_plonk_wiring_proof* out, double* phases_ms) {
    return prove_wiring_impl(ctx, scheme, powers, log_d, false, p_sh, p_mac, w_pub, transcript, out_share, out, phases_ms);
}
// the reference's own wire domain: MixedRadixEvaluationDomain::new(3 * n_gates) = 3 * 2^log_m points (relations/flat.rs:282-300)
int czk_plonk_prove_wiring_mixed(czk_ctx* ctx, int scheme, const czk_bases* powers, unsigned log_m, const czk_vec* p_sh, const czk_vec* p_mac,
                                 const czk_vec* w_pub, const czk_plonk_transcript* transcript, czk_plonk_wiring_proof* out_share,
                                 czk_plonk_wiring_proof* out, double* phases_ms) {
    return prove_wiring_impl(ctx, scheme, powers, log_m, true, p_sh, p_mac, w_pub, transcript, out_share, out, phases_ms);
}

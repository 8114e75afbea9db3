// Additive / SPDZ shares of group elements on the host: the O(1) group arithmetic around the MSMs (the prover loops'
// tails) and the openings of single shared points and field elements.
//   mpc-algebra/src/share/add.rs:178-180, spdz.rs:262-275   group open (reveal) with the SPDZ MAC check
//   mpc-algebra/src/share/add.rs:121-125, spdz.rs:119-131   open of one field element
// The exchanges are host-staged all-gathers of single records (czk_net_allgather_host).
#pragma once
#include "ctx.hpp"

// ------------------------------------------------------------------------------------------ group shares on the host
template <class HF, int LIMBS>
struct GShare {
    typedef HPoint<HF> P;
    P sh, mac;

    static P from_jac_out(const uint64_t* xyz) {  // (x, y, 1) or (1, 1, 0) as written by the MSM entry points
        HF z = HF::from_limbs(xyz + 2 * LIMBS);
        if (z.is_zero()) return P::infinity();
        return P::from_affine(HF::from_limbs(xyz), HF::from_limbs(xyz + LIMBS));
    }
    static P from_affine_limbs(const uint64_t* xy, int inf) {
        if (inf) return P::infinity();
        return P::from_affine(HF::from_limbs(xy), HF::from_limbs(xy + LIMBS));
    }
    static int to_affine_limbs(const P& p, uint64_t* xy) {  // returns the infinity flag; infinity is written as (0, 1)
        HF ax, ay;
        if (!p.to_affine(ax, ay)) {
            HF::zero().to_limbs(xy);
            HF::one().to_limbs(xy + LIMBS);
            return 1;
        }
        ax.to_limbs(xy);
        ay.to_limbs(xy + LIMBS);
        return 0;
    }
};

static inline void fr_canonical(const uint64_t mont[4], uint64_t out[4]) { HFr::from_limbs(mont).from_mont().to_limbs(out); }

// open one shared group element (add.rs:178-180 / spdz.rs:262-275): returns the sum of all parties' sh
template <class HF, int LIMBS>
static int group_open(czk_ctx* ctx, int scheme, const GShare<HF, LIMBS>& s, HPoint<HF>* out) {
    typedef GShare<HF, LIMBS> GS;
    typedef HPoint<HF> P;
    if (scheme == CZK_SCHEME_PLAIN) {
        *out = s.sh;
        return CZK_OK;
    }
    const int n = ctx->nranks;
    const size_t rec = (2 * LIMBS + 1) * 8;  // x | y | inf word
    std::vector<uint64_t> send(2 * LIMBS + 1), recv((size_t)n * (2 * LIMBS + 1));
    send[2 * LIMBS] = (uint64_t)GS::to_affine_limbs(s.sh, send.data());
    CZK_TRY(czk_net_allgather_host(ctx, send.data(), recv.data(), rec));
    P x = P::infinity();
    for (int p = 0; p < n; p++) {
        const uint64_t* r = recv.data() + (size_t)p * (2 * LIMBS + 1);
        x.add(GS::from_affine_limbs(r, (int)r[2 * LIMBS]));
    }
    if (scheme == CZK_SCHEME_SPDZ) {
        // dx_t = x * mac_share - mac ; all dx_t must sum to zero (Pragmatic MPC 6.6.2)
        P dx = ctx->rank == 0 ? x : P::infinity();
        P m = s.mac;
        m.negate();
        dx.add(m);
        send[2 * LIMBS] = (uint64_t)GS::to_affine_limbs(dx, send.data());
        CZK_TRY(czk_net_allgather_host(ctx, send.data(), recv.data(), rec));
        P sum = P::infinity();
        for (int p = 0; p < n; p++) {
            const uint64_t* r = recv.data() + (size_t)p * (2 * LIMBS + 1);
            sum.add(GS::from_affine_limbs(r, (int)r[2 * LIMBS]));
        }
        if (!sum.is_inf()) return fail(ctx, CZK_ERR_PROTOCOL, "SPDZ group MAC check failed (spdz.rs:273 assert!(sum.is_zero()))");
    }
    *out = x;
    return CZK_OK;
}

// open one shared field element given as (sh, mac)
static inline int field_open1(czk_ctx* ctx, int scheme, const HFr& sh, const HFr& mac, HFr* out) {
    if (scheme == CZK_SCHEME_PLAIN) {
        *out = sh;
        return CZK_OK;
    }
    const int n = ctx->nranks;
    std::vector<uint64_t> recv((size_t)n * 4);
    CZK_TRY(czk_net_allgather_host(ctx, sh.l, recv.data(), 32));
    HFr x = HFr::zero();
    for (int p = 0; p < n; p++) x = HFr::add(x, HFr::from_limbs(recv.data() + 4 * p));
    if (scheme == CZK_SCHEME_SPDZ) {
        HFr ms = ctx->rank == 0 ? HFr::one() : HFr::zero();
        HFr dx = HFr::sub(HFr::mul(ms, x), mac);
        CZK_TRY(czk_net_allgather_host(ctx, dx.l, recv.data(), 32));
        HFr sum = HFr::zero();
        for (int p = 0; p < n; p++) sum = HFr::add(sum, HFr::from_limbs(recv.data() + 4 * p));
        if (!sum.is_zero()) return fail(ctx, CZK_ERR_PROTOCOL, "SPDZ MAC check failed (spdz.rs:129 assert!(sum.is_zero()))");
    }
    *out = x;
    return CZK_OK;
}


// ------------------------------------------------------------------------------------------ several openings, one exchange
// The prover tails open a handful of independent values (three shared points and three shared scalars before the scale
// products, then the three proof elements): their records travel together - one all-gather for the values and one for the
// SPDZ sigma records instead of two per item.  Every item is still checked on its own; mpc-net's accounting (one broadcast
// per opened item and per MAC check, multi.rs:145-174) is kept by counting the items, the bytes are the same bytes.
struct OpenItem {
    int kind = 0;  // 0: field element, 1: G1 point, 2: G2 point
    HFr f_sh, f_mac, f_out;
    GShare<HFq, 6> g1;
    GShare<HFq2, 12> g2;
    HG1 g1_out;
    HG2 g2_out;
    static OpenItem field(const HFr& sh, const HFr& mac) {
        OpenItem it;
        it.kind = 0;
        it.f_sh = sh;
        it.f_mac = mac;
        return it;
    }
    static OpenItem point(const GShare<HFq, 6>& s) {
        OpenItem it;
        it.kind = 1;
        it.g1 = s;
        return it;
    }
    static OpenItem point(const GShare<HFq2, 12>& s) {
        OpenItem it;
        it.kind = 2;
        it.g2 = s;
        return it;
    }
    size_t words() const { return kind == 0 ? 4 : (kind == 1 ? 13 : 25); }
};
static inline int open_many(czk_ctx* ctx, int scheme, std::vector<OpenItem>& items) {
    if (items.empty()) return CZK_OK;
    if (scheme == CZK_SCHEME_PLAIN) {
        for (OpenItem& it : items) {
            it.f_out = it.f_sh;
            it.g1_out = it.g1.sh;
            it.g2_out = it.g2.sh;
        }
        return CZK_OK;
    }
    const int n = ctx->nranks;
    size_t rec = 0;
    for (const OpenItem& it : items) rec += it.words();
    std::vector<uint64_t> send(rec), recv((size_t)n * rec);
    auto exchange = [&]() -> int {
        CZK_TRY(czk_net_allgather_host(ctx, send.data(), recv.data(), rec * 8));
        ctx->stats[2] += items.size() - 1;  // one logical broadcast per item
        return CZK_OK;
    };
    size_t off = 0;
    for (const OpenItem& it : items) {
        if (it.kind == 0) it.f_sh.to_limbs(send.data() + off);
        else if (it.kind == 1) send[off + 12] = (uint64_t)GShare<HFq, 6>::to_affine_limbs(it.g1.sh, send.data() + off);
        else send[off + 24] = (uint64_t)GShare<HFq2, 12>::to_affine_limbs(it.g2.sh, send.data() + off);
        off += it.words();
    }
    CZK_TRY(exchange());
    off = 0;
    for (OpenItem& it : items) {
        it.f_out = HFr::zero();
        it.g1_out = HG1::infinity();
        it.g2_out = HG2::infinity();
        for (int p = 0; p < n; p++) {
            const uint64_t* r = recv.data() + (size_t)p * rec + off;
            if (it.kind == 0) it.f_out = HFr::add(it.f_out, HFr::from_limbs(r));
            else if (it.kind == 1) it.g1_out.add(GShare<HFq, 6>::from_affine_limbs(r, (int)r[12]));
            else it.g2_out.add(GShare<HFq2, 12>::from_affine_limbs(r, (int)r[24]));
        }
        off += it.words();
    }
    if (scheme != CZK_SCHEME_SPDZ) return CZK_OK;
    // sigma records: x * mac_share - mac (mac_share = 1 at the king); every item's records must sum to zero
    off = 0;
    for (const OpenItem& it : items) {
        if (it.kind == 0) {
            HFr ms = ctx->rank == 0 ? HFr::one() : HFr::zero();
            HFr::sub(HFr::mul(ms, it.f_out), it.f_mac).to_limbs(send.data() + off);
        } else if (it.kind == 1) {
            HG1 dx = ctx->rank == 0 ? it.g1_out : HG1::infinity(), m = it.g1.mac;
            m.negate();
            dx.add(m);
            send[off + 12] = (uint64_t)GShare<HFq, 6>::to_affine_limbs(dx, send.data() + off);
        } else {
            HG2 dx = ctx->rank == 0 ? it.g2_out : HG2::infinity(), m = it.g2.mac;
            m.negate();
            dx.add(m);
            send[off + 24] = (uint64_t)GShare<HFq2, 12>::to_affine_limbs(dx, send.data() + off);
        }
        off += it.words();
    }
    CZK_TRY(exchange());
    off = 0;
    for (const OpenItem& it : items) {
        bool zero;
        if (it.kind == 0) {
            HFr sum = HFr::zero();
            for (int p = 0; p < n; p++) sum = HFr::add(sum, HFr::from_limbs(recv.data() + (size_t)p * rec + off));
            zero = sum.is_zero();
        } else if (it.kind == 1) {
            HG1 sum = HG1::infinity();
            for (int p = 0; p < n; p++) {
                const uint64_t* r = recv.data() + (size_t)p * rec + off;
                sum.add(GShare<HFq, 6>::from_affine_limbs(r, (int)r[12]));
            }
            zero = sum.is_inf();
        } else {
            HG2 sum = HG2::infinity();
            for (int p = 0; p < n; p++) {
                const uint64_t* r = recv.data() + (size_t)p * rec + off;
                sum.add(GShare<HFq2, 12>::from_affine_limbs(r, (int)r[24]));
            }
            zero = sum.is_inf();
        }
        if (!zero)
            return fail(ctx, CZK_ERR_PROTOCOL, it.kind == 0 ? "SPDZ MAC check failed (spdz.rs:129 assert!(sum.is_zero()))"
                                                            : "SPDZ group MAC check failed (spdz.rs:273 assert!(sum.is_zero()))");
        off += it.words();
    }
    return CZK_OK;
}

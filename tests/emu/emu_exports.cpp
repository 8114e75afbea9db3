// Host build of the DEVICE math headers (csrc/fp.cuh, csrc/ec.cuh) with the PTX carry flag
// emulated (csrc/carry.cuh).  Test infrastructure: lets the CPU-only test tier check the exact
// limb schedule the GPU runs against the oracle.  Not part of the product library.
#include <cstring>
#include "../../collaborative-zksnark_b200/csrc/ec.cuh"
#include "../../collaborative-zksnark_b200/csrc/msm_digits.cuh"
#include "../../collaborative-zksnark_b200/csrc/fq_inverse.cuh"

using namespace czk;
#define EXPORT extern "C" __attribute__((visibility("default")))

template <class F>
static F ld(const uint64_t* p) {
    F r;
    std::memcpy(r.l, p, sizeof r.l);
    return r;
}
template <class F>
static void st(uint64_t* p, const F& v) {
    std::memcpy(p, v.l, sizeof v.l);
}
static Fq2 ld2(const uint64_t* p) { return Fq2{ld<Fq>(p), ld<Fq>(p + 6)}; }
static void st2(uint64_t* p, const Fq2& v) {
    st(p, v.c0);
    st(p + 6, v.c1);
}

EXPORT void emu_fr_mul(uint64_t* r, const uint64_t* a, const uint64_t* b, size_t n) {
    for (size_t i = 0; i < n; i++) st(r + 4 * i, Fr::mul(ld<Fr>(a + 4 * i), ld<Fr>(b + 4 * i)));
}
EXPORT void emu_fr_add(uint64_t* r, const uint64_t* a, const uint64_t* b, size_t n) {
    for (size_t i = 0; i < n; i++) st(r + 4 * i, Fr::add(ld<Fr>(a + 4 * i), ld<Fr>(b + 4 * i)));
}
EXPORT void emu_fr_sub(uint64_t* r, const uint64_t* a, const uint64_t* b, size_t n) {
    for (size_t i = 0; i < n; i++) st(r + 4 * i, Fr::sub(ld<Fr>(a + 4 * i), ld<Fr>(b + 4 * i)));
}
EXPORT void emu_fr_neg(uint64_t* r, const uint64_t* a, size_t n) {
    for (size_t i = 0; i < n; i++) st(r + 4 * i, Fr::neg(ld<Fr>(a + 4 * i)));
}
EXPORT void emu_fr_from_mont(uint64_t* r, const uint64_t* a, size_t n) {
    for (size_t i = 0; i < n; i++) st(r + 4 * i, Fr::from_mont(ld<Fr>(a + 4 * i)));
}
EXPORT void emu_fr_inv(uint64_t* r, const uint64_t* a, size_t n) {
    for (size_t i = 0; i < n; i++) st(r + 4 * i, Fr::inv_fermat(ld<Fr>(a + 4 * i)));
}
EXPORT void emu_fq_mul(uint64_t* r, const uint64_t* a, const uint64_t* b, size_t n) {
    for (size_t i = 0; i < n; i++) st(r + 6 * i, Fq::mul(ld<Fq>(a + 6 * i), ld<Fq>(b + 6 * i)));
}
EXPORT void emu_fq_add(uint64_t* r, const uint64_t* a, const uint64_t* b, size_t n) {
    for (size_t i = 0; i < n; i++) st(r + 6 * i, Fq::add(ld<Fq>(a + 6 * i), ld<Fq>(b + 6 * i)));
}
EXPORT void emu_fq_sub(uint64_t* r, const uint64_t* a, const uint64_t* b, size_t n) {
    for (size_t i = 0; i < n; i++) st(r + 6 * i, Fq::sub(ld<Fq>(a + 6 * i), ld<Fq>(b + 6 * i)));
}
EXPORT void emu_fq_inv(uint64_t* r, const uint64_t* a, size_t n) {
    for (size_t i = 0; i < n; i++) st(r + 6 * i, Fq::inv_fermat(ld<Fq>(a + 6 * i)));
}
EXPORT void emu_fq_mul2(uint64_t* r1, uint64_t* r2, const uint64_t* a, const uint64_t* b, const uint64_t* c, const uint64_t* d, size_t n) {
    for (size_t i = 0; i < n; i++) {
        Fq x, y;
        Fq::mul2(ld<Fq>(a + 6 * i), ld<Fq>(b + 6 * i), ld<Fq>(c + 6 * i), ld<Fq>(d + 6 * i), x, y);
        st(r1 + 6 * i, x);
        st(r2 + 6 * i, y);
    }
}
// the batched-step binary GCD inversion (csrc/fq_inverse.cuh)
EXPORT void emu_fq_inv_bingcd(uint64_t* r, const uint64_t* a, size_t n) {
    for (size_t i = 0; i < n; i++) st(r + 6 * i, BinGcd<FqParams>::inverse(ld<Fq>(a + 6 * i)));
}
EXPORT void emu_fr_inv_bingcd(uint64_t* r, const uint64_t* a, size_t n) {
    for (size_t i = 0; i < n; i++) st(r + 4 * i, BinGcd<FrParams>::inverse(ld<Fr>(a + 4 * i)));
}
EXPORT void emu_fq2_mul(uint64_t* r, const uint64_t* a, const uint64_t* b, size_t n) {
    for (size_t i = 0; i < n; i++) st2(r + 12 * i, Fq2::mul(ld2(a + 12 * i), ld2(b + 12 * i)));
}
EXPORT void emu_fq2_sqr(uint64_t* r, const uint64_t* a, size_t n) {
    for (size_t i = 0; i < n; i++) st2(r + 12 * i, Fq2::sqr(ld2(a + 12 * i)));
}
EXPORT void emu_fq2_inv(uint64_t* r, const uint64_t* a, size_t n) {
    for (size_t i = 0; i < n; i++) st2(r + 12 * i, Fq2::inv_fermat(ld2(a + 12 * i)));
}

// sum of affine points with signs, accumulated exactly like a device bucket; result affine
// (via the device Fermat inversion).  sign[i] != 0 negates point i.  returns 1 if infinity.
EXPORT int emu_g1_sum(uint64_t* out_xy, const uint64_t* xy, const uint8_t* sign, size_t n, int tree) {
    XYZZ<Fq> acc = XYZZ<Fq>::infinity();
    if (!tree) {
        for (size_t i = 0; i < n; i++) {
            Fq x = ld<Fq>(xy + 12 * i), y = ld<Fq>(xy + 12 * i + 6);
            if (sign && sign[i]) y = Fq::neg(y);
            acc.add_affine(x, y);
        }
    } else {  // exercise the XYZZ+XYZZ path: pairwise then fold
        for (size_t i = 0; i < n; i += 2) {
            XYZZ<Fq> t = XYZZ<Fq>::infinity();
            for (size_t k = i; k < i + 2 && k < n; k++) {
                Fq x = ld<Fq>(xy + 12 * k), y = ld<Fq>(xy + 12 * k + 6);
                if (sign && sign[k]) y = Fq::neg(y);
                t.add_affine(x, y);
            }
            acc.add(t);
        }
    }
    if (acc.is_inf()) return 1;
    Fq zi = Fq::inv_fermat(acc.zz), zzzi = Fq::inv_fermat(acc.zzz);
    st(out_xy, Fq::mul(acc.x, zi));
    st(out_xy + 6, Fq::mul(acc.y, zzzi));
    return 0;
}
EXPORT int emu_g2_sum(uint64_t* out_xy, const uint64_t* xy, const uint8_t* sign, size_t n, int tree) {
    XYZZ<Fq2> acc = XYZZ<Fq2>::infinity();
    for (size_t i = 0; i < n; i += 2) {
        XYZZ<Fq2> t = XYZZ<Fq2>::infinity();
        for (size_t k = i; k < i + 2 && k < n; k++) {
            Fq2 x = ld2(xy + 24 * k), y = ld2(xy + 24 * k + 12);
            if (sign && sign[k]) y = Fq2::neg(y);
            if (tree) t.add_affine(x, y);
            else acc.add_affine(x, y);
        }
        if (tree) acc.add(t);
    }
    if (acc.is_inf()) return 1;
    Fq2 zi = Fq2::inv_fermat(acc.zz), zzzi = Fq2::inv_fermat(acc.zzz);
    st2(out_xy, Fq2::mul(acc.x, zi));
    st2(out_xy + 12, Fq2::mul(acc.y, zzzi));
    return 0;
}
// signed-digit decomposition used by the MSM kernels (csrc/msm_digits.cuh)
EXPORT void emu_signed_digits(int32_t* out, const uint64_t* scalar_canonical, unsigned c, unsigned nwin) {
    uint32_t s[8];
    std::memcpy(s, scalar_canonical, 32);
    signed_digits(s, c, nwin, out);
}

// The two-lane Fq2 product of the G2 bucket kernel (csrc/msm_batched.cu): lane 0 owns c0, lane 1 owns c1; each lane
// computes ONE lazily reduced sum of two products (Fp::mul_sum2).  Here both lanes are evaluated one after the other.
//   c0 = a0 b0 + (-5 a1) b1      c1 = a1 b0 + a0 b1
EXPORT void emu_fq2_mul_two_lanes(uint64_t* r, const uint64_t* a, const uint64_t* b, size_t n) {
    for (size_t i = 0; i < n; i++) {
        Fq2 x = ld2(a + 12 * i), y = ld2(b + 12 * i);
        uint32_t m5[12];
        Fq::neg_times5_unreduced(x.c1, m5);
        Fq2 out;
        out.c0 = Fq::mul_sum2(x.c0, y.c0, m5, y.c1);
        out.c1 = Fq::mul_sum2(x.c1, y.c0, x.c0.l, y.c1);
        st2(r + 12 * i, out);
    }
}
EXPORT void emu_fr_mul_sum2(uint64_t* r, const uint64_t* a, const uint64_t* b, const uint64_t* c, const uint64_t* d, size_t n) {
    for (size_t i = 0; i < n; i++) st(r + 4 * i, Fr::mul_sum2(ld<Fr>(a + 4 * i), ld<Fr>(b + 4 * i), ld<Fr>(c + 4 * i).l, ld<Fr>(d + 4 * i)));
}

// ---- the NTT tile arithmetic (csrc/ntt_tile.cuh) run thread by thread on the host: every pass, tile, phase and thread of
// ntt_run_tiles, with the tile accessor addressing the vector directly.  Same composition of passes, scalings and bit
// reversal as ntt_batch_dev (csrc/api.cu).  op: 0 fft, 1 ifft, 2 coset fft, 3 coset ifft, 4 ifft then coset fft.
#include <vector>
#include "../../collaborative-zksnark_b200/csrc/ntt_tile.cuh"
namespace {
struct HostTile {
    uint64_t* data;
    NttTileGeom g;
    size_t idx(unsigned e) const { return g.base | ((size_t)(e >> g.cl) << g.L) | (size_t)(e & ((1u << g.cl) - 1u)); }
    Fr load(unsigned e) const { return ld<Fr>(data + 4 * idx(e)); }
    void store(unsigned e, const Fr& v) const { st(data + 4 * idx(e), v); }
};
Fr host_ldtw(const uint32_t* table, size_t i) { return ntt_ld_words(table + 8 * i); }
void host_run_tiles(uint64_t* data, const uint32_t* tw, int n, bool inverse, bool dit, const NttScale& pre, const NttScale& post, int tile_log) {
    const NttPlan plan = ntt_make_plan(n, tile_log);
    for (int k = 0; k < plan.npass; k++) {
        const NttPass& p = plan.pass[dit ? plan.npass - 1 - k : k];
        const int tl = p.r + p.cl, L = n - p.s - p.r;
        const size_t blocks = (size_t)1 << (n - tl);
        for (size_t blk = 0; blk < blocks; blk++) {
            const size_t lowblk = blk & (((size_t)1 << (L - p.cl)) - 1), hi = blk >> (L - p.cl);
            HostTile tile{data, NttTileGeom{n, p.s, p.r, p.cl, L, (hi << (p.r + L)) | (lowblk << p.cl)}};
            const int nph = ntt_num_phases(p.r);
            for (int ph = 0; ph < nph; ph++) {
                int kp, jlo, ns;
                ntt_phase_geom(p.r, p.cl, dit, ph, kp, jlo, ns);
                for (unsigned u = 0; u < (1u << tl) / 8; u++) {
                    const bool first = k == 0 && ph == 0, last = k == plan.npass - 1 && ph == nph - 1;
                    if (dit) ntt_phase_thread<true, true>(tile, u, tile.g, inverse, kp, jlo, ns, tw, first, last, pre, post, host_ldtw);
                    else ntt_phase_thread<false, true>(tile, u, tile.g, inverse, kp, jlo, ns, tw, first, last, pre, post, host_ldtw);
                }
            }
        }
    }
}
std::vector<uint32_t> powers(const Fr& base, const Fr& c, size_t count) {
    std::vector<uint32_t> t(count * 8);
    Fr cur = c;
    for (size_t k = 0; k < count; k++) {
        for (int i = 0; i < 8; i++) t[8 * k + i] = cur.l[i];
        cur = Fr::mul(cur, base);
    }
    return t;
}
}  // namespace
EXPORT int emu_ntt(uint64_t* data, int n, int op, int tile_log, const uint64_t* omega, const uint64_t* gen, const uint64_t* gen_inv,
                   const uint64_t* size_inv) {
    if (n < 3 || tile_log < 4) return 0;  // smaller domains take the device's tiny kernel
    const size_t d = (size_t)1 << n;
    const Fr one = Fr::one(), w = ld<Fr>(omega), g = ld<Fr>(gen), gi = ld<Fr>(gen_inv), sinv = ld<Fr>(size_inv);
    const std::vector<uint32_t> tw = powers(w, one, d / 2 + 1);
    const int lo_log = n < 10 ? n : 10;
    const size_t nlo = (size_t)1 << lo_log, nhi = (size_t)1 << (n - lo_log);
    auto table_pair = [&](const Fr& base, const Fr& c, std::vector<uint32_t>& lo, std::vector<uint32_t>& hi) {
        lo = powers(base, one, nlo);
        hi = powers(Fr::pow_u64(base, nlo), c, nhi);
    };
    std::vector<uint32_t> g_lo, g_hi, g_hi_sinv, gi_lo, gi_hi;
    table_pair(g, one, g_lo, g_hi);
    table_pair(g, sinv, g_lo, g_hi_sinv);
    table_pair(gi, sinv, gi_lo, gi_hi);
    auto tables = [&](const std::vector<uint32_t>& lo, const std::vector<uint32_t>& hi, bool bitrev) {
        NttScale sc;
        sc.mode = 2;
        sc.bitrev = bitrev ? 1 : 0;
        sc.lo = lo.data();
        sc.hi = hi.data();
        sc.lo_log = lo_log;
        return sc;
    };
    const NttScale none;
    const bool inverse = op == 1 || op == 3 || op == 4;
    if (op == 4) {
        host_run_tiles(data, tw.data(), n, true, false, none, none, tile_log);
        host_run_tiles(data, tw.data(), n, false, true, tables(g_lo, g_hi_sinv, true), none, tile_log);
        return 1;
    }
    if (op == 5) {  // the same pair with the product's one-factor-per-position table (api.cu: g_sinv_br[p] = D^-1 g^bitrev(p))
        std::vector<uint32_t> nat = powers(g, sinv, d), br(nat.size());
        for (size_t pos = 0; pos < d; pos++) std::memcpy(&br[8 * pos], &nat[8 * (size_t)ntt_bitrev((uint32_t)pos, n)], 32);
        NttScale direct;
        direct.mode = 3;
        direct.lo = br.data();
        host_run_tiles(data, tw.data(), n, true, false, none, none, tile_log);
        host_run_tiles(data, tw.data(), n, false, true, direct, none, tile_log);
        return 1;
    }
    host_run_tiles(data, tw.data(), n, inverse, false, op == 2 ? tables(g_lo, g_hi, false) : none, none, tile_log);
    // k_bitrev_scale: swap into natural order, the inverse scalings ride along
    std::vector<uint64_t> tmp(data, data + 4 * d);
    for (size_t i = 0; i < d; i++) {
        const size_t r = ntt_bitrev((uint32_t)i, n);
        Fr v = ld<Fr>(tmp.data() + 4 * r);  // the element at position r lands at i
        if (op == 1) v = Fr::mul(v, sinv);
        if (op == 3) v = Fr::mul(v, Fr::mul(host_ldtw(gi_lo.data(), i & (nlo - 1)), host_ldtw(gi_hi.data(), i >> lo_log)));
        st(data + 4 * i, v);
    }
    return 1;
}

// A whole mixed-radix transform over 3 * 2^n points the way the device runs it (api.cu, ntt_mixed_batch_dev): de-interleave,
// three radix-2 transforms through the tile emulation above (op 0 forward / 1 inverse, natural order out), then
// ntt_mixed_combine3 - the device's own combining arithmetic - with w^j (or w^-j), zeta = w^M and 3^-1 on the inverse.
// omega2 / size_inv2: the radix-2 domain of 2^n points; omega3: get_root_of_unity(3 * 2^n) (or its inverse for inverse = 1).
EXPORT int emu_ntt_mixed(uint64_t* data, int n, int inverse, int tile_log, const uint64_t* omega2, const uint64_t* gen, const uint64_t* gen_inv,
                         const uint64_t* size_inv2, const uint64_t* omega3, const uint64_t* third_inv) {
    const size_t M = (size_t)1 << n;
    std::vector<uint64_t> sub(3 * M * 4);
    for (size_t i = 0; i < 3 * M; i++) std::memcpy(&sub[((i % 3) * M + i / 3) * 4], data + 4 * i, 32);
    for (int r = 0; r < 3; r++)
        if (!emu_ntt(sub.data() + r * M * 4, n, inverse ? 1 : 0, tile_log, omega2, gen, gen_inv, size_inv2)) return 0;
    const Fr w = ld<Fr>(omega3), c = ld<Fr>(third_inv), zeta = Fr::pow_u64(w, M);
    Fr wj = Fr::one();
    for (size_t j = 0; j < M; j++) {
        const Fr y[3] = {ld<Fr>(sub.data() + 4 * j), ld<Fr>(sub.data() + 4 * (M + j)), ld<Fr>(sub.data() + 4 * (2 * M + j))};
        Fr out[3];
        ntt_mixed_combine3(y, wj, zeta, c, inverse != 0, out);
        for (int s3 = 0; s3 < 3; s3++) st(data + 4 * (j + (size_t)s3 * M), out[s3]);
        wj = Fr::mul(wj, w);
    }
    return 1;
}


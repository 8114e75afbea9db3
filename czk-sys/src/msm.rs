//! `Msm` on the device: replaces `AffineMsm` (mpc-algebra/src/share/msm.rs:31-36) and, through it,
//! `AffineCurve::multi_scalar_mul` -> `VariableBaseMSM::multi_scalar_mul` (algebra/ec/src/lib.rs:302-311,
//! algebra/ec/src/msm/variable_base.rs:12-106) for BLS12-377 G1 and G2.
use crate::{check, ffi, fr_limbs, with_ctx, DevVec};
use ark_bls12_377::{Fq, Fq2, Fr, G1Affine, G1Projective, G2Affine, G2Projective};
use ark_ec::{AffineCurve, ProjectiveCurve};
use mpc_algebra::msm::Msm;
use std::marker::PhantomData;

/// x | y limbs and infinity bytes of a slice of affine points (`GroupAffine` is repr(Rust): marshalled explicitly,
/// short_weierstrass_jacobian.rs:43-49).  One O(n) copy - the reference copies its bases on every call too
/// (wire/pairing.rs:751-753); a resident `Bases` pays it once per CRS.
pub fn marshal_g1(points: &[G1Affine]) -> (Vec<u64>, Vec<u8>) {
    let mut xy = Vec::with_capacity(12 * points.len());
    let mut inf = Vec::with_capacity(points.len());
    for p in points {
        xy.extend_from_slice(&(p.x.0).0);
        xy.extend_from_slice(&(p.y.0).0);
        inf.push(p.infinity as u8);
    }
    (xy, inf)
}
pub fn marshal_g2(points: &[G2Affine]) -> (Vec<u64>, Vec<u8>) {
    let mut xy = Vec::with_capacity(24 * points.len());
    let mut inf = Vec::with_capacity(points.len());
    for p in points {
        for c in [&p.x.c0, &p.x.c1, &p.y.c0, &p.y.c1].iter() {
            xy.extend_from_slice(&(c.0).0);
        }
        inf.push(p.infinity as u8);
    }
    (xy, inf)
}
fn fq(l: &[u64]) -> Fq {
    let mut b = ark_ff::BigInteger384::default();
    b.0.copy_from_slice(&l[..6]);
    ark_ff::Fp384::<ark_bls12_377::FqParameters>(b, PhantomData)
}
/// The library returns an affine-normalised Jacobian triple: (x, y, 1), or (1, 1, 0) for the identity
/// (short_weierstrass_jacobian.rs:440-448).
pub fn g1_from_out(o: &[u64; 18]) -> G1Projective {
    G1Projective::new(fq(&o[0..6]), fq(&o[6..12]), fq(&o[12..18]))
}
pub fn g2_from_out(o: &[u64; 36]) -> G2Projective {
    G2Projective::new(Fq2::new(fq(&o[0..6]), fq(&o[6..12])), Fq2::new(fq(&o[12..18]), fq(&o[18..24])), Fq2::new(fq(&o[24..30]), fq(&o[30..36])))
}

/// Drop-in for `AffineMsm<G>`: `impl Msm<G, G::ScalarField>` with the MSM on the GPU.
pub struct GpuAffineMsm<G: AffineCurve>(PhantomData<G>);

impl Msm<G1Affine, Fr> for GpuAffineMsm<G1Affine> {
    fn msm(bases: &[G1Affine], scalars: &[Fr]) -> G1Affine {
        let n = bases.len().min(scalars.len()); // variable_base.rs:16
        let (xy, inf) = marshal_g1(&bases[..n]);
        let mut out = [0u64; 18];
        with_ctx(|c| check(c, "czk_msm_g1", unsafe {
            ffi::czk_msm_g1(c, xy.as_ptr(), inf.as_ptr(), fr_limbs(&scalars[..n]).as_ptr(), 1, n, out.as_mut_ptr())
        }));
        g1_from_out(&out).into_affine()
    }
}
impl Msm<G2Affine, Fr> for GpuAffineMsm<G2Affine> {
    fn msm(bases: &[G2Affine], scalars: &[Fr]) -> G2Affine {
        let n = bases.len().min(scalars.len());
        let (xy, inf) = marshal_g2(&bases[..n]);
        let mut out = [0u64; 36];
        with_ctx(|c| check(c, "czk_msm_g2", unsafe {
            ffi::czk_msm_g2(c, xy.as_ptr(), inf.as_ptr(), fr_limbs(&scalars[..n]).as_ptr(), 1, n, out.as_mut_ptr())
        }));
        g2_from_out(&out).into_affine()
    }
}

/// A CRS query kept on the device across proofs (the reference re-passes the same `pk.*_query` every time):
/// `czk_bases_upload` + `czk_bases_precompute` once, `czk_msm_bases` per proof.
pub struct Bases {
    pub ptr: *mut ffi::czk_bases,
    pub curve: i32,
}
impl Bases {
    pub fn g1(points: &[G1Affine]) -> Self {
        let (xy, inf) = marshal_g1(points);
        Self::upload(1, &xy, &inf, points.len())
    }
    pub fn g2(points: &[G2Affine]) -> Self {
        let (xy, inf) = marshal_g2(points);
        Self::upload(2, &xy, &inf, points.len())
    }
    fn upload(curve: i32, xy: &[u64], inf: &[u8], n: usize) -> Self {
        with_ctx(|c| {
            let mut p = std::ptr::null_mut();
            check(c, "czk_bases_upload", unsafe { ffi::czk_bases_upload(c, curve, xy.as_ptr(), inf.as_ptr(), n, &mut p) });
            check(c, "czk_bases_precompute", unsafe { ffi::czk_bases_precompute(c, p, 0) });
            Bases { ptr: p, curve }
        })
    }
    /// sum_i scalars[i] * bases[off + i] with the scalars already on the device
    pub fn msm_g1(&self, off: usize, scalars: &DevVec, sc_off: usize, n: usize) -> G1Projective {
        assert_eq!(self.curve, 1);
        let mut out = [0u64; 18];
        with_ctx(|c| check(c, "czk_msm_bases", unsafe { ffi::czk_msm_bases(c, self.ptr, off, scalars.ptr, sc_off, 1, n, out.as_mut_ptr()) }));
        g1_from_out(&out)
    }
    pub fn msm_g2(&self, off: usize, scalars: &DevVec, sc_off: usize, n: usize) -> G2Projective {
        assert_eq!(self.curve, 2);
        let mut out = [0u64; 36];
        with_ctx(|c| check(c, "czk_msm_bases", unsafe { ffi::czk_msm_bases(c, self.ptr, off, scalars.ptr, sc_off, 1, n, out.as_mut_ptr()) }));
        g2_from_out(&out)
    }
}
impl Drop for Bases {
    fn drop(&mut self) {
        with_ctx(|c| unsafe { ffi::czk_bases_free(c, self.ptr) });
    }
}

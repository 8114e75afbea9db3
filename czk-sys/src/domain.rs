//! `EvaluationDomain` on the device: replaces the bodies of `Radix2EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place`
//! (algebra/poly/src/domain/radix2/mod.rs:99-117, radix2/fft.rs:22-260) for BLS12-377 Fr and for shares of Fr.
//!
//! The transforms are linear, so a vector of shares is transformed component by component (value vector, MAC vector):
//! exactly what the generic `T: DomainCoeff<F>` code does element-wise.  `Gpu::transform_batch` runs several vectors in
//! one grid per pass (czk_ntt_fr_batch), `IFFT_COSET_FFT` fuses the witness map's pair (r1cs_to_qap.rs:85-90).
use crate::{check, ffi, fr_limbs_mut, with_ctx, DevVec};
use ark_bls12_377::Fr;
use ark_poly::{EvaluationDomain, Radix2EvaluationDomain};

pub const FFT: i32 = ffi::CZK_NTT_FFT;
pub const IFFT: i32 = ffi::CZK_NTT_IFFT;
pub const COSET_FFT: i32 = ffi::CZK_NTT_COSET_FFT;
pub const COSET_IFFT: i32 = ffi::CZK_NTT_COSET_IFFT;
pub const IFFT_COSET_FFT: i32 = ffi::CZK_NTT_IFFT_COSET_FFT;

/// In-place transform of a host vector of plain field elements (what `fft_in_place::<Fr>` does).
/// `coeffs` is resized to the domain size with zeros first, like radix2/mod.rs:100-101.
pub fn transform_in_place(domain: &Radix2EvaluationDomain<Fr>, coeffs: &mut Vec<Fr>, op: i32) {
    assert!(coeffs.len() <= domain.size());
    coeffs.resize(domain.size(), Fr::from(0u64));
    let (inverse, coset) = (op & 1, (op >> 1) & 1);
    with_ctx(|c| check(c, "czk_ntt_fr", unsafe {
        ffi::czk_ntt_fr(c, fr_limbs_mut(coeffs).as_mut_ptr(), domain.log_size_of_group, inverse, coset)
    }));
}

/// The same transform over several device-resident vectors, one grid per pass.
pub fn transform_batch(log_size: u32, vecs: &[&DevVec], op: i32) {
    let ptrs: Vec<*mut ffi::czk_vec> = vecs.iter().map(|v| v.ptr).collect();
    with_ctx(|c| check(c, "czk_ntt_vec_batch", unsafe { ffi::czk_ntt_vec_batch(c, ptrs.as_ptr(), ptrs.len() as i32, log_size, op) }));
}

/// Patch for algebra/poly/src/domain/radix2/mod.rs (the four method bodies):
///
/// ```ignore
/// fn fft_in_place<T: DomainCoeff<F>>(&self, coeffs: &mut Vec<T>) {
///     if let Some(v) = czk_sys::domain::as_fr_vec(coeffs) { return czk_sys::domain::transform_in_place(self.as_fr(), v, FFT); }
///     if let Some((sh, mac)) = czk_sys::shares::as_share_vecs(coeffs) { /* upload both, transform_batch, download */ }
///     coeffs.resize(self.size(), T::zero()); self.in_order_fft_in_place(&mut *coeffs)      // everything else: unchanged
/// }
/// ```
pub fn domain_constants(log_size: u32) -> ([u64; 4], [u64; 4], [u64; 4], [u64; 4]) {
    let (mut g, mut gi, mut si, mut geni) = ([0u64; 4], [0u64; 4], [0u64; 4], [0u64; 4]);
    let rc = unsafe { ffi::czk_domain_params(log_size, g.as_mut_ptr(), gi.as_mut_ptr(), si.as_mut_ptr(), geni.as_mut_ptr()) };
    assert_eq!(rc, ffi::CZK_OK, "czk_domain_params");
    (g, gi, si, geni)
}

/// `MixedRadixEvaluationDomain` of 3 * 2^log_m points (algebra/poly/src/domain/mixed_radix.rs:130-157): the Plonk prover's
/// wire domain (mpc-plonk/src/relations/flat.rs:282-300).  Same ops, device-resident vectors of 3 * 2^log_m elements.
///
/// ```ignore
/// // mixed_radix.rs, fft_in_place:  if self.size == 3 << self.log_size_of_group { upload; transform_batch_mixed(..); download }
/// ```
pub fn transform_batch_mixed(log_m: u32, vecs: &[&DevVec], op: i32) {
    let ptrs: Vec<*mut ffi::czk_vec> = vecs.iter().map(|v| v.ptr).collect();
    with_ctx(|c| check(c, "czk_ntt_mixed_vec_batch", unsafe { ffi::czk_ntt_mixed_vec_batch(c, ptrs.as_ptr(), ptrs.len() as i32, log_m, op) }));
}
pub fn mixed_domain_constants(log_m: u32) -> ([u64; 4], [u64; 4], [u64; 4], [u64; 4]) {
    let (mut g, mut gi, mut si, mut geni) = ([0u64; 4], [0u64; 4], [0u64; 4], [0u64; 4]);
    let rc = unsafe { ffi::czk_mixed_domain_params(log_m, g.as_mut_ptr(), gi.as_mut_ptr(), si.as_mut_ptr(), geni.as_mut_ptr()) };
    assert_eq!(rc, ffi::CZK_OK, "czk_mixed_domain_params");
    (g, gi, si, geni)
}
/// In-place mixed-radix transform of a host vector of plain field elements (`fft_in_place::<Fr>` of a 3 * 2^k-point domain).
pub fn transform_in_place_mixed(log_m: u32, coeffs: &mut Vec<Fr>, op: i32) {
    assert!(coeffs.len() <= 3usize << log_m);
    coeffs.resize(3usize << log_m, Fr::from(0u64));
    let (inverse, coset) = (op & 1, (op >> 1) & 1);
    with_ctx(|c| check(c, "czk_ntt_mixed_fr", unsafe { ffi::czk_ntt_mixed_fr(c, fr_limbs_mut(coeffs).as_mut_ptr(), log_m, inverse, coset) }));
}

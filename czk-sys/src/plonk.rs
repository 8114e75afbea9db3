//! `Prover::prove_wiring` on the device (mpc-plonk/src/lib.rs:199-258, with prove_unit_product :110-197 and eval / commit
//! :343-400) with the CALLER'S Fiat-Shamir transcript: the library calls back into `FiatShamirRng<Blake2s>` wherever the
//! reference calls `fs_rng.absorb` / `fs_rng.gen`, so the challenges are the reference's.
use crate::msm::Bases;
use crate::{check, ffi, with_ctx, DevVec};
use ark_bls12_377::{Fq, Fr, G1Affine};
use mpc_plonk::FiatShamirRng;
use std::os::raw::{c_int, c_void};

type Transcript = FiatShamirRng<blake2::Blake2s>;

unsafe extern "C" fn absorb_g1(user: *mut c_void, xy: *const u64, inf: c_int) {
    let t = &mut *(user as *mut Transcript);
    let l = std::slice::from_raw_parts(xy, 12);
    let p = G1Affine::new(fq(&l[0..6]), fq(&l[6..12]), inf != 0);
    // lib.rs:393-396: fs_rng.absorb(&to_bytes![c]) on the (publicized) commitment
    t.absorb(&ark_ff::to_bytes![p].expect("failed serialization"));
}
unsafe extern "C" fn challenge(user: *mut c_void, out: *mut u64) {
    let t = &mut *(user as *mut Transcript);
    let f: Fr = t.gen();
    std::ptr::copy_nonoverlapping((f.0).0.as_ptr(), out, 4);
}
fn fq(l: &[u64]) -> Fq {
    let mut b = ark_ff::BigInteger384::default();
    b.0.copy_from_slice(l);
    ark_ff::Fp384::<ark_bls12_377::FqParameters>(b, std::marker::PhantomData)
}

/// p: this party's shares of the wire polynomial's coefficients (value, and MAC under SPDZ); w: pk.w (public).
/// `domain_size` is `circ.domains.wires.size()`: 3 * 2^k for the reference's mixed-radix wire domain, or a power of two.
pub fn prove_wiring(scheme: i32, powers_of_g: &Bases, domain_size: usize, p_sh: &DevVec, p_mac: Option<&DevVec>, w: &DevVec,
                    fs_rng: &mut Transcript) -> (ffi::czk_plonk_wiring_proof, ffi::czk_plonk_wiring_proof) {
    let log_size = domain_size.trailing_zeros();
    if domain_size == 3usize << log_size {
        let tr = ffi::czk_plonk_transcript { user: fs_rng as *mut _ as *mut c_void, absorb_g1: Some(absorb_g1), challenge: Some(challenge) };
        let mut share: ffi::czk_plonk_wiring_proof = unsafe { std::mem::zeroed() };
        let mut revealed: ffi::czk_plonk_wiring_proof = unsafe { std::mem::zeroed() };
        with_ctx(|c| check(c, "czk_plonk_prove_wiring_mixed", unsafe {
            ffi::czk_plonk_prove_wiring_mixed(c, scheme, powers_of_g.ptr, log_size, p_sh.ptr,
                                              p_mac.map(|m| m.ptr as *const _).unwrap_or(std::ptr::null()), w.ptr, &tr, &mut share, &mut revealed,
                                              std::ptr::null_mut())
        }));
        return (share, revealed);
    }
    assert_eq!(domain_size, 1usize << log_size, "wire domain must have 2^k or 3 * 2^k points");
    let tr = ffi::czk_plonk_transcript { user: fs_rng as *mut _ as *mut c_void, absorb_g1: Some(absorb_g1), challenge: Some(challenge) };
    let mut share: ffi::czk_plonk_wiring_proof = unsafe { std::mem::zeroed() };
    let mut revealed: ffi::czk_plonk_wiring_proof = unsafe { std::mem::zeroed() };
    with_ctx(|c| check(c, "czk_plonk_prove_wiring", unsafe {
        ffi::czk_plonk_prove_wiring(c, scheme, powers_of_g.ptr, log_size, p_sh.ptr, p_mac.map(|m| m.ptr as *const _).unwrap_or(std::ptr::null()),
                                    w.ptr, &tr, &mut share, &mut revealed, std::ptr::null_mut())
    }));
    (share, revealed)
}

//! Rust side of the drop-in boundary of czk-b200 (see INTEGRATION.md at the repository root).
//!
//! `ffi` is the raw, generated declaration of every function of `include/*.h`.  The other modules are the safe layer the
//! reference's call sites switch to - each keeps the reference's own trait or function signature:
//!
//! | module    | replaces (reference file:line)                                                                  |
//! |-----------|--------------------------------------------------------------------------------------------------|
//! | `msm`     | `Msm::msm` / `AffineMsm` (mpc-algebra/src/share/msm.rs:6-48), `AffineCurve::multi_scalar_mul`      |
//! | `domain`  | `Radix2EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place` (poly/src/domain/radix2/mod.rs:99-117) |
//! | `shares`  | `FieldShare::{batch_open,batch_mul,batch_inv,batch_div,partial_products}` (share/field.rs:44-182)  |
//! | `net`     | `MpcNet` (mpc-net/src/lib.rs:28-70) over NCCL                                                       |
//! | `groth16` | `create_proof` + `reveal` on shares (mpc-snarks/src/groth/prover.rs:66-177)                         |
//! | `plonk`   | `Prover::prove_wiring` (mpc-plonk/src/lib.rs:110-258,343-400) with the caller's FiatShamirRng       |
//!
//! Error convention: the reference panics on misuse, IO failure, a failed MAC check or a failed degree check; every wrapper
//! here turns a non-zero status into `panic!` with `czk_last_error`, so call sites keep their signatures.
//!
//! This crate is NOT compiled in the repository's build image (no Rust toolchain there); it is written against the
//! reference workspace and checked for symbol coverage by tests/test_czk_sys_crate.py.
pub mod ffi;

pub mod domain;
pub mod groth16;
pub mod msm;
pub mod net;
#[cfg(feature = "plonk")]
pub mod plonk;
pub mod shares;

use std::ffi::CStr;
use std::sync::Mutex;

/// The process-wide context of this party (one party = one process = one GPU), mirroring the reference's global
/// `Mutex<Connections>` singleton (mpc-net/src/multi.rs:14-16).  Created on first use on device `LOCAL_RANK` (0 if unset).
pub struct Ctx(pub *mut ffi::czk_ctx);
unsafe impl Send for Ctx {}

lazy_static::lazy_static! {
    static ref CTX: Mutex<Ctx> = {
        let dev: i32 = std::env::var("LOCAL_RANK").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        let mut p: *mut ffi::czk_ctx = std::ptr::null_mut();
        let rc = unsafe { ffi::czk_ctx_create(dev, &mut p) };
        if rc != ffi::CZK_OK {
            panic!("czk_ctx_create({}): {}", dev, last_error(std::ptr::null()));
        }
        Mutex::new(Ctx(p))
    };
}

/// Run `f` with the context pointer; calls are serialised like the reference's `get_ch!()` accesses.
pub fn with_ctx<R>(f: impl FnOnce(*mut ffi::czk_ctx) -> R) -> R {
    let g = CTX.lock().unwrap();
    f(g.0)
}

pub fn last_error(ctx: *const ffi::czk_ctx) -> String {
    unsafe {
        let p = ffi::czk_last_error(ctx);
        if p.is_null() { String::new() } else { CStr::from_ptr(p).to_string_lossy().into_owned() }
    }
}

/// status code -> panic (the reference has no error returns on this path)
pub fn check(ctx: *mut ffi::czk_ctx, what: &str, rc: std::os::raw::c_int) {
    if rc != ffi::CZK_OK {
        panic!("{}: czk error {}: {}", what, rc, last_error(ctx));
    }
}

/// A device-resident `Vec<Fr>` (RAII over `czk_vec`).
pub struct DevVec {
    pub ptr: *mut ffi::czk_vec,
    pub len: usize,
}
impl DevVec {
    pub fn zeros(n: usize) -> Self {
        with_ctx(|c| {
            let mut p = std::ptr::null_mut();
            check(c, "czk_vec_alloc", unsafe { ffi::czk_vec_alloc(c, n, &mut p) });
            DevVec { ptr: p, len: n }
        })
    }
    /// `limbs`: n x 4 Montgomery limbs - `&[Fr]` reinterpreted (Fp256 is repr(transparent) over [u64; 4] in the patched fork)
    pub fn from_limbs(limbs: &[u64]) -> Self {
        assert_eq!(limbs.len() % 4, 0);
        let v = Self::zeros(limbs.len() / 4);
        with_ctx(|c| check(c, "czk_vec_upload", unsafe { ffi::czk_vec_upload(c, v.ptr, 0, limbs.as_ptr(), v.len) }));
        v
    }
    pub fn to_limbs(&self, out: &mut [u64]) {
        assert_eq!(out.len(), 4 * self.len);
        with_ctx(|c| check(c, "czk_vec_download", unsafe { ffi::czk_vec_download(c, self.ptr, 0, out.as_mut_ptr(), self.len) }));
    }
}
impl Drop for DevVec {
    fn drop(&mut self) {
        with_ctx(|c| unsafe { ffi::czk_vec_free(c, self.ptr) });
    }
}

/// `&[Fr]` as limbs.  Requires `#[repr(transparent)]` on `Fp256` and `BigInteger256` (one-line patches to the fork, no
/// behavioural effect: algebra/ff/src/fields/macros.rs:103-108, algebra/ff/src/biginteger/macros.rs).
pub fn fr_limbs(v: &[ark_bls12_377::Fr]) -> &[u64] {
    unsafe { std::slice::from_raw_parts(v.as_ptr() as *const u64, 4 * v.len()) }
}
pub fn fr_limbs_mut(v: &mut [ark_bls12_377::Fr]) -> &mut [u64] {
    unsafe { std::slice::from_raw_parts_mut(v.as_mut_ptr() as *mut u64, 4 * v.len()) }
}

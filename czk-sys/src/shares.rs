//! `FieldShare` batch operations on the device: replaces, for additive / SPDZ / GSZ shares of BLS12-377 Fr,
//!   batch_open        mpc-algebra/src/share/add.rs:121-125, spdz.rs:166-185, gsz20/mod.rs:286-299
//!   batch_mul         share/field.rs:97-127 (Beaver, stub triples wire/field.rs:41-77), gsz20/mod.rs:309-315
//!   batch_inv / batch_div / partial_products   share/field.rs:135-182
//! reached from the `MpcField` hooks `batch_product_in_place`, `batch_division_in_place`, `partial_products_in_place`
//! (wire/field.rs:358-455).
use crate::{check, ffi, with_ctx, DevVec};

#[derive(Clone, Copy, PartialEq, Eq)]
pub enum Scheme {
    Additive = ffi::CZK_SCHEME_ADDITIVE as isize,
    Spdz = ffi::CZK_SCHEME_SPDZ as isize,
    Gsz = ffi::CZK_SCHEME_GSZ as isize,
}

/// A shared vector on the device: value shares and (SPDZ) MAC shares.
pub struct SharedVec {
    pub sh: DevVec,
    pub mac: Option<DevVec>,
}
impl SharedVec {
    fn mac_ptr(&self) -> *mut ffi::czk_vec {
        self.mac.as_ref().map(|m| m.ptr).unwrap_or(std::ptr::null_mut())
    }
}

/// `batch_open`: every party learns the k values (SPDZ: with the MAC check; a failure panics like spdz.rs:182).
pub fn batch_open(scheme: Scheme, x: &SharedVec) -> DevVec {
    let out = DevVec::zeros(x.sh.len);
    with_ctx(|c| match scheme {
        Scheme::Gsz => panic!("use gsz_open(degree)"),
        _ => check(c, "czk_batch_open", unsafe { ffi::czk_batch_open(c, scheme as i32, x.sh.ptr, x.mac_ptr(), out.ptr, x.sh.len) }),
    });
    out
}
/// `batch_mul`: x <- x * y on shares; one fused exchange for both Beaver opens (csrc/shares.cu).
pub fn batch_mul(scheme: Scheme, x: &mut SharedVec, y: &SharedVec) {
    with_ctx(|c| check(c, "czk_beaver_batch_mul", unsafe {
        ffi::czk_beaver_batch_mul(c, scheme as i32, x.sh.ptr, x.mac_ptr(), y.sh.ptr, y.mac_ptr(), x.sh.len)
    }));
}
pub fn batch_inv(scheme: Scheme, x: &mut SharedVec) {
    with_ctx(|c| check(c, "czk_share_batch_inv", unsafe { ffi::czk_share_batch_inv(c, scheme as i32, x.sh.ptr, x.mac_ptr(), x.sh.len) }));
}
/// x <- x / y; y returns holding the shares of 1 / y (share/field.rs:155-158)
pub fn batch_div(scheme: Scheme, x: &mut SharedVec, y: &mut SharedVec) {
    with_ctx(|c| check(c, "czk_share_batch_div", unsafe {
        ffi::czk_share_batch_div(c, scheme as i32, x.sh.ptr, x.mac_ptr(), y.sh.ptr, y.mac_ptr(), x.sh.len)
    }));
}
pub fn partial_products(scheme: Scheme, x: &mut SharedVec) {
    with_ctx(|c| check(c, "czk_share_partial_products", unsafe {
        ffi::czk_share_partial_products(c, scheme as i32, x.sh.ptr, x.mac_ptr(), x.sh.len)
    }));
}
/// GSZ20: open_degree_vec at `degree` (gsz20/mod.rs:440-459); a failed degree check panics like the reference's assert.
pub fn gsz_open(x: &DevVec, degree: u32) -> DevVec {
    let out = DevVec::zeros(x.len);
    with_ctx(|c| check(c, "czk_gsz_open", unsafe { ffi::czk_gsz_open(c, x.ptr, degree, out.ptr, x.len) }));
    out
}
/// GSZ20 batch_mult (gsz20/mod.rs:559-594): the triple is queued for `check_accumulated_field_products`.
pub fn gsz_batch_mul(x: &mut DevVec, y: &DevVec) {
    with_ctx(|c| check(c, "czk_gsz_batch_mul", unsafe { ffi::czk_gsz_batch_mul(c, x.ptr, y.ptr, x.len, 1) }));
}
pub fn gsz_check_accumulated_field_products() {
    with_ctx(|c| check(c, "czk_gsz_check_products", unsafe { ffi::czk_gsz_check_products(c, std::ptr::null_mut()) }));
}

//! The Groth16 prover loop on shares: replaces `create_proof` + `pf.reveal()` on the squaring-circuit path
//! (mpc-snarks/src/groth/prover.rs:66-177, r1cs_to_qap.rs:47-112, proof.rs:130-139) and, for any circuit, the same loop fed
//! with the constraint matrices (`czk_groth16_prove_r1cs`).
use crate::msm::{marshal_g1, marshal_g2};
use crate::{check, ffi, fr_limbs, with_ctx};
use ark_bls12_377::{Bls12_377, Fr, G1Affine, G2Affine};

/// A `ProvingKey<Bls12_377>` resident on the device (queries + merged-window tables).
pub struct DevicePk(pub *mut ffi::czk_pk);

pub fn upload_pk(n_squarings: usize, pk: &ark_groth16::ProvingKey<Bls12_377>) -> DevicePk {
    let (a, ai) = marshal_g1(&pk.a_query);
    let (b1, b1i) = marshal_g1(&pk.b_g1_query);
    let (b2, b2i) = marshal_g2(&pk.b_g2_query);
    let (h, hi) = marshal_g1(&pk.h_query);
    let (l, li) = marshal_g1(&pk.l_query);
    let (vk1, _) = marshal_g1(&[pk.vk.alpha_g1, pk.beta_g1, pk.delta_g1]);
    let (vk2, _) = marshal_g2(&[pk.vk.beta_g2, pk.vk.gamma_g2, pk.vk.delta_g2]);
    with_ctx(|c| {
        let mut p = std::ptr::null_mut();
        check(c, "czk_groth16_pk_upload", unsafe {
            ffi::czk_groth16_pk_upload(c, n_squarings, a.as_ptr(), ai.as_ptr(), b1.as_ptr(), b1i.as_ptr(), b2.as_ptr(), b2i.as_ptr(), h.as_ptr(),
                                       hi.as_ptr(), l.as_ptr(), li.as_ptr(), vk1.as_ptr(), vk2.as_ptr(), &mut p)
        });
        DevicePk(p)
    })
}

/// The revealed proof and this party's share of it, as affine limbs A (12) | B (24) | C (12) + three infinity bytes.
pub struct ProofLimbs {
    pub share: [u64; 48],
    pub share_inf: [u8; 3],
    pub revealed: [u64; 48],
    pub revealed_inf: [u8; 3],
}

/// `chain`: this party's shares of w_0 .. w_{n-1}, out (n + 1 elements); `r`, `s`: its shares of the prover randomness.
pub fn prove(scheme: i32, pk: &DevicePk, chain: &[Fr], r: &Fr, s: &Fr) -> ProofLimbs {
    let mut out = ProofLimbs { share: [0; 48], share_inf: [0; 3], revealed: [0; 48], revealed_inf: [0; 3] };
    with_ctx(|c| check(c, "czk_groth16_prove", unsafe {
        ffi::czk_groth16_prove(c, scheme, pk.0, fr_limbs(chain).as_ptr(), fr_limbs(std::slice::from_ref(r)).as_ptr(),
                               fr_limbs(std::slice::from_ref(s)).as_ptr(), out.share.as_mut_ptr(), out.share_inf.as_mut_ptr(),
                               out.revealed.as_mut_ptr(), out.revealed_inf.as_mut_ptr())
    }));
    out
}
impl Drop for DevicePk {
    fn drop(&mut self) {
        with_ctx(|c| unsafe { ffi::czk_groth16_pk_free(c, self.0) });
    }
}
#[allow(dead_code)]
fn _types(_: G1Affine, _: G2Affine) {}

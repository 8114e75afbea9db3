//! `MpcNet` over NCCL: replaces `MpcMultiNet` (mpc-net/src/multi.rs:145-242) inside one NVSwitch domain.
//! One process per party, one GPU per process; `init_from_file(path, id)` becomes (rank, n_parties, ncclUniqueId).
use crate::{check, ffi, with_ctx};
use mpc_net::{MpcNet, Stats};

pub struct MpcNcclNet;

/// The launcher (torchrun, mpirun, or the reference's `bench.zsh`) hands every party the 128-byte id made by party 0.
pub fn init(rank: usize, n_parties: usize, unique_id: &[u8; 128]) {
    with_ctx(|c| check(c, "czk_net_init", unsafe { ffi::czk_net_init(c, rank as i32, n_parties as i32, unique_id.as_ptr()) }));
}
pub fn unique_id() -> [u8; 128] {
    let mut id = [0u8; 128];
    assert_eq!(unsafe { ffi::czk_net_unique_id(id.as_mut_ptr()) }, ffi::CZK_OK, "czk_net_unique_id");
    id
}

impl MpcNet for MpcNcclNet {
    fn am_king() -> bool { Self::party_id() == 0 }
    fn n_parties() -> usize { with_ctx(|c| unsafe { ffi::czk_net_n_parties(c) } as usize) }
    fn party_id() -> usize { with_ctx(|c| unsafe { ffi::czk_net_party_id(c) } as usize) }
    fn init_from_file(_path: &str, _party_id: usize) {
        // The hosts file has no meaning inside one box: parties are ranks.  RANK / WORLD_SIZE come from the launcher,
        // the unique id from party 0 through the launcher's rendezvous (see collaborative-zksnark_b200/launch.py).
        unimplemented!("use czk_sys::net::init(rank, n_parties, unique_id)")
    }
    fn is_init() -> bool { Self::n_parties() >= 1 }
    fn deinit() { with_ctx(|c| unsafe { ffi::czk_net_deinit(c) }) }
    fn reset_stats() { with_ctx(|c| unsafe { ffi::czk_net_reset_stats(c) }) }
    fn stats() -> Stats {
        let mut s = [0u64; 5];
        with_ctx(|c| check(c, "czk_net_stats", unsafe { ffi::czk_net_stats(c, s.as_mut_ptr()) }));
        Stats { bytes_sent: s[0] as usize, bytes_recv: s[1] as usize, broadcasts: s[2] as usize, to_king: s[3] as usize, from_king: s[4] as usize }
    }
    /// broadcast_bytes (multi.rs:145-174): every party contributes `bytes`, receives all of them in party order
    fn broadcast_bytes(bytes: &[u8]) -> Vec<Vec<u8>> {
        let n = Self::n_parties();
        let mut recv = vec![0u8; n * bytes.len()];
        with_ctx(|c| check(c, "czk_net_allgather_host", unsafe {
            ffi::czk_net_allgather_host(c, bytes.as_ptr() as *const _, recv.as_mut_ptr() as *mut _, bytes.len())
        }));
        recv.chunks(bytes.len().max(1)).map(|s| s.to_vec()).collect()
    }
    /// send_bytes_to_king (multi.rs:176-209) on host buffers: staged through the all-gather (O(1)-sized messages on this
    /// path; the bulk king traffic of GSZ stays on the device, czk_net_gather_to_king_dev)
    fn send_bytes_to_king(bytes: &[u8]) -> Option<Vec<Vec<u8>>> {
        let all = Self::broadcast_bytes(bytes);
        if Self::am_king() { Some(all) } else { None }
    }
    /// recv_bytes_from_king (multi.rs:211-242): the king's i-th message reaches party i
    fn recv_bytes_from_king(bytes: Option<Vec<Vec<u8>>>) -> Vec<u8> {
        let n = Self::n_parties();
        let me = Self::party_id();
        let flat: Vec<u8> = match bytes {
            Some(v) => v.into_iter().flatten().collect(),
            None => Vec::new(),
        };
        // equal-length messages (multi.rs:227): the king's concatenation is broadcast, every party keeps its slice
        let mut len = [flat.len() as u64 / n.max(1) as u64];
        let lens = Self::broadcast_bytes(unsafe { std::slice::from_raw_parts(len.as_mut_ptr() as *const u8, 8) });
        let m = u64::from_le_bytes([lens[0][0], lens[0][1], lens[0][2], lens[0][3], lens[0][4], lens[0][5], lens[0][6], lens[0][7]]) as usize;
        let mut buf = if Self::am_king() { flat } else { vec![0u8; n * m] };
        let all = Self::broadcast_bytes(&buf);
        buf = all[0].clone();
        buf[me * m..(me + 1) * m].to_vec()
    }
}

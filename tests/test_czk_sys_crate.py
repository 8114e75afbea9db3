"""The czk-sys crate (the Rust side of the boundary) cannot be compiled here - no rustc in the image - so
this checks what can be checked: src/ffi.rs is exactly what tools/gen_czk_sys.py generates from include/*.h, it declares
every symbol the headers export (and nothing else), the library really exports each one, and the hand-written wrapper
modules only call functions that exist."""
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _header_symbols():
    names = []
    for h in sorted((ROOT / "include").glob("*.h")):
        text = re.sub(r"/\*.*?\*/", " ", h.read_text(), flags=re.S)
        names += re.findall(r"CZK_API\s+[^;{}]+?\b(czk_\w+)\s*\(", text)
    return names


def test_ffi_is_generated_and_covers_every_exported_symbol():
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "gen_czk_sys.py"), "--check"])
    assert r.returncode == 0, "czk-sys/src/ffi.rs is stale: run python tools/gen_czk_sys.py"
    ffi = (ROOT / "czk-sys" / "src" / "ffi.rs").read_text()
    declared = re.findall(r"pub fn (czk_\w+)\(", ffi)
    hdr = _header_symbols()
    assert len(hdr) > 100 and sorted(declared) == sorted(hdr)
    assert len(set(declared)) == len(declared)


def test_wrapper_modules_call_only_declared_functions():
    ffi = (ROOT / "czk-sys" / "src" / "ffi.rs").read_text()
    declared = set(re.findall(r"pub fn (czk_\w+)\(", ffi))
    consts = set(re.findall(r"pub const (CZK_\w+)", ffi))
    used = set()
    for f in (ROOT / "czk-sys" / "src").glob("*.rs"):
        if f.name == "ffi.rs":
            continue
        text = f.read_text()
        used |= set(re.findall(r"ffi::(czk_\w+)\s*\(", text))
        for c in re.findall(r"ffi::(CZK_\w+)", text):
            assert c in consts, (f.name, c)
    assert used and used <= declared, used - declared
    # the safe layer reaches every part of the boundary: MSM, NTT, shares, GSZ, net, Groth16, Plonk
    for must in ("czk_msm_g1", "czk_msm_bases", "czk_ntt_fr", "czk_ntt_vec_batch", "czk_batch_open", "czk_beaver_batch_mul", "czk_gsz_open",
                 "czk_net_init", "czk_net_allgather_host", "czk_groth16_prove", "czk_plonk_prove_wiring"):
        assert must in used, must


def test_crate_files_exist():
    for f in ("Cargo.toml", "build.rs", "src/lib.rs", "src/ffi.rs", "src/msm.rs", "src/domain.rs", "src/shares.rs", "src/net.rs",
              "src/groth16.rs", "src/plonk.rs"):
        assert (ROOT / "czk-sys" / f).exists(), f

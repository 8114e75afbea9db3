"""CPU tier: the C-ABI library loads and exports every symbol include/czk.h declares; without a GPU the
product refuses to run (no CPU fallback); the host-side planners behave."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    out = set()
    for h in (ROOT / "include").glob("*.h"):
        out |= set(re.findall(r"CZK_API[^;(]*?\b(czk_\w+)\s*\(", h.read_text()))
    return out


def test_library_exports_every_declared_symbol(czk):
    lib = czk.load_library()
    names = declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/*.h but not exported"
    assert lib.czk_version().startswith(b"czk-b200")


def test_binding_covers_header(czk):
    from czk_b200 import binding

    missing = declared_symbols() - set(binding._SIGS) - set(binding._OPTIONAL_SIGS)
    assert not missing, missing


def test_no_cpu_fallback(czk):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(czk.CzkError) as e:
        czk.Context(0)
    assert e.value.code == 3  # CZK_ERR_NO_DEVICE


def test_product_does_not_import_oracle():
    for p in (ROOT / "collaborative-zksnark_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".hpp", ".cpp", ".h") and p.is_file():
            txt = p.read_text()
            assert "oracle" not in txt.replace("independent of oracle/", ""), p


def test_domain_params_host_side(czk, oracle):
    # computed on the host inside the library (no GPU needed): must equal Radix2EvaluationDomain::new
    for log_d in (1, 11, 21):
        a = czk.domain_params(log_d)
        b = oracle.domain_params(1 << log_d)
        for k in a:
            assert (a[k] == b[k]).all()

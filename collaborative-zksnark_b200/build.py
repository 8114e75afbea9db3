"""Build libczk_b200.so (sm_100a) in-tree with nvcc.  `python build.py [--force]`."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OBJ = HERE / "build"
LIB = HERE / "libczk_b200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CUFLAGS = ARCH + ["-O3", "-std=c++17", "-lineinfo", "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fPIC,-fvisibility=hidden",
                  "-I", str(HERE.parent / "include"), "--expt-relaxed-constexpr"] + os.environ.get("CZK_EXTRA_NVCC_FLAGS", "").split()
SOURCES = ["api.cu", "ntt.cu", "msm.cu", "msm_batched.cu", "fr_ops.cu", "shares.cu", "microbench.cu", "groth16.cu", "gsz.cu", "poly.cu", "plonk.cu", "serialize.cu", "pairing.cu", "fixed_base.cu"]


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    headers = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.hpp")) + list((HERE.parent / "include").glob("*.h"))
    srcs = [s for s in SOURCES if (CSRC / s).exists()]
    jobs = []
    for s in srcs:
        src = CSRC / s
        obj = OBJ / (src.stem + ".o")
        if force or _stale(obj, [src] + headers):
            jobs.append([NVCC] + CUFLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", str(src), "-o", str(obj)])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        for out in ex.map(run, jobs):
            if verbose and out:
                print(out)
    objs = [OBJ / (Path(s).stem + ".o") for s in srcs]
    if force or jobs or _stale(LIB, objs):
        cmd = [NVCC] + ARCH + ["-shared", "-ccbin", "/usr/bin/g++", "-o", str(LIB)] + [str(o) for o in objs] + \
              ["-cudart", "static", "-ldl", "-lpthread", "-lrt"]
        run(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

#!/usr/bin/env python3
"""Extract the BLS12-377 known-answer literals from the reference source tree.

TEST INFRASTRUCTURE.  Run once in the build container (where /root/reference
exists); the JSON it writes is committed so the tests never read
/root/reference at run time.

Sources (all relative to /root/reference/curves/bls12_377/src):
  fields/fr.rs:11-108   Fr: MODULUS, R, R2, INV, GENERATOR, TWO_ADIC_ROOT_OF_UNITY,
                        LARGE_SUBGROUP_ROOT_OF_UNITY, T, ... (Montgomery-form limbs)
  fields/fq.rs:11-118   Fq: same set
  fields/fq2.rs:13      NONRESIDUE = -5
  curves/g1.rs:17-51    b = 1, cofactor, generator (decimal)
  curves/g2.rs:14-86    b' , cofactor, generator (decimal)

These literals are the only known-answer values the reference holds for the
hot path: they pin the Montgomery representation
(R, R2, INV), the coset generator and the root-of-unity derivation.
"""
import json
import re
import sys
from pathlib import Path

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference") / "curves/bls12_377/src"


def strip_comments(src: str) -> str:
    return "\n".join(l for l in src.splitlines() if not l.strip().startswith("//"))


def bigints(src: str):
    """name -> int for every `const NAME: ... BigInteger([ ... ])` literal."""
    out = {}
    pat = re.compile(r"const\s+(\w+)\s*:\s*[^=]*=\s*(?:Some\()?\s*BigInteger\(\[(.*?)\]\)", re.S)
    for name, body in pat.findall(src):
        limbs = [int(t.replace("u64", "").replace("_", ""), 0) for t in re.findall(r"0x[0-9a-fA-F_]+|\d[\d_]*(?:u64)?", body)]
        out[name] = sum(l << (64 * i) for i, l in enumerate(limbs))
    for name, val in re.findall(r"const\s+(\w+)\s*:\s*u(?:32|64)\s*=\s*(?:Self::\w+\s*-\s*1|(\d+)(?:u32|u64)?)\s*;", src):
        if val:
            out[name] = int(val)
    for name, val in re.findall(r"const\s+(\w+)\s*:\s*Option<u32>\s*=\s*Some\((\d+)\)", src):
        out[name] = int(val)
    return out


def decimals(src: str):
    """NAME -> int for `pub const NAME: Fq = field_new!(Fq, "<decimal>")`."""
    out = {}
    for name, val in re.findall(r"const\s+(\w+)\s*:\s*\w+\s*=\s*field_new!\(\s*\w+\s*,\s*\"(-?\d+)\"\s*\)", src):
        out[name] = int(val)
    return out


def main():
    fr = bigints(strip_comments((REF / "fields/fr.rs").read_text()))
    fq = bigints(strip_comments((REF / "fields/fq.rs").read_text()))
    g1s = strip_comments((REF / "curves/g1.rs").read_text())
    g2s = strip_comments((REF / "curves/g2.rs").read_text())
    g1 = decimals(g1s)
    g2 = decimals(g2s)
    m = re.search(r"FQ_ZERO,\s*field_new!\(Fq,\s*\"(\d+)\"\)", g2s)
    g2["COEFF_B_C1"] = int(m.group(1))
    cof1 = [int(x, 16) for x in re.findall(r"0x[0-9a-f]+", re.search(r"COFACTOR: &'static \[u64\] = &\[(.*?)\];", g1s, re.S).group(1))]
    cof2 = [int(x, 16) for x in re.findall(r"0x[0-9a-f]+", re.search(r"COFACTOR: &'static \[u64\] = &\[(.*?)\];", g2s, re.S).group(1))]
    g1["COFACTOR"] = sum(l << (64 * i) for i, l in enumerate(cof1))
    g2["COFACTOR"] = sum(l << (64 * i) for i, l in enumerate(cof2))
    fq2 = decimals(strip_comments((REF / "fields/fq2.rs").read_text()))
    doc = {
        "_source": "curves/bls12_377/src/{fields/fr.rs,fields/fq.rs,fields/fq2.rs,curves/g1.rs,curves/g2.rs} @ reference commit 8cff2c2",
        "_note": "field literals are Montgomery-form integers exactly as written in the reference (little-endian u64 limbs folded into one int), decimal strings",
        "fr": {k: str(v) for k, v in sorted(fr.items())},
        "fq": {k: str(v) for k, v in sorted(fq.items())},
        "fq2": {k: str(v) for k, v in sorted(fq2.items())},
        "g1": {k: str(v) for k, v in sorted(g1.items())},
        "g2": {k: str(v) for k, v in sorted(g2.items())},
    }
    out = Path(__file__).with_name("bls12_377_constants.json")
    out.write_text(json.dumps(doc, indent=1) + "\n")
    print("wrote", out, {k: len(v) for k, v in doc.items() if isinstance(v, dict)})


if __name__ == "__main__":
    main()

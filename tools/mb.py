import sys
sys.path.insert(0, '.')
import czk_b200
ctx = czk_b200.Context(0)
names = {15: "fq_mul_two_chains", 13: "imad_wide_carry_clean", 14: "imad_wide+iadd3x", 12: "fq_mul_regmod", 3: "fq_mul", 0: "imad_wide", 5: "imad_wide_carry", 7: "imad_lo", 1: "imad_lo_hi_pair", 6: "iadd3_x", 2: "fr_mul", 4: "g1_madd"}
for kind, name in names.items():
    for bps, thr in ((1, 128), (4, 128), (8, 256)):
        iters = 4000 if kind in (0, 1, 5, 6, 7, 13, 14) else (400 if kind in (2, 3, 12, 15) else 60)
        ops, ms = ctx.microbench(kind, bps, thr, iters)
        print(f"{name:18s} blocks/SM={bps} threads={thr}: {ops:.3e} ops/s  ({ms:.3f} ms)  per-SM-clk={ops/148/1.965e9:.2f}", flush=True)

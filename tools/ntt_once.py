"""Device time of the NTT at one size: `python tools/ntt_once.py [log_d]` (CUDA events on the context stream)."""
import sys
sys.path.insert(0, '.')
import numpy as np
import torch
import czk_b200
log_d = int(sys.argv[1]) if len(sys.argv) > 1 else 21
ctx = czk_b200.Context(0)
n = 1 << log_d
rng = np.random.Generator(np.random.PCG64(1))
def rand():
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    return a
vecs = [ctx.vec_from(rand()) for _ in range(6)]
stream = torch.cuda.ExternalStream(ctx.stream, device=f"cuda:{ctx.device}")
names = {0: "fft", 1: "ifft", 2: "coset_fft", 3: "coset_ifft", 4: "ifft+coset_fft"}
for op in (0, 1, 2, 3, 4):
    for cnt in (1, 6):
        ctx.ntt_batch(vecs[:cnt], log_d, op)
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(reps):
                ctx.ntt_batch(vecs[:cnt], log_d, op)
            e1.record(stream)
        ctx.sync()
        ms = e0.elapsed_time(e1) / reps
        print(f"ntt 2^{log_d} {names[op]:15s} x{cnt}: {ms:.3f} ms  ({ms / cnt / (2 if op == 4 else 1):.3f} ms per transform)", flush=True)

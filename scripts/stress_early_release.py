"""Stress of the early-release chain: many forwards of random batch sizes, queued without synchronisation in between,
each compared bit for bit with the same kernels issued one by one through the public entries (plain stream order).
    python scripts/stress_early_release.py [iterations]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import koemorph_b200 as K
from oracle import koemorph_oracle as O
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = torch.device("cuda", 0)
w = O.make_weights(1242, 30, style="stress")
m = K.SequentialDualStreamModel().to(dev).eval()
m.load_state_dict(O.model_state_dict(w), strict=True)
m.set_compression_layer(torch.from_numpy(w["compression.weight"]), torch.from_numpy(w["compression.bias"]))
m.precision = "bf16"
N = 900
g = torch.Generator(device=dev).manual_seed(1)
audio = 0.1 * torch.randn(N, 136000, device=dev, generator=g)
audio[::5] *= 1e-3
eg = torch.randn(N, 264, device=dev, generator=g)
fe = m._frontend(dev)
power, fmax = fe.power(audio, 533, 257)
want, _, _ = m._core_windows([power], [fmax], 0, N, 257, 1, 1, 257, m._check_egemaps(eg, N, dev), False)
torch.cuda.synchronize()
rng = np.random.default_rng(0)
bad = 0
pending = []
for it in range(iters):
    n = int(rng.integers(149, N + 1))
    lo = int(rng.integers(0, N - n + 1))
    out = m(audio[lo:lo + n], egemaps=eg[lo:lo + n])["blendshapes"]
    pending.append((lo, n, out))
    if len(pending) == 8:      # eight forwards in flight, then check
        torch.cuda.synchronize()
        for lo_, n_, o_ in pending:
            if not torch.equal(o_.reshape(n_, 1, 52), want[lo_:lo_ + n_]):
                bad += 1
                print(f"MISMATCH: clips [{lo_}, {lo_ + n_})  max |d| = {(o_.reshape(n_, 1, 52) - want[lo_:lo_ + n_]).abs().max().item():.3e}")
        pending = []
torch.cuda.synchronize()
print(f"{iters} forwards of 149..{N} clips: {bad} mismatches")
sys.exit(1 if bad else 0)

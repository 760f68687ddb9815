import sys, os
sys.path.insert(0, "/root/repo")
import torch
from koemorph_b200.features.mel_frontend import LogMelFrontend
B, hop, n_frames = 512, 533, 257
audio = 0.1 * torch.randn(B, 136448, device="cuda")
fe = LogMelFrontend.get("cuda")
db, fmax = fe.power(audio, hop, n_frames)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for _ in range(3):
    fe.power(audio, hop, n_frames, out=(db, fmax))
ts = []
for _ in range(10):
    ev[0].record()
    for _ in range(10):
        fe.power(audio, hop, n_frames, out=(db, fmax))
    ev[1].record(); torch.cuda.synchronize()
    ts.append(ev[0].elapsed_time(ev[1]) / 10)
ts.sort()
print(os.environ.get("KOE_K1_DEBUG"), "median %.1f us" % (ts[5] * 1e3))

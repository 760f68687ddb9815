"""Phase timeline of the tcgen05 core inside a streaming hop (ring mode, 4096 streams): CTA 0's first two windows."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import koemorph_oracle as O
import koemorph_b200 as K
from koemorph_b200 import _lib
from koemorph_b200.streaming import StreamingEngine
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
w = O.make_weights(1235, 30, style="stress")
m = K.SequentialDualStreamModel().cuda().eval()
m.load_state_dict(O.model_state_dict(w), strict=True)
m.precision = "bf16"
eng = StreamingEngine(m, S)
eng.set_egemaps(torch.randn(S, 264, device="cuda"))
hops = [0.1 * torch.randn(S, m.hop_length, device="cuda") for _ in range(8)]
for i in range(m.mel_sequence_length + 8):
    eng.step(hops[i % 8])
torch.cuda.synchronize()
dbg = torch.zeros(128 + 2 * 148, dtype=torch.int64, device="cuda")
lib = _lib.load()
lib.koe_debug_set_tc_timestamps.argtypes = [C.c_void_p]
lib.koe_debug_set_tc_timestamps(dbg.data_ptr())
eng.step(hops[0]); torch.cuda.synchronize()
lib.koe_debug_set_tc_timestamps(None)
d = dbg.cpu().tolist()
t0 = d[0]
span = np.array(d[128:128 + 296]).reshape(148, 2)
print("core kernel: first CTA start -> last CTA end %.1f us; per-CTA duration median %.1f us (%d windows per CTA)" % (
    (span[:, 1].max() - span[:, 0].min()) / 1e3, np.median(span[:, 1] - span[:, 0]) / 1e3, (S + 147) // 148))
names_s = ["start", "staged", "G1 done", "E1 done", "S/VT done", "E2/3 done", "PV done", "E4 done", "H1 done", "E5 done"]
names_m = ["go1", "G1 issued", "go2", "S/VT issued", "go3", "PV issued", "go4", "H1 issued"]
for wdw in range(2):
    print("window", wdw)
    ev = [(d[16 * wdw + i] - t0, "SIMT " + names_s[i]) for i in range(10)] + \
         [(d[16 * wdw + 10 + i] - t0, "SIMT   staging: " + n) for i, n in enumerate(
             ["dB reference done", "mel stage 0 staged", "mel stage 2 staged", "mel stages staged"]) if d[16 * wdw + 10 + i]] + \
         [(d[64 + 16 * wdw + i] - t0, "MMA  " + names_m[i]) for i in range(8)] + \
         [(d[32 + 16 * wdw + i] - t0, "TMA      issue " + n) for i, n in enumerate(
             ["H1 s0", "H1 s1", "H1 s2", "H1 s3", "G1 s0", "G1 last", "S/VT s0", "S/VT last", "mel s0", "mel last"])]
    prev = None
    for t, n in sorted(ev):
        print(f"  {t:8d} cyc  (+{0 if prev is None else t - prev:6d})  {n}")
        prev = t

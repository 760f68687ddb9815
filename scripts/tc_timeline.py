"""Phase timeline of the tcgen05 core kernel (CTA 0, first two windows) from clock64 stamps."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import koemorph_oracle as O
import koemorph_b200 as K
from koemorph_b200 import _lib
w = O.make_weights(1235, 30, style="stress")
m = K.SequentialDualStreamModel().cuda().eval()
m.load_state_dict(O.model_state_dict(w), strict=True)
m.precision = "bf16"
B = 512
a = 0.1 * torch.randn(B, 136000, device="cuda"); e = torch.randn(B, 264, device="cuda")
for _ in range(3): m(a, egemaps=e)
dbg = torch.zeros(128 + 2 * 148, dtype=torch.int64, device="cuda")
lib = _lib.load()
lib.koe_debug_set_tc_timestamps.argtypes = [C.c_void_p]
lib.koe_debug_set_tc_timestamps(dbg.data_ptr())
m(a, egemaps=e); torch.cuda.synchronize()
lib.koe_debug_set_tc_timestamps(None)
d = dbg.cpu().tolist()
t0 = d[0]
print('CTA 0: kernel entry %d, prologue done %d, first window staged %d, last window done %d (cycles, relative to window 0 SIMT start)' % tuple(d[120 + i] - d[0] for i in range(4)))
import numpy as np
span = np.array(d[128:128 + 296]).reshape(148, 2)
t_first = span[:, 0].min()
print('per-CTA wall clock (us after the first CTA started): start min/median/max %.1f %.1f %.1f, end min/median/max %.1f %.1f %.1f' % (
    *(np.percentile(span[:, 0] - t_first, [0, 50, 100]) / 1e3), *(np.percentile(span[:, 1] - t_first, [0, 50, 100]) / 1e3)))
print('  CTAs with 4 windows (blockIdx < 68): end median %.1f us; with 3 windows: %.1f us' % (np.median(span[:68, 1] - t_first) / 1e3, np.median(span[68:, 1] - t_first) / 1e3))
names_s = ["start", "staged", "G1 done", "E1 done", "S/VT done", "E2/3 done", "PV done", "E4 done", "H1 done", "E5 done"]
names_m = ["go1", "G1 issued", "go2", "S/VT issued", "go3", "PV issued", "go4", "H1 issued"]
for wdw in range(2):
    print("window", wdw)
    ev = [(d[16 * wdw + i] - t0, "SIMT " + names_s[i]) for i in range(10)] + \
         [(d[16 * wdw + 10 + i] - t0, "SIMT   staging: " + n) for i, n in enumerate(
             ["dB reference done", "mel stage 0 staged", "mel stage 2 staged", "mel stages staged"]) if d[16 * wdw + 10 + i]] + \
         [(d[64 + 16 * wdw + i] - t0, "MMA  " + names_m[i]) for i in range(8)] + \
         [(d[64 + 16 * wdw + 8 + i] - t0, f"MMA    H1 stage {i} arrived") for i in range(4)] + \
         [(d[64 + 16 * wdw + 12 + i] - t0, f"MMA    G1 stage {(0, 1, 2, 'last')[i]} arrived") for i in range(4)] + \
         [(d[32 + 16 * wdw + i] - t0, "TMA      issue " + n) for i, n in enumerate(
             ["H1 s0", "H1 s1", "H1 s2", "H1 s3", "G1 s0", "G1 last", "S/VT s0", "S/VT last", "mel s0", "mel last"])]
    prev = None
    for t, n in sorted(ev):
        print(f"  {t:8d} cyc  (+{0 if prev is None else t - prev:6d})  {n}")
        prev = t

# whole-launch time of the core kernel alone for the same batch (events around 10 launches), for scale
import ctypes
power, fmax = m._frontend(a.device).power(a, m.hop_length, 257)
eg = m._check_egemaps(e, B, a.device)
for _ in range(3):
    m._core_windows([power], [fmax], 0, B, 257, 1, 1, 257, eg, False)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    m._core_windows([power], [fmax], 0, B, 257, 1, 1, 257, eg, False)
e1.record(); torch.cuda.synchronize()
print("emotion + core kernels, 512 windows: %.1f us per call (clock %d MHz nominal 1965)" % (e0.elapsed_time(e1) * 100, 1965))

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
dbg = torch.zeros(128, dtype=torch.int64, device="cuda")
lib = _lib.load()
lib.koe_debug_set_tc_timestamps.argtypes = [C.c_void_p]
lib.koe_debug_set_tc_timestamps(dbg.data_ptr())
m(a, egemaps=e); torch.cuda.synchronize()
lib.koe_debug_set_tc_timestamps(None)
d = dbg.cpu().tolist()
t0 = d[0]
names_s = ["start", "staged", "G1 done", "E1 done", "S/VT done", "E2/3 done", "PV done", "E4 done", "H1 done", "E5 done"]
names_m = ["go1", "G1 issued", "go2", "S/VT issued", "go3", "PV issued", "go4", "H1 issued"]
for wdw in range(2):
    print("window", wdw)
    ev = [(d[16 * wdw + i] - t0, "SIMT " + names_s[i]) for i in range(10)] + \
         [(d[64 + 16 * wdw + i] - t0, "MMA  " + names_m[i]) for i in range(8)]
    prev = None
    for t, n in sorted(ev):
        print(f"  {t:8d} cyc  (+{0 if prev is None else t - prev:6d})  {n}")
        prev = t

"""Launch-level picture of one forward's chain (globaltimer stamps): when the core's CTAs start / end and when the emotion
kernel's CTAs (queued behind the core) start, finish their work and leave."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
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
n_emo = (B + 15) // 16 if os.environ.get('KOE_EMO_CLIPS', '16') == '16' else (B + 7) // 8
dbg = torch.zeros(128 + 2 * 148 + 3 * n_emo + 8, dtype=torch.int64, device="cuda")
lib = _lib.load()
lib.koe_debug_set_tc_timestamps.argtypes = [C.c_void_p]
lib.koe_debug_set_tc_timestamps(dbg.data_ptr())
for _ in range(3): m(a, egemaps=e)   # the stamps of the last forward stay; earlier ones keep the queue full
torch.cuda.synchronize()
lib.koe_debug_set_tc_timestamps(None)
d = np.array(dbg.cpu().tolist())
core = d[128:128 + 296].reshape(148, 2)
emo = d[128 + 296:128 + 296 + 3 * n_emo].reshape(n_emo, 3)
ph = d[128 + 296 + 3 * n_emo:]
t0 = core[:, 0].min()
pct = lambda x: "min %.1f median %.1f max %.1f" % tuple(np.percentile((x - t0) / 1e3, [0, 50, 100]))
print("core CTAs    start:", pct(core[:, 0]), "  end:", pct(core[:, 1]), "(us after the first core CTA started)")
print("emotion CTAs start:", pct(emo[:, 0]))
print("             work done:", pct(emo[:, 1]), "  left:", pct(emo[:, 2]))
print("emotion CTA work time (us): min %.1f median %.1f max %.1f" % tuple(np.percentile((emo[:, 1] - emo[:, 0]) / 1e3, [0, 50, 100])))
n4 = B % 148 if B % 148 else 148
for name, sl in (("CTAs with one window more (blockIdx < %d)" % n4, slice(0, n4)), ("the others", slice(n4, 148))):
    if core[sl].size:
        print("%-46s start: %s   end: %s" % (name, pct(core[sl, 0]), pct(core[sl, 1])))
order = np.argsort(core[:, 0])
print("blockIdx of the first 12 core CTAs to start:", order[:12].tolist(), " of the last 12:", order[-12:].tolist())

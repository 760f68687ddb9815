"""Time koe_emotion_stream alone for 512 / 4096 clips (KOE_EMO_CLIPS forces clips per CTA in experiment builds)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import koemorph_oracle as O
import koemorph_b200 as K
from koemorph_b200 import _lib
w = O.make_weights(1235, 30, style="stress")
m = K.SequentialDualStreamModel().cuda().eval()
m.load_state_dict(O.model_state_dict(w), strict=True)
kw = m.dual_stream_attention.kernel_weights(m._compression)
lib = _lib.load()
for B in (512, 4096):
    eg = torch.randn(B, 264, device="cuda")
    out = torch.empty(B, device="cuda")
    st = _lib.stream_ptr(eg.device)
    for _ in range(5):
        lib.koe_emotion_stream(C.byref(kw.struct), eg.data_ptr(), B, out.data_ptr(), st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(10):
        e0.record()
        for _ in range(20):
            lib.koe_emotion_stream(C.byref(kw.struct), eg.data_ptr(), B, out.data_ptr(), st)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 20 * 1e3)
    ts.sort()
    print(os.environ.get("KOE_EMO_CLIPS"), B, "clips: median %.1f us" % ts[5])

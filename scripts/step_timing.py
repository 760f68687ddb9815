"""Step timing of the batch forward (512 clips, bf16 core) with CUDA events; prints ms per step.  Env switches of the
library (read once per process) select experiment variants."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import koemorph_b200 as K
from oracle import koemorph_oracle as O
dev = torch.device("cuda", 0)
w = O.make_weights(1234, 30, style="init")
m = K.SequentialDualStreamModel().to(dev).eval()
m.load_state_dict(O.model_state_dict(w), strict=True)
m.set_compression_layer(torch.from_numpy(w["compression.weight"]), torch.from_numpy(w["compression.bias"]))
m.precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
audio = 0.1 * torch.randn(512, 136000, device=dev)
eg = torch.randn(512, 264, device=dev)
kept = torch.empty(20, 512, 1, 52, device=dev)
for i in range(5):
    m(audio, egemaps=eg, out=kept[i])
torch.cuda.synchronize()
ts = []
for rep in range(7):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        m(audio, egemaps=eg, out=kept[i])
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 20)
ts.sort()
print(f"step: median {ts[3]*1e3:.1f} us, min {ts[0]*1e3:.1f} us ({'no programmatic launch' if os.environ.get('KOE_NO_PDL') else 'default'})")

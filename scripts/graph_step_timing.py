import os, sys
sys.path.insert(0, "/root/repo")
import torch
import koemorph_b200 as K
from oracle import koemorph_oracle as O
dev = torch.device("cuda", 0)
w = O.make_weights(1234, 30, style="init")
m = K.SequentialDualStreamModel().to(dev).eval()
m.load_state_dict(O.model_state_dict(w), strict=True)
m.set_compression_layer(torch.from_numpy(w["compression.weight"]), torch.from_numpy(w["compression.bias"]))
m.precision = "bf16"
audio = 0.1 * torch.randn(512, 136000, device=dev)
eg = torch.randn(512, 264, device=dev)
kept = torch.empty(20, 512, 1, 52, device=dev)
for i in range(5):
    m(audio, egemaps=eg, out=kept[i])
torch.cuda.synchronize()
ref = kept[:5].clone()
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    m(audio, egemaps=eg, out=kept[0])
torch.cuda.current_stream().wait_stream(s)
kept.zero_()
with torch.cuda.graph(g):
    for i in range(20):
        m(audio, egemaps=eg, out=kept[i])
g.replay()
torch.cuda.synchronize()
print("graph result equals eager:", torch.equal(kept[:5], ref))
ts = []
for rep in range(7):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 20)
ts.sort()
print(f"graph step: median {ts[3]*1e3:.1f} us, min {ts[0]*1e3:.1f} us")

import sys, cProfile, pstats, time
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
t0 = time.perf_counter()
for i in range(200):
    m(audio, egemaps=eg, out=kept[i % 20])
t1 = time.perf_counter()
torch.cuda.synchronize()
print("host submit per call (GPU-bound loop)", (t1 - t0) / 200 * 1e6, "us")
# host-only cost: tiny batch so the GPU never back-pressures
a2, e2 = audio[:2].contiguous(), eg[:2].contiguous()
k2 = torch.empty(2, 1, 52, device=dev)
for i in range(5):
    m(a2, egemaps=e2, out=k2)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(500):
    m(a2, egemaps=e2, out=k2)
t1 = time.perf_counter()
torch.cuda.synchronize()
print("host cost per call (tiny batch)", (t1 - t0) / 500 * 1e6, "us")
pr = cProfile.Profile()
pr.enable()
for i in range(500):
    m(a2, egemaps=e2, out=k2)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)

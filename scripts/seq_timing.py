"""Timing of a sliding-window forward (n_out > 1: edge-variant frontend launches + EMA scan in the chain).
    python scripts/seq_timing.py [fps] [extra_hops] [clips]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import koemorph_b200 as K
from oracle import koemorph_oracle as O
fps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
extra = int(sys.argv[2]) if len(sys.argv) > 2 else 30
B = int(sys.argv[3]) if len(sys.argv) > 3 else 512
W = 256 if fps == 30 else 512
dev = torch.device("cuda", 0)
w = O.make_weights(1234, fps, style="init")
m = K.SequentialDualStreamModel(target_fps=fps, mel_sequence_length=W).to(dev).eval()
m.load_state_dict(O.model_state_dict(w), strict=True)
m.set_compression_layer(torch.from_numpy(w["compression.weight"]), torch.from_numpy(w["compression.bias"]))
m.precision = "bf16"
L = (W + extra) * m.hop_length
audio = 0.1 * torch.randn(B, L, device=dev)
eg = torch.randn(B, 264, device=dev)
n_out = m.num_output_frames(L)
for _ in range(3):
    m(audio, egemaps=eg)
torch.cuda.synchronize()
ts = []
for rep in range(7):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
        m(audio, egemaps=eg)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 10)
ts.sort()
print(f"{fps} fps, {B} clips x {n_out} output frames: median {ts[3]*1e3:.1f} us per forward")

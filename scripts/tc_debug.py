"""Bring-up check of the tcgen05 core: bf16 tensor path vs the fp32 CUDA-core path and the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import koemorph_oracle as O
import koemorph_b200 as K

torch.manual_seed(0)
w = O.make_weights(1235, 30, style="stress")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
audio, eg = O.make_inputs(5679, B, 136000, "speechlike")
m = K.SequentialDualStreamModel().cuda().eval()
m.load_state_dict(O.model_state_dict(w), strict=True)
m.set_compression_layer(torch.from_numpy(w["compression.weight"]), torch.from_numpy(w["compression.bias"]))
a, e = torch.from_numpy(audio).cuda(), torch.from_numpy(eg).cuda()
ref = m(a, return_attention=True, egemaps=e)
torch.cuda.synchronize()
print("fp32 ok", ref["blendshapes"].shape, flush=True)
m.precision = "bf16"
out = m(a, return_attention=True, egemaps=e)
torch.cuda.synchronize()
print("bf16 ran", flush=True)
for k in ("blendshapes", "mel_blendshapes", "mel_attention_weights"):
    d = (out[k] - ref[k]).abs()
    print(k, "max|d|", float(d.max()), "mean|d|", float(d.mean()), "ref max", float(ref[k].abs().max()))
print("ref sig", ref["mel_blendshapes"][0, 0, 14:20].tolist())
print("tc  sig", out["mel_blendshapes"][0, 0, 14:20].tolist())
print("ref attn", ref["mel_attention_weights"][0, 0, 0, :6].tolist())
print("tc  attn", out["mel_attention_weights"][0, 0, 0, :6].tolist())

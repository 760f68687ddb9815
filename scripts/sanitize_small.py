"""Small end-to-end run for compute-sanitizer (one tool per gpurun call): batch forward (warp-specialised log-mel kernel,
emotion stream, both cores), a sequence with edge variants, a few 30 fps and 60 fps streaming hops."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import koemorph_b200 as K
from koemorph_b200.streaming import StreamingEngine
from oracle import koemorph_oracle as O

for fps, W in ((30, 256), (60, 512)):
    w = O.make_weights(1235, fps, style="stress")
    m = K.SequentialDualStreamModel(target_fps=fps, mel_sequence_length=W).cuda().eval()
    m.load_state_dict(O.model_state_dict(w), strict=True)
    m.set_compression_layer(torch.from_numpy(w["compression.weight"]), torch.from_numpy(w["compression.bias"]))
    hop = m.hop_length
    audio, eg = O.make_inputs(7, 3, 136000 + 5 * hop + 17, "speechlike")
    a, e = torch.from_numpy(audio).cuda(), torch.from_numpy(eg).cuda()
    for prec in ("fp32", "bf16"):
        m.precision = prec
        out = m(a, egemaps=e, return_attention=True)["blendshapes"]
        torch.cuda.synchronize()
        print(fps, prec, tuple(out.shape), float(out.sum()))
    eng = StreamingEngine(m, 3)
    eng.set_egemaps(e)
    n_hops = W + 3
    for n in range(n_hops):
        o = eng.step(a[:, n * hop:(n + 1) * hop].contiguous())
    torch.cuda.synchronize()
    print(fps, "stream", float(o.sum()))
print("done")

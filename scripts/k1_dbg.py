import sys, os
sys.path.insert(0, "/root/repo")
import torch
from koemorph_b200.features.mel_frontend import LogMelFrontend
B, hop, n_frames = 512, 533, 257
audio = 0.1 * torch.randn(B, 136448, device="cuda")
fe = LogMelFrontend.get("cuda")
os.environ["KOE_K1_DEBUG"] = "1"
for _ in range(3):
    db, fmax = fe.power(audio, hop, n_frames)
torch.cuda.synchronize()
os.environ["KOE_K1_DEBUG"] = "2"
fe.power(audio, hop, n_frames)
torch.cuda.synchronize()

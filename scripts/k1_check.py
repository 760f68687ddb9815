"""Bring-up harness for the log-mel kernel (K1): parity against a float64 torch.stft reference on the GPU, then timing.

    python scripts/k1_check.py [n_clips]

Test infrastructure only (torch.stft is the checker, never the product path)."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from koemorph_b200.features.mel_frontend import LogMelFrontend

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
hop, n_frames = 533, 257
torch.manual_seed(0)
audio = 0.1 * torch.randn(B, 136448, device="cuda")
audio[1] *= torch.linspace(1e-4, 1.0, 136448, device="cuda")            # level ramp
t = torch.arange(136448, device="cuda") / 16000.0
audio[2] = 0.5 * torch.sin(2 * torch.pi * 440.0 * t)                      # pure tone: exercises the error floor
audio[3, :70000] = 0                                                      # silence then noise
fe = LogMelFrontend.get("cuda")
db, fmax = fe.power(audio, hop, n_frames)
torch.cuda.synchronize()

nref = min(B, 8)
fb = torch.from_numpy(fe.filterbank()).cuda().double()
win = torch.hann_window(1024, periodic=True, dtype=torch.float64, device="cuda")
S = torch.stft(audio[:nref].double(), 1024, hop, 1024, win, center=True, pad_mode="constant", return_complex=True)
mel = torch.einsum("mf,bft->btm", fb, S.abs() ** 2)[:, :n_frames]
ref = 10 * torch.log10(mel.clamp_min(1e-10))
got = db[:nref].double()
for b in range(nref):
    big = mel[b] > 1e-6 * mel[b].max()
    e_db = (got[b] - ref[b]).abs()[big].max().item()
    e_pw = ((10 ** (got[b] / 10)) - mel[b].clamp_min(1e-10)).abs().max().item() / max(mel[b].max().item(), 1e-10)
    e_mx = (fmax[b].double() - got[b].max(dim=1).values).abs().max().item()
    print(f"clip {b}: max dB err (within 60 dB of peak) {e_db:.3e}   abs power err / peak {e_pw:.3e}   fmax err {e_mx:.1e}")
    assert e_db < 8.7e-4 and e_pw < 2e-6 and e_mx == 0, "parity failure"
assert torch.isfinite(db).all()

# edge variants and a ragged frame count
db2, _ = fe.power(audio[:4, :100000], hop, 100, frame_offset=3, frame_step=2, lo_rel=-1, hi_rel=1)
torch.cuda.synchronize()

ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for _ in range(3):
    fe.power(audio, hop, n_frames, out=(db, fmax))
times = []
REP = 10  # back-to-back launches per timed region, so that host launch overhead stays off the critical path
for _ in range(20):
    ev[0].record()
    for _ in range(REP):
        fe.power(audio, hop, n_frames, out=(db, fmax))
    ev[1].record()
    torch.cuda.synchronize()
    times.append(ev[0].elapsed_time(ev[1]) / REP)
times.sort()
algo = B * (136000 * 4 + n_frames * 324)
print(f"K1: {B} clips x {n_frames} frames: median {times[10]*1e3:.1f} us, min {times[0]*1e3:.1f} us -> "
      f"{algo / times[10] / 1e6:.0f} GB/s algorithmic")

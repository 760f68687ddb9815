"""K1: cost of the window-edge frame pairs.  Times the same launch shape (512 clips x 128 pairs) with and without the
two pairs per clip that touch the zero padding (frames 0 and 256 of the 257-frame window)."""
import sys
sys.path.insert(0, "/root/repo")
import torch
from koemorph_b200.features.mel_frontend import LogMelFrontend

B, hop = 512, 533
audio = 0.1 * torch.randn(B, 136448 + 4 * hop, device="cuda")
fe = LogMelFrontend.get("cuda")


def run(n_frames, frame_offset):
    db, fmax = fe.power(audio, hop, n_frames, frame_offset=frame_offset)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for _ in range(3):
        fe.power(audio, hop, n_frames, frame_offset=frame_offset, out=(db, fmax))
    ts = []
    for _ in range(10):
        ev[0].record()
        for _ in range(10):
            fe.power(audio, hop, n_frames, frame_offset=frame_offset, out=(db, fmax))
        ev[1].record()
        torch.cuda.synchronize()
        ts.append(ev[0].elapsed_time(ev[1]) / 10)
    ts.sort()
    pairs = B * ((n_frames + 1) // 2)
    print(f"n_frames {n_frames} offset {frame_offset}: median {ts[5] * 1e3:.1f} us, {ts[5] * 1e6 / pairs:.3f} ns/pair")


run(256, 0)   # pair 0 of every clip reads the left padding
run(256, 1)   # all pairs interior
run(256, 2)
run(257, 0)   # the shipped shape: pair 0 and pair 128 (single frame) are edge pairs

# the bench shape: 136,000-sample clips, 257 frames (frame 255 also runs past the end of the clip: three edge pairs)
audio = audio[:, :136000].contiguous()
run(257, 0)

"""K1 (log-mel kernel) scheduling experiments: every variant must reproduce the shipped kernel bit for bit; then timing.

    python scripts/k1_variants.py [n_clips]

Variants are selected through koe_debug_k1_variant(group_warps, threads, stagger_cycles, late_store_mask):
  group_warps  16 = one barrier group (shipped); 8 / 4 = independent named-barrier groups of 8 / 4 warps, each with its own
               16 / 8 frames per mel phase (lane = frame, upper lanes idle), started `stagger` cycles apart
  threads      512 (16 warps per SM) or fewer (occupancy probe)
  late mask    warp classes (warp / 4) whose store phase runs after the NEXT iteration's FFT instead of before it
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from koemorph_b200 import _lib
from koemorph_b200.features.mel_frontend import LogMelFrontend

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
hop, n_frames = 533, 257
torch.manual_seed(0)
audio = 0.1 * torch.randn(B, 136000, device="cuda")
audio[1] *= torch.linspace(1e-4, 1.0, 136000, device="cuda")
t = torch.arange(136000, device="cuda") / 16000.0
audio[2] = 0.5 * torch.sin(2 * torch.pi * 440.0 * t)
audio[3, :70000] = 0
fe = LogMelFrontend.get("cuda")
lib = _lib.load()
lib.koe_debug_k1_variant.argtypes = [__import__("ctypes").c_int] * 2


def set_variant(order, ws=0):
    rc = lib.koe_debug_k1_variant(order, ws)
    assert rc == 0, lib.koe_last_error()


def timed(out):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for _ in range(3):
        fe.power(audio, hop, n_frames, out=out)
    ts = []
    for _ in range(15):
        ev[0].record()
        for _ in range(10):
            fe.power(audio, hop, n_frames, out=out)
        ev[1].record()
        torch.cuda.synchronize()
        ts.append(ev[0].elapsed_time(ev[1]) / 10)
    ts.sort()
    return ts[len(ts) // 2] * 1e3, ts[0] * 1e3


set_variant(0)
ref_db, ref_fm = fe.power(audio, hop, n_frames)
torch.cuda.synchronize()
ref_db, ref_fm = ref_db.clone(), ref_fm.clone()

# store_order: 2 bits per warp class (warp / 4): where the store phase of the previous iteration's rows sits --
# 0 = before the FFT, 1 = after the next pair's loads, 2 = between the FFT and the loads
variants = [("all store-first 0x00", 0x00), ("all after-loads 0x55", 0x55), ("all after-fft 0xAA", 0xAA),
            ("classes 1,3 after-loads 0x44", 0x44), ("classes 2,3 after-loads 0x50", 0x50),
            ("classes 1,3 after-fft 0x88", 0x88), ("0,1,2,1 -> 0x64", 0x64), ("class 3 after-loads 0x40", 0x40)]

variants = [(n, o, 0) for n, o in variants[:4]] + [("warp-specialised: 16 producers @96 regs + 8 consumers @48 regs (shipped)", 0, 1)]
out = (torch.empty_like(ref_db), torch.empty_like(ref_fm))
for name, order, ws in variants:
    set_variant(order, ws)
    out[0].fill_(float("nan"))
    out[1].fill_(float("nan"))
    fe.power(audio, hop, n_frames, out=out)
    torch.cuda.synchronize()
    same = torch.equal(out[0], ref_db) and torch.equal(out[1], ref_fm)
    med, mn = timed(out)
    print(f"{name:36s} median {med:7.1f} us  min {mn:7.1f} us  bit-identical {same}", flush=True)
    assert same, name
set_variant(0x44, 1)

# timing probe (results are not written): FFT + loads only, no CTA barrier / mel phase / store phase
set_variant(0x100)
med, mn = timed(out)
print(f"{'probe: FFT + loads only, free running':36s} median {med:7.1f} us  min {mn:7.1f} us")
set_variant(0x44, 1)

# timing probe: the warp-specialised kernel with the producers NOT waiting for the consumers (results invalid): what full
# decoupling of producers and consumers could buy at most
set_variant(0x200, 1)
med, mn = timed(out)
print(f"{'probe: ws kernel, producers never wait for the consumers':60s} median {med:7.1f} us  min {mn:7.1f} us")
set_variant(0x44, 1)

#!/usr/bin/env python
"""Secondary workloads of BASELINE.json (configs[2], [3], [4]); bench.py covers configs[1].

    python scripts/bench_configs.py [--configs c3,c4,c5] [--streams 4096] [--corpus-clips 1000000] [--precision bf16]
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_configs.py --configs c5

Prints one JSON line per config (rank 0).  Synthetic data, random-init weights (oracle.make_weights, test infra).
  c3  60 fps mode (hop 266, 512-frame window, K = 515): 512 clips x 8.5 s, one output frame per clip, both cores
  c4  streaming: S concurrent streams, rings pre-filled with 8.5 s, then 300 steps of one hop per stream; per-step
      latency measured on the host (submit -> outputs complete), p50 / p99
  c5  corpus sweep: N clips sharded by clip over the ranks, 512-clip chunks from a small pool of device buffers,
      one all_gather of the (N, 52) outputs at the end
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

import koemorph_b200 as K
from koemorph_b200 import _lib
from koemorph_b200.parallel import shard_range, gather_outputs
from koemorph_b200.streaming import StreamingEngine
from oracle import koemorph_oracle as O  # weight generator only


def make_model(fps, dev, precision):
    w = O.make_weights(1234, fps, style="init")
    m = K.SequentialDualStreamModel(target_fps=fps, mel_sequence_length=256 if fps == 30 else 512).to(dev).eval()
    m.load_state_dict(O.model_state_dict(w), strict=True)
    m.set_compression_layer(torch.from_numpy(w["compression.weight"]), torch.from_numpy(w["compression.bias"]))
    m.precision = precision
    return m


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def c3(dev, args):
    m = make_model(60, dev, args.precision)
    B = 512
    audio = 0.1 * torch.randn(B, 136000, device=dev)
    eg = torch.randn(B, 264, device=dev)
    ms = timed(lambda: m(audio, egemaps=eg), 20)
    m.precision = "fp32" if args.precision == "bf16" else "bf16"
    ms_other = timed(lambda: m(audio, egemaps=eg), 20)
    return {"config": "c3: 60 fps (hop 266, 512-frame window, K=515), 512 x 8.5 s clips, 1 frame/clip",
            "dtype": args.precision, "ms_per_step": ms, "value": B * 8.5 / (ms * 1e-3), "unit": "audio-s/s", "n_gpus": 1,
            "other_precision": {"precision": m.precision, "ms_per_step": ms_other, "value": B * 8.5 / (ms_other * 1e-3)}}


def c4(dev, args):
    m = make_model(30, dev, args.precision)
    S = args.streams
    eng = StreamingEngine(m, S)
    eng.native = not args.stream_python_driver
    eng.set_egemaps(torch.randn(S, 264, device=dev))
    hops = [0.1 * torch.randn(S, m.hop_length, device=dev) for _ in range(8)]
    for i in range(m.mel_sequence_length + 8):          # fill the 8.5 s context, then a few emitting warm-up steps
        eng.step(hops[i % 8])
    torch.cuda.synchronize()
    lat = []
    for i in range(args.stream_steps):
        t0 = time.perf_counter()
        out = eng.step(hops[i % 8])
        torch.cuda.synchronize()
        lat.append((time.perf_counter() - t0) * 1e3)
        assert out is not None
    lat.sort()
    p = lambda q: lat[min(len(lat) - 1, int(q * len(lat)))]
    return {"config": f"c4: streaming, {S} concurrent streams, stride 1 hop (533 samples), 8.5 s context", "dtype": args.precision,
            "latency_ms": {"p50": p(0.50), "p90": p(0.90), "p99": p(0.99), "max": lat[-1], "steps": len(lat)},
            "value": S * (m.hop_length / 16000.0) / (p(0.50) * 1e-3), "unit": "audio-s/s (new audio per wall second at p50)",
            "real_time_factor_p99": (m.hop_length / 16000.0 * 1e3) / p(0.99), "n_gpus": 1,
            "driver": "koe_stream_push (one native call per step)" if eng.native else "six calls per step from Python",
            "note": "latency = host submit of one hop for every stream -> all outputs complete on the device (host sync)"}


def c5(dev, args, rank, world):
    m = make_model(30, dev, args.precision)
    N = args.corpus_clips
    lo, hi = shard_range(N, rank, world)
    B = 512
    pool = [(0.1 * torch.randn(B, 136000, device=dev), torch.randn(B, 264, device=dev)) for _ in range(4)]
    out = torch.empty(hi - lo, 1, 52, device=dev)
    def run():
        i = 0
        for c0 in range(0, hi - lo, B):
            n = min(B, hi - lo - c0)
            a, e = pool[i & 3]
            out[c0:c0 + n] = m(a[:n], egemaps=e[:n])["blendshapes"]
            i += 1
        return gather_outputs(out, N) if world > 1 else out
    run()  # warm-up pass over the whole shard
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    full = run()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    assert full.shape[0] == N
    return {"config": f"c5: corpus sweep, {N} synthetic 8.5 s clips sharded by clip over {world} GPU(s), 512-clip chunks "
                      f"(pool of 4 device buffers), one all_gather of (N,1,52) at the end", "dtype": args.precision,
            "seconds": ms * 1e-3, "value": N * 8.5 / (ms * 1e-3), "unit": "audio-s/s", "n_gpus": world,
            "clips_per_second": N / (ms * 1e-3)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c3,c4,c5")
    ap.add_argument("--stream-python-driver", action="store_true",
                    help="c4: issue the step call by call from Python instead of through koe_stream_push")
    ap.add_argument("--streams", type=int, default=4096)
    ap.add_argument("--stream-steps", type=int, default=300)
    ap.add_argument("--corpus-clips", type=int, default=1000000)
    ap.add_argument("--precision", default="bf16", choices=["fp32", "bf16"])
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    with torch.no_grad():
        for name in args.configs.split(","):
            if name in ("c3", "c4") and rank != 0:
                continue
            res = {"c3": lambda: c3(dev, args), "c4": lambda: c4(dev, args), "c5": lambda: c5(dev, args, rank, world)}[name]()
            if rank == 0:
                res["data"] = "synthetic"
                print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Benchmark of the hot path: batched audio -> blendshape inference (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision fp32|tf32|bf16]

One step = SequentialDualStreamModel.forward over 512 synthetic 8.5 s clips (16 kHz, 30 fps, one output
frame per clip) per GPU.  BASELINE.json quotes configs[1] in "fp32 and bf16": the line's value / dtype are the bf16
tensor-core path (log-mel frontend in fp32 either way), `other_precision` carries the all-fp32 path of the same run.  Prints ONE JSON line (see the task contract): `value` is the whole-job
audio-seconds per second with inputs resident in HBM; `e2e` is the same metric through the host-buffer API
(pinned host audio -> H2D -> kernels -> D2H); `roofline` describes the dominant kernel; `cpu_baseline` is the
oracle port of the reference's CPU forward timed on this box's host cores.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CLIPS_PER_GPU = 512
CLIP_SAMPLES = 136000          # 8.5 s at 16 kHz
CLIP_SECONDS = 8.5
METRIC = "audio-seconds/sec (30fps, 8.5s ctx)"
UNIT = "audio-s/s"


# ----------------------------------------------------------------------------- CPU baseline (oracle port)
_G = {}


def _cpu_init(clips_per_worker):
    """Per-process set-up (not timed): weights, synthetic clips, one warm-up forward."""
    import torch
    from oracle import koemorph_oracle as O
    torch.set_num_threads(1)
    w = O.make_weights(1234, 30, style="init")
    audio, eg = O.make_inputs(1000 + os.getpid() % 1000, clips_per_worker, CLIP_SAMPLES, "noise")
    O.forward_sequence(w, audio[:1], eg[:1])
    _G.update(w=w, audio=audio, eg=eg, fwd=O.forward_sequence)


def _cpu_pass(reps):
    t0 = time.perf_counter()
    for _ in range(reps):
        _G["fwd"](_G["w"], _G["audio"], _G["eg"])
    return time.perf_counter() - t0


def _cpu_pool(clips_per_worker):
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    # one single-threaded worker per core: without this every worker's BLAS/OpenMP runtime spawns a thread per core and
    # the oversubscribed pool runs ~50x slower (which would flatter the GPU numbers)
    saved = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS")}
    os.environ.update({k: "1" for k in saved})
    try:
        pool = mp.get_context("spawn").Pool(cores, initializer=_cpu_init, initargs=(clips_per_worker,))
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    pool.map(_cpu_pass, [0] * cores, chunksize=1)  # make sure every worker is up
    return pool, cores


_CPU_NOTE = ("oracle.forward_sequence (per-clip Python loop like the reference, one single-threaded process per core); "
             "eGeMAPS extraction excluded (synthetic input); mel stage is the librosa restatement, not librosa")


def cpu_baseline(target_seconds=12.0):
    """Oracle port of SequentialDualStreamModel.forward on all host cores, on a bounded sample of the workload."""
    clips_per_worker = 4
    pool, cores = _cpu_pool(clips_per_worker)
    with pool:
        t_probe = max(pool.map(_cpu_pass, [1] * cores, chunksize=1))
        reps = max(1, int(target_seconds / max(t_probe, 1e-3)))
        t0 = time.perf_counter()
        pool.map(_cpu_pass, [reps] * cores, chunksize=1)
        wall = time.perf_counter() - t0
    n = cores * clips_per_worker * reps
    return {"value": n * CLIP_SECONDS / wall, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} clips of 8.5 s ({cores} processes x {clips_per_worker} clips x {reps} passes, "
                      f"{wall:.1f} s wall) through " + _CPU_NOTE}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port: the reference itself needs
    librosa/opensmile and cannot travel to the GPU box).  Rank 0 only; other ranks exit without work."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    clips_per_worker = 4
    pool, cores = _cpu_pool(clips_per_worker)
    # one step = the GPU arm's step: ~512 clips, i.e. every worker forwards its 4 clips `reps` times
    reps = max(1, -(-CLIPS_PER_GPU // (cores * clips_per_worker)))
    times = []
    with pool:
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            pool.map(_cpu_pass, [reps] * cores, chunksize=1)
            if i >= args.warmup:
                times.append(time.perf_counter() - t0)
    n = cores * clips_per_worker * reps
    step = sum(times) / len(times)
    value = n * CLIP_SECONDS / step
    sample = f"{n} clips of 8.5 s per step ({cores} processes x {clips_per_worker} clips x {reps} passes) through " + _CPU_NOTE
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": "512 x 8.5 s clips @16 kHz per GPU, 30 fps, 256x80 mel context, 1 frame/clip "
                                   "(BASELINE.json configs[1])", "step_sample": f"{n} clips per step"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _loop(self):
        nv = self._nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                break
            time.sleep(0.001)

    def start(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    import koemorph_b200 as K
    from koemorph_b200 import _lib
    from koemorph_b200.infer import HostPipeline
    from oracle import koemorph_oracle as O  # weights generator + cpu_baseline only (never on the timed path)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: koemorph_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from koemorph_b200.infer import bind_host_thread_to_gpu_node
    numa_node = bind_host_thread_to_gpu_node(local) if world > 1 else None   # before any pinned allocation
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    w = O.make_weights(1234, 30, style="init")
    model = K.SequentialDualStreamModel().to(dev).eval()
    model.load_state_dict(O.model_state_dict(w), strict=True)
    model.set_compression_layer(torch.from_numpy(w["compression.weight"]), torch.from_numpy(w["compression.bias"]))
    model.precision = args.precision

    B = CLIPS_PER_GPU
    g = torch.Generator(device=dev).manual_seed(5678 + rank)
    audio = 0.1 * torch.randn(B, CLIP_SAMPLES, device=dev, generator=g)       # 278.5 MB > 126 MB L2
    eg = torch.randn(B, 264, device=dev, generator=g)
    kept = torch.empty(args.steps, B, 1, 52, device=dev) if world > 1 else None
    gathered = torch.empty(world * args.steps, B, 1, 52, device=dev) if world > 1 else None
    step_no = [0]

    fe = model._frontend(dev)
    n_frames = model.window_frames + 1
    k1_events = []

    def step(record):
        # the dominant kernel (log-mel power) is bracketed by events inside the timed region for the roofline;
        # model.forward launches exactly the same kernels (see sequential_dual_stream_model.py)
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            power, fmax = fe.power(audio, model.hop_length, n_frames)
            e1.record()
            k1_events.append((e0, e1))
        else:
            power, fmax = fe.power(audio, model.hop_length, n_frames)
        # multi-GPU: the step's frames land directly in their slot of the buffer that is gathered at the end
        slot = None
        if kept is not None:
            slot = kept[step_no[0] % kept.shape[0]]
            step_no[0] += 1
        out, _, _ = model._core_windows([power], [fmax], 0, B, n_frames, 1, 1, n_frames, eg, False, out=slot)
        return out

    def gather_all():
        # the path's only collective (north_star: "NCCL over NVLink used only for the final output gather"): every
        # rank's results of all the steps, once, inside the timed region
        if kept is not None:
            dist.all_gather_into_tensor(gathered, kept)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # sanity: the decomposed step equals the public forward
    ref = model(audio[:8].contiguous(), egemaps=eg[:8].contiguous())["blendshapes"]
    for _ in range(max(args.warmup, 3)):
        out = step(False)
    gather_all()  # warm-up covers the collective too (first use sets up NCCL's channels for this size)
    step_no[0] = 0
    barrier()
    assert torch.equal(out[:8], ref), "bench step diverges from SequentialDualStreamModel.forward"

    sampler = ClockSampler(local)
    _lib.reset_launch_count()
    barrier()
    sampler.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        step(True)
    gather_all()
    t1.record()
    barrier()
    clocks = sampler.stop()
    launches = _lib.launch_count()
    elapsed_ms = t0.elapsed_time(t1)
    k1_ms = sum(a.elapsed_time(b) for a, b in k1_events) / len(k1_events)
    tt = torch.tensor([elapsed_ms], device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    elapsed_ms = float(tt.item())
    value = world * B * CLIP_SECONDS * args.steps / (elapsed_ms * 1e-3)

    # ---- the other precision of BASELINE.json configs[1] ("fp32 and bf16"), same workload, device-resident, fewer steps ----
    other = None
    if not args.no_other_precision:
        other_name = "bf16" if args.precision == "fp32" else "fp32"
        model.precision = other_name
        for _ in range(3):
            step(False)
        barrier()
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_other = max(3, min(args.steps, 20))
        o0.record()
        for _ in range(n_other):
            step(False)
        o1.record()
        barrier()
        to = torch.tensor([o0.elapsed_time(o1)], device=dev)
        if world > 1:
            dist.all_reduce(to, op=dist.ReduceOp.MAX)
        other = {"precision": other_name, "value": world * B * CLIP_SECONDS * n_other / (float(to.item()) * 1e-3),
                 "unit": UNIT, "ms_per_step": float(to.item()) / n_other, "steps": n_other}
        model.precision = args.precision

    # ---- end to end through the host-buffer API (pinned host -> H2D -> kernels -> D2H) ----
    e2e_value, e2e_pcm_value, e2e_steps = None, None, 0
    if not args.no_e2e:
        pipe = HostPipeline(model, chunk_clips=64)
        audio_h = torch.empty(B, CLIP_SAMPLES, dtype=torch.float32, pin_memory=True)
        audio_h.copy_(audio)
        eg_h = torch.empty(B, 264, dtype=torch.float32, pin_memory=True)
        eg_h.copy_(eg)
        out_h = torch.empty(B, 1, 52, dtype=torch.float32, pin_memory=True)
        e2e_steps = max(3, min(args.steps, 20))

        def time_pipe(a_h):
            ts = time.perf_counter()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(e2e_steps):
                pipe(a_h, eg_h, out_h)
            e1.record()
            barrier()
            ms = max(e0.elapsed_time(e1), (time.perf_counter() - ts) * 1e3)
            te = torch.tensor([ms], device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            return world * B * CLIP_SECONDS * e2e_steps / (float(te.item()) * 1e-3)

        for _ in range(2):
            pipe(audio_h, eg_h, out_h)
        expect = step(False).clone()  # (multi-GPU steps write into the gather buffer, which the legs above have reused)
        barrier()
        assert torch.allclose(out_h, expect.cpu(), atol=1e-7), "host pipeline diverges from the device path"
        e2e_value = time_pipe(audio_h)

        # the same clips handed over as 16-bit PCM (the samples' format in a WAV file): half the PCIe bytes.  Reported
        # beside the float32 number, not instead of it; checked against the float path fed the same quantised samples.
        pcm_h = torch.empty(B, CLIP_SAMPLES, dtype=torch.int16, pin_memory=True)
        pcm_h.copy_((audio.clamp(-1.0, 32767.0 / 32768.0) * 32768.0).round().to(torch.int16))
        audio_h.copy_(pcm_h.to(torch.float32) / 32768.0)
        pipe(audio_h, eg_h, out_h)
        want = out_h.clone()
        for _ in range(2):
            pipe(pcm_h, eg_h, out_h)
        barrier()
        assert torch.equal(out_h, want), "PCM16 host path differs from the float path on the same samples"
        e2e_pcm_value = time_pipe(pcm_h)

    if rank == 0:
        peaks, peak_src = None, "fallback"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
            peak_src = "measured"
        except Exception:
            pass
        hbm_peak = float(peaks["hbm_gbs"]) if peaks else 6650.0
        # algorithmic bytes of one log-mel launch: audio read once + mel power and frame maxima written once
        k1_bytes = B * (CLIP_SAMPLES * 4 + n_frames * 80 * 4 + n_frames * 4)
        achieved = k1_bytes / (k1_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                traffic = json.load(f).get("logmel_power_kernel_dram_bytes_per_launch")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": {"fp32": "f32", "tf32": "tf32", "bf16": "bf16"}[args.precision],
            "data": "synthetic",
            "config": {"workload": "512 x 8.5 s clips @16 kHz per GPU, 30 fps, 256x80 mel context, 1 frame/clip "
                                   "(BASELINE.json configs[1])", "clips_per_gpu": B, "parallelism": f"clip-shard x{world}",
                       "precision": args.precision, "collective": f"one all_gather of the ({args.steps},B,1,52) results at the end of the timed region" if world > 1 else "none",
                       "l2": "inputs 278.5 MB per GPU > 126 MB L2, no flush needed"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * (CLIP_SAMPLES + 264) * 4,
                    "d2h_bytes_per_step": B * 52 * 4, "steps": e2e_steps,
                    "api": "koemorph_b200.infer.HostPipeline (pinned host tensors, 64-clip chunks, 2 streams)",
                    "host_numa_node_rank0": numa_node,
                    "pcm16_input": {"value": e2e_pcm_value, "unit": UNIT,
                                    "h2d_bytes_per_step": B * (CLIP_SAMPLES * 2 + 264 * 4),
                                    "note": "same API, audio handed over as int16 PCM and converted on the device "
                                            "(bit-identical results); the headline e2e above is float32 host audio"}},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "logmel_power_kernel", "bound": "hbm", "achieved": achieved, "peak": hbm_peak,
                         "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
                         "peak_source": peak_src, "kernel_ms": k1_ms, "share_of_step": k1_ms / (elapsed_ms / args.steps),
                         "algorithmic_bytes_per_launch": k1_bytes},
        }
        if other is not None:
            line["other_precision"] = other
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="koemorph_b200", choices=["koemorph_b200", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["fp32", "tf32", "bf16"],
                    help="core kernel: bf16 = tcgen05 tensor path (bf16 operands, fp32 accumulation; default), fp32 = CUDA-core FMA; "
                         "the other one is timed too and reported under other_precision")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-buffer leg")
    ap.add_argument("--no-other-precision", action="store_true", help="skip the secondary (bf16 / fp32) leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

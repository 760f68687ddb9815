"""Host-buffer entry point: pinned host audio in, host blendshapes out, copies overlapped with compute.

This is the call an application makes when its audio lives in host memory (the reference's own calling
convention is host numpy -> per-clip librosa, src/model/simplified_dual_stream_model.py:184-229).
Clips are independent, so the batch is cut into chunks that flow through a two-stage pipeline on two CUDA
streams: while chunk i runs the kernels, chunk i+1 is crossing PCIe.

The audio may also be handed over as int16 PCM, the format the samples have in a WAV file (the reference's loaders turn
it into float32 / 32768 on the host, src/data/io.py:71): PCIe then carries half the bytes and the conversion runs on the
device (koe_pcm16_to_float), with bit-identical results.
"""
from __future__ import annotations

from typing import Optional

import torch

from .features.mel_frontend import pcm16_to_float


def bind_host_thread_to_gpu_node(device_index: int) -> Optional[int]:
    """Pin the calling process to the CPUs of the NUMA node its GPU hangs off, so that pinned buffers allocated afterwards
    (first touch) and the copy threads live next to the PCIe root of that GPU -- with one process per GPU on a two-socket
    box, cross-socket pinned memory otherwise caps the aggregate host->device bandwidth.  Returns the node, or None when
    the topology is not exposed (containers without /sys access): nothing is changed then."""
    import os
    try:
        bdf = torch.cuda.get_device_properties(device_index).pci_bus_id if hasattr(
            torch.cuda.get_device_properties(device_index), "pci_bus_id") else None
        if bdf is None:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            bdf = pynvml.nvmlDeviceGetPciInfo(h).busId
            bdf = bdf.decode() if isinstance(bdf, bytes) else bdf
        bdf = bdf.lower()
        if len(bdf.split(":")[0]) == 8:          # nvml prints an 8-digit domain, sysfs uses 4
            bdf = bdf[4:]
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


class HostPipeline:
    def __init__(self, model, chunk_clips: int = 64):
        self.model = model
        self.chunk = int(chunk_clips)
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("HostPipeline needs the model on a CUDA device")
        self._streams = [torch.cuda.Stream(self.device) for _ in range(2)]
        self._bufs = {}

    def _buffers(self, slot: int, n: int, L: int):
        key = (slot, L)
        if key not in self._bufs:
            self._bufs[key] = (torch.empty(self.chunk, L, dtype=torch.float32, device=self.device),
                               torch.empty(self.chunk, 264, dtype=torch.float32, device=self.device))
        a, e = self._bufs[key]
        return a[:n], e[:n]

    def _pcm_buffer(self, slot: int, n: int, L: int):
        key = (slot, L, "pcm16")
        if key not in self._bufs:
            self._bufs[key] = torch.empty(self.chunk, L, dtype=torch.int16, device=self.device)
        return self._bufs[key][:n]

    @torch.no_grad()
    def __call__(self, audio_host: torch.Tensor, egemaps_host: torch.Tensor,
                 out_host: Optional[torch.Tensor] = None) -> torch.Tensor:
        """audio_host (B, L) float32 -- or int16 PCM, see the module docstring -- and egemaps_host (B, 264) in (ideally
        pinned) host memory -> (B, T_out, 52) host tensor.  Synchronises before returning."""
        if audio_host.is_cuda or egemaps_host.is_cuda:
            raise ValueError("HostPipeline takes host tensors; call the model directly for device tensors")
        if audio_host.dtype not in (torch.float32, torch.int16):
            raise ValueError(f"audio_host must be float32 or int16 PCM, got {audio_host.dtype}")
        pcm = audio_host.dtype == torch.int16
        B, L = audio_host.shape
        eg = egemaps_host.reshape(B, 264)
        model = self.model
        sequential = hasattr(model, "num_output_frames")
        n_out = model.num_output_frames(L) if sequential else None
        shape = (B, n_out, 52) if sequential else (B, 52)
        if out_host is None:
            out_host = torch.empty(shape, dtype=torch.float32, pin_memory=True)
        # every chunk runs the STATELESS part of the forward and writes its rows into one device buffer; the model's
        # cross-call smoothing state is touched once, after the loop, on the caller's stream -- chunking must not change the
        # result (SimplifiedDualStreamModel.forward blends a call with the previous call's output, and consecutive chunks
        # hold different clips)
        full = torch.empty(shape, dtype=torch.float32, device=self.device)
        cur = torch.cuda.current_stream(self.device)
        for s in self._streams:
            s.wait_stream(cur)
        for ci, c0 in enumerate(range(0, B, self.chunk)):
            n = min(self.chunk, B - c0)
            s = self._streams[ci & 1]
            a_dev, e_dev = self._buffers(ci & 1, n, L)
            with torch.cuda.stream(s):
                if pcm:
                    p_dev = self._pcm_buffer(ci & 1, n, L)
                    p_dev.copy_(audio_host[c0:c0 + n], non_blocking=True)
                    pcm16_to_float(p_dev, a_dev)
                else:
                    a_dev.copy_(audio_host[c0:c0 + n], non_blocking=True)
                e_dev.copy_(eg[c0:c0 + n], non_blocking=True)
                model._forward_frames(a_dev, e_dev, False, out=full[c0:c0 + n])
                if sequential:
                    out_host[c0:c0 + n].copy_(full[c0:c0 + n], non_blocking=True)
        for s in self._streams:
            cur.wait_stream(s)
        if sequential:
            # what SequentialDualStreamModel.forward leaves behind (reference :99,136): the last frame of every clip
            model.prev_blendshapes = full[:, -1] if model.use_temporal_smoothing else None
        else:
            out_host.copy_(model.apply_temporal_smoothing(full), non_blocking=True)
        cur.synchronize()
        return out_host

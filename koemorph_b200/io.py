"""The data formats either side of the path: the sliding-window contract of the
reference's ``SequentialKoeMorphDataset`` (``src/data/sequential_dataset.py:181-206``) as index arithmetic, and the
``{"timestamp", "blendshapes"}`` JSON frames the reference reads as ground truth (``README.md:95-100``) and emits from
its real-time loop over a file or UDP (``scripts/rt.py:209-231``).  Host-side glue only: no compute."""
from __future__ import annotations

import json
import socket
from typing import IO, Iterable, Iterator, List, Optional, Sequence, Tuple

import torch


def window_grid(num_frames: int, window_frames: int = 256, stride_frames: int = 1, hop_length: int = 533
                ) -> List[Tuple[int, int, int, int]]:
    """(start_frame, end_frame, start_sample, end_sample) of every full window of a recording of ``num_frames``
    label frames -- ``num_windows = (num_frames - window_frames) // stride_frames + 1`` (sequential_dataset.py:182),
    window i = frames [i*stride, i*stride + window) = samples [start_frame*hop, end_frame*hop) (:186-192).
    These are exactly the windows ``SequentialDualStreamModel.forward`` slides over one long clip."""
    if window_frames < 1 or stride_frames < 1 or hop_length < 1:
        raise ValueError("window_frames, stride_frames and hop_length must be positive")
    n = (num_frames - window_frames) // stride_frames + 1
    out = []
    for i in range(max(0, n)):
        s = i * stride_frames
        out.append((s, s + window_frames, s * hop_length, (s + window_frames) * hop_length))
    return out


def frame_records(blendshapes, fps: float, t0: float = 0.0) -> Iterator[dict]:
    """(T, 52) -> one ``{"timestamp": seconds, "blendshapes": [52 floats]}`` record per frame; frame i carries
    ``t0 + (i + 1) / fps`` (the README's example: 0.0333, 0.0667, ... at 30 fps)."""
    x = torch.as_tensor(blendshapes).detach().float().cpu()
    if x.dim() != 2 or x.shape[1] != 52:
        raise ValueError(f"blendshapes must be (T, 52), got {tuple(x.shape)}")
    for i, row in enumerate(x.tolist()):
        yield {"timestamp": t0 + (i + 1) / float(fps), "blendshapes": row}


def write_jsonl(blendshapes, fps: float, fh: IO[str], t0: float = 0.0) -> int:
    """One JSON line per frame (the reference's file mode, rt.py:224-231).  Returns the number of frames written."""
    n = 0
    for rec in frame_records(blendshapes, fps, t0):
        fh.write(json.dumps(rec) + "\n")
        n += 1
    return n


def read_jsonl(lines: Iterable[str]) -> Tuple[torch.Tensor, torch.Tensor]:
    """Inverse of write_jsonl: (timestamps (T,), blendshapes (T, 52)); blank lines are skipped."""
    ts, rows = [], []
    for line in lines:
        line = line.strip()
        if not line:
            continue
        rec = json.loads(line)
        if len(rec["blendshapes"]) != 52:
            raise ValueError("a frame must carry 52 coefficients")
        ts.append(float(rec["timestamp"]))
        rows.append(rec["blendshapes"])
    return torch.tensor(ts, dtype=torch.float64), torch.tensor(rows, dtype=torch.float32).reshape(-1, 52)


class BlendshapeStreamer:
    """The reference's ``BlendshapeStreamer`` (scripts/rt.py:173-238) for the two self-contained modes: "udp" (one JSON
    datagram per frame) and "file" (JSON lines).  OSC needs python-osc, which is not part of this path."""

    def __init__(self, output_mode: str = "udp", host: str = "127.0.0.1", port: int = 9001,
                 output_file: Optional[str] = None):
        self.output_mode, self.host, self.port = output_mode, host, port
        self._sock = self._fh = None
        if output_mode == "udp":
            self._sock = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
        elif output_mode == "file":
            if not output_file:
                raise ValueError("output_file required for file mode")
            self._fh = open(output_file, "w")
        else:
            raise ValueError(f"Unknown output mode: {output_mode}")

    def send(self, blendshapes: Sequence[float], timestamp: float) -> None:
        row = torch.as_tensor(blendshapes).detach().float().cpu().reshape(-1).tolist()
        if len(row) != 52:
            raise ValueError("a frame must carry 52 coefficients")
        msg = json.dumps({"timestamp": float(timestamp), "blendshapes": row})
        if self._sock is not None:
            self._sock.sendto(msg.encode("utf-8"), (self.host, self.port))
        else:
            self._fh.write(msg + "\n")
            self._fh.flush()

    def close(self) -> None:
        if self._sock is not None:
            self._sock.close()
        if self._fh is not None:
            self._fh.close()
        self._sock = self._fh = None


# ---- window slicing of a long recording (SequentialKoeMorphDataset._process_file_pair, sequential_dataset.py:157-206) ------
def align_recording(n_samples: int, n_label_frames: int, hop_length: int) -> Tuple[int, int]:
    """(samples kept, label frames kept): when the label count is more than one frame off ``n_samples // hop`` the
    reference trims BOTH to the shorter one (sequential_dataset.py:169-180); otherwise nothing is trimmed."""
    expected = n_samples // hop_length
    if abs(n_label_frames - expected) > 1:
        n = min(n_label_frames, expected)
        return min(n_samples, n * hop_length), n
    return n_samples, n_label_frames


def recording_windows(audio: torch.Tensor, blendshapes: torch.Tensor, window_frames: int = 256, stride_frames: int = 1,
                      hop_length: int = 533):
    """All full windows of one recording, as zero-copy strided VIEWS on the tensors' own device (no gather kernel, no
    host loop): ``audio`` (L,) and ``blendshapes`` (T, 52) -> ``{"audio": (N, window_frames * hop), "blendshapes":
    (N, window_frames, 52), "start_frames": (N,)}``.  Window i covers label frames [i*stride, i*stride + W) and samples
    [start*hop, end*hop); windows whose audio would run past the (aligned) recording are dropped, like the reference's
    size check (sequential_dataset.py:182-196).  Feeding ``out["audio"]`` to ``SequentialDualStreamModel.forward`` gives
    one frame per window -- the same frames the model produces when it slides over the whole recording itself."""
    if audio.dim() != 1 or blendshapes.dim() != 2:
        raise ValueError(f"audio must be (L,) and blendshapes (T, C), got {tuple(audio.shape)}, {tuple(blendshapes.shape)}")
    if window_frames < 1 or stride_frames < 1 or hop_length < 1:
        raise ValueError("window_frames, stride_frames and hop_length must be positive")
    n_samples, n_frames = align_recording(audio.shape[0], blendshapes.shape[0], hop_length)
    audio, blendshapes = audio[:n_samples], blendshapes[:n_frames]
    w_samples = window_frames * hop_length
    n = (n_frames - window_frames) // stride_frames + 1
    # the reference yields a window only if its audio slice is complete
    while n > 0 and (n - 1) * stride_frames * hop_length + w_samples > n_samples:
        n -= 1
    n = max(n, 0)
    if n == 0:
        return {"audio": audio.new_zeros((0, w_samples)), "blendshapes": blendshapes.new_zeros((0, window_frames, blendshapes.shape[1])),
                "start_frames": torch.zeros(0, dtype=torch.long, device=audio.device)}
    a = audio.as_strided((n, w_samples), (stride_frames * hop_length * audio.stride(0), audio.stride(0)))
    b = blendshapes.as_strided((n, window_frames, blendshapes.shape[1]),
                               (stride_frames * blendshapes.stride(0), blendshapes.stride(0), blendshapes.stride(1)))
    return {"audio": a, "blendshapes": b,
            "start_frames": torch.arange(n, device=audio.device, dtype=torch.long) * stride_frames}

"""CPU tests of the host side: weight folding algebra, module surface, C-ABI exports (no compute calls)."""
import ctypes
import math
import os
import re

import numpy as np
import pytest
import torch

from oracle import koemorph_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _emulate_kernels(cw, x259, eg, k_mel):
    """What the CUDA kernels compute from the folded pack, written with torch on the CPU (float64)."""
    t = {k: v.double() if v.is_floating_point() else v for k, v in cw.tensors.items()}
    B = x259.shape[0]
    z = x259.double() @ t["wc_t"][:k_mel] + t["bc"]
    enc = torch.nn.functional.layer_norm(z, (256,), t["ln_g"], t["ln_b"], 1e-5)        # (B, 80, 256)
    s = enc @ t["qk_t"][:, :224]                                                          # (B, 80, 224)
    p = torch.softmax(s, dim=1)
    v = enc @ t["wv_t"] + t["bv"]                                                         # (B, 80, 256)
    o = torch.zeros(B, 28, 256, dtype=torch.float64)
    for h in range(8):
        ph = p[:, :, h * 28:(h + 1) * 28].transpose(1, 2)                                 # (B, 28, 80)
        o[:, :, 32 * h:32 * h + 32] = ph @ v[:, :, 32 * h:32 * h + 32]
    y_m = torch.sigmoid(torch.relu(o @ t["wa_t"] + t["ba"]) @ t["w2"] + cw.struct.b2)      # (B, 28)
    ze = torch.nn.functional.layer_norm(eg.double() @ t["we1_t"][:eg.shape[1]] + t["be1"], (256,), t["eln_g"],
                                        t["eln_b"], 1e-5)
    y_e = torch.sigmoid(torch.relu(ze @ t["we2_t"] + t["be2"]) @ t["w2"] + cw.struct.b2)   # (B,)
    sig = torch.zeros(B, 52, dtype=torch.float64)
    sig[:, t["mouth_idx"].long()] = y_m
    sig[:, t["expr_idx"].long()] = y_e[:, None]
    attn = p.reshape(B, 80, 8, 28).mean(2).transpose(1, 2)
    return torch.clamp(t["coef"] * sig, 0, 1), sig, attn


@pytest.mark.parametrize("fps,style", [(30, "init"), (30, "stress"), (60, "stress")])
def test_folding_matches_oracle(fps, style):
    from koemorph_b200.model.folding import fold
    w = O.make_weights(11, fps, style=style)
    mel_seq = 256 if fps == 30 else 512
    rng = np.random.default_rng(3)
    B = 3
    long_t = rng.uniform(0, 1, (B, mel_seq, 80)).astype(np.float32)
    short_t = rng.uniform(0, 1, (B, 3, 80)).astype(np.float32)
    eg = rng.standard_normal((B, 264)).astype(np.float32)
    emo = O.compress_egemaps(eg, w, torch.float64)
    ref = O.dual_stream_core(w, long_t, short_t, emo, mel_sequence_length=mel_seq, return_attention=True,
                             dtype=torch.float64)
    sd = {k[len("dual_stream_attention."):]: torch.from_numpy(v) for k, v in w.items()
          if k.startswith("dual_stream_attention.")}
    comp = {"weight": torch.from_numpy(w["compression.weight"]), "bias": torch.from_numpy(w["compression.bias"])}
    cw = fold(sd, 8, 1.0, "cpu", comp)
    assert cw.struct.k_mel == mel_seq + 3 and cw.struct.k_mel_pad % 16 == 0 and cw.struct.emo_in == 264
    x = torch.cat([torch.from_numpy(long_t).transpose(1, 2), torch.from_numpy(short_t).transpose(1, 2)], dim=2)
    out, sig, attn = _emulate_kernels(cw, x, torch.from_numpy(eg), mel_seq + 3)
    # folded weights are stored in float32: agreement is bounded by that rounding, not by the algebra
    assert (sig - ref["sigmoid"]).abs().max() < 2e-6
    assert (out - ref["blendshapes"]).abs().max() < 1e-7
    assert (attn - ref["mel_attention_weights"]).abs().max() < 1e-6
    # standalone core: no compression folded in, emotion input is 256-D
    cw2 = fold(sd, 8, 1.0, "cpu", None)
    assert cw2.struct.emo_in == 256
    out2, _, _ = _emulate_kernels(cw2, x, emo.float(), mel_seq + 3)
    assert (out2 - ref["blendshapes"]).abs().max() < 1e-7


def test_state_dict_contract():
    """SURVEY.md section 8 a-W: parameter names and shapes of the reference."""
    import koemorph_b200 as K
    m = K.SequentialDualStreamModel()
    sd = m.state_dict()
    w = O.make_weights(1, 30)
    expect = {k: v.shape for k, v in w.items() if not k.startswith("compression.")}
    assert {k: tuple(v.shape) for k, v in sd.items()} == {k: tuple(s) for k, s in expect.items()}
    m.load_state_dict(O.model_state_dict(w), strict=True)
    m60 = K.SequentialDualStreamModel(target_fps=60, mel_sequence_length=512)
    assert m60.hop_length == 266 and m60.state_dict()["dual_stream_attention.mel_channel_encoder.weight"].shape == (256, 515)
    assert m.hop_length == 533 and m.window_samples == 256 * 533
    assert [m.num_output_frames(n) for n in (136000, 136448 + 533, 320000)] == [1, 2, 345]
    assert K.MOUTH_INDICES == O.MOUTH_INDICES and K.EXPRESSION_INDICES == O.EXPRESSION_INDICES
    assert len(K.ARKIT_BLENDSHAPES) == 52 and K.ARKIT_BLENDSHAPES[17] == "jawOpen" and K.ARKIT_BLENDSHAPES[51] == "tongueOut"


def test_no_cpu_fallback():
    import koemorph_b200 as K
    m = K.SimplifiedDualStreamModel()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 16000), egemaps=torch.zeros(1, 264))
    core = K.DualStreamCrossAttention()
    with pytest.raises(RuntimeError, match="CUDA"):
        core(torch.zeros(1, 256, 80), torch.zeros(1, 3, 80), torch.zeros(1, 256))
    with pytest.raises(RuntimeError, match="real-time"):
        m.process_audio_frame_realtime(np.zeros(533, np.float32))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "koemorph_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f"{f} imports the oracle"


def test_c_abi_exports_every_declared_symbol():
    from koemorph_b200 import _lib
    header = open(os.path.join(ROOT, "include", "koemorph_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(koe_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations parsed"
    if not os.path.exists(_lib.LIB_PATH):
        from koemorph_b200.build import build
        build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"library does not export {missing}"
    assert set(_lib.exported_symbols()) <= set(declared)
    assert lib.koe_version() >= 100


def test_numa_binding_helper_is_a_no_op_without_topology():
    """bind_host_thread_to_gpu_node must never raise: without a GPU / sysfs topology it returns None and leaves the
    affinity mask alone."""
    from koemorph_b200.infer import bind_host_thread_to_gpu_node
    before = os.sched_getaffinity(0)
    assert bind_host_thread_to_gpu_node(0) in (None, 0, 1, 2, 3)
    if not __import__("torch").cuda.is_available():
        assert os.sched_getaffinity(0) == before


def test_ctypes_mirrors_have_the_size_of_the_c_structs():
    """The argument blocks cross the C ABI by pointer: a ctypes mirror that drifts from include/koemorph_b200.h would
    scribble over the fields behind the drift (the library reports its own sizeof for exactly this check)."""
    import ctypes as C
    from koemorph_b200 import _lib
    lib = _lib.load()
    for which, mirror in enumerate((_lib.FrontendConfig, _lib.LogmelArgs, _lib.CoreWeightsStruct, _lib.StreamArgs,
                                    _lib.ForwardArgs)):
        assert lib.koe_sizeof_struct(which) == C.sizeof(mirror), mirror.__name__
    assert lib.koe_sizeof_struct(99) == -1


def test_public_header_is_plain_c():
    """include/koemorph_b200.h is the contract a cgo / JNI / ctypes binding compiles against: it must parse as C99 on
    its own (no C++ types, no CUDA or torch headers)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler here")
    header = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "koemorph_b200.h")
    res = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", header],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr

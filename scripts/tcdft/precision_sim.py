"""Precision of a two-stage (32 x 32) DFT whose GEMMs run on split-fp16 operands with fp32 accumulation,
against the float64 oracle, measured with the criteria of tests/test_gpu_parity.py::test_mel_power_vs_float64_oracle.
CPU only (numpy emulation: rounding to 11 significant bits, float32 matmul)."""
import sys
import numpy as np
import scipy.fft
sys.path.insert(0, ".")
from oracle import koemorph_oracle as O


def rbits(x, bits):
    """round float32 array to `bits` significant bits (no range limit)"""
    x = np.asarray(x, np.float32)
    m, e = np.frexp(x)
    return np.ldexp(np.round(m * (1 << bits)) / (1 << bits), e).astype(np.float32)


def split(x, bits, terms):
    out, r = [], np.asarray(x, np.float32)
    for _ in range(terms):
        h = rbits(r, bits)
        out.append(h)
        r = (r - h).astype(np.float32)
    return out


def gemm_split(A, B, bits, terms_a, terms_b, order):
    """sum of the products a_i b_j with i + j < order, fp32 accumulation; A (M, K), B (K, N)"""
    As, Bs = split(A, bits, terms_a), split(B, bits, terms_b)
    acc = np.zeros((A.shape[0], B.shape[1]), np.float32)
    # small terms first is what a chained accumulate could do; here hi*hi last
    pairs = [(i, j) for i in range(terms_a) for j in range(terms_b) if i + j < order]
    for i, j in sorted(pairs, key=lambda p: -(p[0] + p[1])):
        acc = (acc + As[i] @ Bs[j]).astype(np.float32)
    return acc


def frames_of(y, hop):
    y = np.pad(y, 512)
    n = 1 + (len(y) - 1024) // hop
    return np.lib.stride_tricks.as_strided(y, (n, 1024), (y.strides[0] * hop, y.strides[0]))


def power_tc(fr, bits=11, ta=2, tb=2, order=2):
    hann = O.hann_window().astype(np.float32)
    xw = (fr * hann).astype(np.float32)                       # (F, 1024), n = 32 n1 + n2
    F_ = xw.shape[0]
    x = xw.reshape(F_, 32, 32)                                # [f, n1, n2]
    k = np.arange(32)
    ang = 2 * np.pi * np.outer(k, k) / 32
    C1 = np.concatenate([np.cos(ang), -np.sin(ang)], 0)      # (64 (k1,c), 32 n1)
    A = x.transpose(0, 2, 1).reshape(F_ * 32, 32)            # rows (f, n2), K = n1
    Y = gemm_split(A, C1.T.astype(np.float64), bits, ta, tb, order)      # (F*32, 64)
    Y = Y.reshape(F_, 32, 2, 32)                              # [f, n2, c, k1]
    Yr, Yi = Y[:, :, 0, :], Y[:, :, 1, :]
    n2 = np.arange(32)[:, None]; k1 = np.arange(32)[None, :]
    tw = np.exp(-2j * np.pi * n2 * k1 / 1024)
    twr, twi = tw.real.astype(np.float32), tw.imag.astype(np.float32)
    Zr = (Yr * twr - Yi * twi).astype(np.float32)            # [f, n2, k1]
    Zi = (Yr * twi + Yi * twr).astype(np.float32)
    # stage 2: rows (f, k1), K = (n2, c), N = (k2, c)
    A2 = np.concatenate([Zr.transpose(0, 2, 1), Zi.transpose(0, 2, 1)], 2).reshape(F_ * 32, 64)   # K = [re n2 | im n2]
    G = np.zeros((64, 64))
    G[:32, :32] = np.cos(ang); G[32:, :32] = np.sin(ang)     # Re X = sum Zr cos + Zi sin
    G[:32, 32:] = -np.sin(ang); G[32:, 32:] = np.cos(ang)    # Im X = -Zr sin + Zi cos
    X = gemm_split(A2, G, bits, ta, tb, order).reshape(F_, 32, 2, 32)     # [f, k1, c, k2]
    P = (X[:, :, 0, :] ** 2 + X[:, :, 1, :] ** 2).astype(np.float32)     # [f, k1, k2], k = k1 + 32 k2
    P = P.transpose(0, 2, 1).reshape(F_, 1024)
    return P[:, :513]


def power_fp32(fr):
    hann = O.hann_window().astype(np.float32)
    X = scipy.fft.rfft((fr * hann).astype(np.float32), axis=1)
    assert X.dtype == np.complex64
    return (X.real ** 2 + X.imag ** 2).astype(np.float32)


def check(P, ref, fb):
    mel = (P.astype(np.float32) @ fb.T.astype(np.float32)).astype(np.float64)
    db = 10 * np.log10(np.maximum(mel, 1e-10))
    ref_db = 10 * np.log10(np.maximum(ref, 1e-10))
    big = ref > 1e-6 * ref.max()
    e1 = np.abs(db - ref_db)[big].max()
    e2 = np.abs(mel - np.maximum(ref, 1e-10)).max() / max(ref.max(), 1e-10)
    # the log-mel criterion (|d| <= 1e-4 |ref| + 2e-5 in (dB + 80) / 80 units after the clip reference and clamp)
    def norm(d):
        d = d - d.max()
        return (np.maximum(d, -80) + 80) / 80
    a, b = norm(db), norm(ref_db)
    e3 = (np.abs(a - b) - 1e-4 * np.abs(b)).max()
    return e1, e2, e3


if __name__ == "__main__":
    fb = O._fb(16000, 1024, 80, 80.0, 8000.0)
    print("limits: dB(big) < 8.7e-4, abs/peak <= 2e-6, logmel excess <= 2e-5")
    for kind in ["noise", "speechlike", "level_step", "sine", "silence_burst"]:
        audio, _ = O.make_inputs(31, 3, 136000, kind)
        for b in range(1):
            ref = O.melspectrogram(audio[b], hop_length=533, exact=True).T
            fr = frames_of(audio[b], 533)[: ref.shape[0]]
            rows = [("fp32 fft", power_fp32(fr)),
                    ("fp16 2x2 order2 (3 products)", power_tc(fr, 11, 2, 2, 2)),
                    ("fp16 2x2 order3 (4 products)", power_tc(fr, 11, 2, 2, 3)),
                    ("bf16 3x3 order3 (6 products)", power_tc(fr, 8, 3, 3, 3))]
            for name, P in rows:
                e1, e2, e3 = check(P, ref, fb)
                print(f"{kind:14s} {name:30s} dB(big) {e1:.2e}  abs/peak {e2:.2e}  logmel excess {e3:.2e}")

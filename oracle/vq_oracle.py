"""ORACLE (test infrastructure only): Python face of oracle/vq_oracle.c plus a numpy restatement of the
reference formula used to pin it (muse/modeling_maskgit_vqgan.py:303-324,342-348)."""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_SO = _DIR / "_build" / "libvq_oracle.so"
_lib = None


def build():
    """Compiles vq_oracle.c with gcc (called by __graft_entry__.build(); building the checker is not using it)."""
    if not _SO.exists() or _SO.stat().st_mtime < (_DIR / "vq_oracle.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(_DIR), "_build/libvq_oracle.so"], check=True, capture_output=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(str(_SO))
        _lib.vq_oracle_argmin.restype = None
        _lib.vq_oracle_argmin.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] * 3 + [ctypes.c_void_p]
    return _lib


def argmin(z: np.ndarray, codebook: np.ndarray):
    """z [n, D] fp32, codebook [ncodes, D] fp32 -> (ids int64 [n], dmin fp32 [n]) with the pinned arithmetic."""
    z = np.ascontiguousarray(z, dtype=np.float32)
    cb = np.ascontiguousarray(codebook, dtype=np.float32)
    n, D = z.shape
    ids = np.empty(n, dtype=np.int64)
    dmin = np.empty(n, dtype=np.float32)
    ws = np.empty(cb.shape[0], dtype=np.float32)
    _load().vq_oracle_argmin(z.ctypes.data, cb.ctypes.data, ids.ctypes.data, dmin.ctypes.data, n, cb.shape[0], D,
                             ws.ctypes.data)
    return ids, dmin


def distances_numpy(z: np.ndarray, codebook: np.ndarray) -> np.ndarray:
    """compute_distances (:303-316) restated with numpy: (|z|^2 + |e|^2) - 2 z e^T, fp32 throughout."""
    z = z.astype(np.float32)
    cb = codebook.astype(np.float32)
    zn = (z ** np.float32(2.0)).sum(axis=1, keepdims=True, dtype=np.float32)
    en = (cb.T ** np.float32(2.0)).sum(axis=0, keepdims=True, dtype=np.float32)
    return (zn + en) + np.float32(-2.0) * (z @ cb.T)


def nchw_to_rows(z_nchw: np.ndarray) -> np.ndarray:
    """(B,C,H,W) -> (B*H*W, C): the permute(0,2,3,1).reshape(-1, C) of :276,:305."""
    b, c, h, w = z_nchw.shape
    return np.ascontiguousarray(np.transpose(z_nchw, (0, 2, 3, 1))).reshape(b * h * w, c)


def codebook_entry_nchw(ids: np.ndarray, codebook: np.ndarray) -> np.ndarray:
    """get_codebook_entry (:318-324): ids (B, T) -> (B, C, sqrt T, sqrt T)."""
    b, t = ids.shape
    s = int(round(t ** 0.5))
    return np.transpose(codebook[ids].reshape(b, s, s, -1), (0, 3, 1, 2))


def soft_code(z: np.ndarray, codebook: np.ndarray, temp: float = 1.0, expo_noise: np.ndarray | None = None):
    """get_soft_code (:327-340): soft = softmax(-d / temp) over the codebook; code = argmin d (stochastic=False) or
    multinomial(soft, 1) restated as argmax soft / q for q ~ Exp(1) (what torch.multinomial computes for one draw)."""
    d = distances_numpy(z, codebook)
    x = -d / np.float32(temp)
    x = x - x.max(axis=1, keepdims=True)
    e = np.exp(x, dtype=np.float32)
    soft = e / e.sum(axis=1, keepdims=True, dtype=np.float32)
    if expo_noise is None:
        ids, _ = argmin(z, codebook)
    else:
        ids = np.argmax(soft / expo_noise.astype(np.float32), axis=1).astype(np.int64)
    return soft, ids

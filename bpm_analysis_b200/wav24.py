"""24-bit PCM WAV files without scipy's full expansion.

``scipy.io.wavfile.read`` (bpm_analysis.py:1014) cannot memory-map 3-byte samples; it reads the whole
data chunk and expands it to int32 (the 24 bits in the upper three bytes).  The decimate-first front
end only ever looks at one frame in ``ds`` (:1033), so the file is mapped as bytes here and the kept
frames are expanded straight out of the page cache by ``bpm_host_gather_s24`` (libbpm_host.so).
Anything this parser does not recognise (RF64, big-endian RIFX, a truncated data chunk, other
encodings) is left to scipy: :func:`map_s24` returns None and the caller falls back.
"""
from __future__ import annotations

import ctypes as C
import struct
from typing import Optional, Tuple

import numpy as np

from .classifier import load_host_library

WAVE_FORMAT_PCM, WAVE_FORMAT_EXTENSIBLE = 0x0001, 0xFFFE


class S24Recording:
    """A 24-bit PCM recording left in its file mapping.  Looks like the int32 array scipy would have
    returned (``shape``, ``dtype``, ``len``; ``np.asarray`` expands it in full), and hands the kept
    frames of a decimation to a caller-provided buffer without expanding the rest."""

    dtype = np.dtype(np.int32)

    def __init__(self, raw: np.ndarray, channels: int, n_frames: int):
        self.raw, self.channels, self.n_frames = raw, int(channels), int(n_frames)
        self.shape = (self.n_frames,) if self.channels == 1 else (self.n_frames, self.channels)
        self.ndim = len(self.shape)

    def __len__(self) -> int:
        return self.n_frames

    def gather_into(self, out_ptr: int, stride: int, n_threads: int = 0) -> None:
        """out[j, c] = frame j * stride, channel c, as scipy's int32 (``audio_data[::stride]``)."""
        rc = load_host_library().bpm_host_gather_s24(C.c_void_p(self.raw.ctypes.data), self.channels, self.n_frames,
                                                     int(stride), C.c_void_p(out_ptr), int(n_threads))
        if rc != 0:
            raise RuntimeError(f"bpm_host_gather_s24 failed ({rc})")

    def decimated(self, stride: int) -> np.ndarray:
        m = (self.n_frames + stride - 1) // stride
        out = np.empty((m,) if self.channels == 1 else (m, self.channels), dtype=np.int32)
        if m:
            self.gather_into(out.ctypes.data, stride)
        return out

    def __array__(self, dtype=None, copy=None):
        a = self.decimated(1)
        return a if dtype is None else a.astype(dtype, copy=False)


def _chunks(buf: memoryview, start: int, end: int):
    pos = start
    while pos + 8 <= end:
        cid, size = bytes(buf[pos:pos + 4]), struct.unpack_from("<I", buf, pos + 4)[0]
        yield cid, pos + 8, size
        pos += 8 + size + (size & 1)                                  # chunks are word-aligned


def map_s24(file_path: str) -> Optional[Tuple[int, S24Recording]]:
    """(sample_rate, recording) for a little-endian RIFF/WAVE file holding 24-bit integer PCM whose data
    chunk lies completely inside the file, else None."""
    try:
        raw = np.memmap(file_path, dtype=np.uint8, mode="r")
    except (OSError, ValueError):
        return None
    buf = memoryview(raw)
    if raw.size < 44 or bytes(buf[0:4]) != b"RIFF" or bytes(buf[8:12]) != b"WAVE":
        return None
    fmt = None
    for cid, off, size in _chunks(buf, 12, raw.size):
        if cid == b"fmt ":
            if size < 16 or off + size > raw.size:
                return None
            tag, channels, rate, _, block_align, bits = struct.unpack_from("<HHIIHH", buf, off)
            if tag == WAVE_FORMAT_EXTENSIBLE and size >= 40:
                tag = struct.unpack_from("<H", buf, off + 24)[0]      # first field of the sub-format GUID
            fmt = (tag, channels, rate, block_align, bits)
        elif cid == b"data":
            if fmt is None:
                return None
            tag, channels, rate, block_align, bits = fmt
            if tag != WAVE_FORMAT_PCM or bits != 24 or channels < 1 or block_align != 3 * channels:
                return None
            if size == 0 or size % block_align or off + size > raw.size:
                return None
            return int(rate), S24Recording(raw[off:off + size], channels, size // block_align)
    return None

"""Device JPEG loader of the embedding loop (reference compute_feats.py:26-29,55: `Image.open(path)` +
`VF.to_tensor` in 4 DataLoader workers, then 77 MB of fp32 per batch over PCIe at compute_feats.py:72).

Here the FILES cross PCIe (~2 MB per 128-patch batch): the host only parses the marker segments
(`dsmil_jpeg_parse_batch`, libdsmil_host.so) and the device does the rest (`dsmil_jpeg_decode_batch`,
libdsmil_b200.so: Huffman decoding one warp per patch, IDCT, chroma upsampling, colour, /255) -- bit for bit what PIL's
libjpeg produces.  Files the device path does not take (progressive, CMYK, ...) are reported, never silently decoded
differently: the caller (embed.embed_bag) routes such a batch through PIL, the reference's own decoder.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _hostlib, _lib

STATUS = {0: "ok", -1: "corrupt JPEG data", -2: "not decodable on the device (progressive / CMYK / other size ...)"}


def header_bytes() -> int:
    hb = int(_hostlib.load().dsmil_jpeg_header_bytes())
    return hb


class ParsedBatch:
    """n files back to back in one pinned buffer + their parsed headers (pinned), ready for one H2D copy each."""
    __slots__ = ("n", "blob", "blob_bytes", "headers", "offsets", "bad", "H", "W", "statuses")

    def __init__(self, n, blob, blob_bytes, headers, offsets, bad, H, W, statuses):
        self.n, self.blob, self.blob_bytes, self.headers, self.offsets = n, blob, blob_bytes, headers, offsets
        self.bad, self.H, self.W, self.statuses = bad, H, W, statuses

    def file_bytes(self, i: int) -> bytes:
        """The bytes of file i (for a decoder outside the device path)."""
        return self.blob.numpy()[int(self.offsets[i]):int(self.offsets[i + 1])].tobytes()


class _Pinned:
    """A growable pinned byte buffer (re-pinning is expensive: grow geometrically, reuse across batches)."""

    def __init__(self):
        self.t: Optional[torch.Tensor] = None

    def get(self, nbytes: int, pin: bool) -> torch.Tensor:
        if self.t is None or self.t.numel() < nbytes:
            cap = max(nbytes, int(1.5 * (self.t.numel() if self.t is not None else 0)), 1 << 16)
            t = torch.empty(cap, dtype=torch.uint8)
            self.t = t.pin_memory() if pin else t
        return self.t


def _parse_blob(blob: torch.Tensor, total: int, offsets: np.ndarray, n: int, hdr_buf: Optional[_Pinned], pin: bool
                ) -> ParsedBatch:
    lib = _hostlib.load()
    hb = header_bytes()
    headers = (hdr_buf or _Pinned()).get(max(n, 1) * hb, pin)
    bad = int(lib.dsmil_jpeg_parse_batch(blob.data_ptr(), offsets.ctypes.data, n, headers.data_ptr())) if n else 0
    if bad < 0:
        raise RuntimeError("dsmil_jpeg_parse_batch: bad argument")
    hv = headers.numpy()[:n * hb].reshape(n, hb)
    # int32 fields at the head of the record: file_off (8 bytes), width, height, ..., status at byte 48
    i32 = hv[:, :56].copy().view(np.int32)
    widths, heights, statuses = i32[:, 2], i32[:, 3], i32[:, 12]
    H = int(heights[0]) if n and statuses[0] == 0 else 0
    W = int(widths[0]) if n and statuses[0] == 0 else 0
    if n and bad == 0 and (np.any(widths != W) or np.any(heights != H)):
        bad = int(np.sum((widths != W) | (heights != H)))
    return ParsedBatch(n, blob, total, headers, offsets, bad, H, W, statuses.copy())


def parse_batch(files: Sequence[bytes], blob_buf: Optional[_Pinned] = None, hdr_buf: Optional[_Pinned] = None,
                pin: bool = False) -> ParsedBatch:
    """Concatenates the files (bytes objects) into one (optionally pinned) buffer and parses their headers."""
    n = len(files)
    sizes = np.fromiter((len(f) for f in files), dtype=np.int64, count=n)
    offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(sizes, out=offsets[1:])
    total = int(offsets[-1])
    blob = (blob_buf or _Pinned()).get(total + 16, pin)
    view = blob.numpy()
    for i, f in enumerate(files):
        view[offsets[i]:offsets[i + 1]] = np.frombuffer(f, dtype=np.uint8)
    return _parse_blob(blob, total, offsets, n, hdr_buf, pin)


def parse_paths(paths: Sequence[str], blob_buf: Optional[_Pinned] = None, hdr_buf: Optional[_Pinned] = None,
                pin: bool = False, threads: int = 4) -> ParsedBatch:
    """Reads the files natively (libdsmil_host.so, `threads` reader threads, the GIL released) straight into the
    (optionally pinned) blob and parses their headers: no per-file Python objects."""
    import ctypes as C
    import os
    lib = _hostlib.load()
    n = len(paths)
    arr = (C.c_char_p * max(n, 1))(*[os.fsencode(p) for p in paths])
    offsets = np.zeros(n + 1, dtype=np.int64)
    rc = int(lib.dsmil_files_offsets(arr, n, offsets.ctypes.data))
    if rc < 0:
        raise FileNotFoundError(f"cannot stat patch file {paths[-rc - 1]}")
    total = int(offsets[-1])
    blob = (blob_buf or _Pinned()).get(total + 16, pin)
    rc = int(lib.dsmil_files_read(arr, n, offsets.ctypes.data, blob.data_ptr(), int(threads)))
    if rc < 0:
        raise OSError(f"cannot read patch file {paths[-rc - 1]}")
    return _parse_blob(blob, total, offsets, n, hdr_buf, pin)


class JpegBatchDecoder:
    """Reusable device-side state of the loader: device copies of blob / headers, workspace, status."""

    def __init__(self, device: torch.device):
        if torch.device(device).type != "cuda":
            raise RuntimeError("the JPEG loader decodes on a CUDA device (no CPU path)")
        self.device = torch.device(device)
        self.lib = _lib.load()
        if int(self.lib.dsmil_jpeg_header_bytes_dev()) != header_bytes():
            raise RuntimeError("libdsmil_b200.so and libdsmil_host.so disagree on the JPEG header record: rebuild both")
        self._blob = self._hdr = self._ws = self._status = None

    def _dev(self, cur: Optional[torch.Tensor], nbytes: int) -> torch.Tensor:
        if cur is None or cur.numel() < nbytes:
            cur = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=self.device)
        return cur

    def decode(self, pb: ParsedBatch, out_f32: Optional[torch.Tensor] = None, out_u8: Optional[torch.Tensor] = None,
               stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
        """Copies blob + headers to the device and launches the three kernels on `stream` (default: current).
        Returns the device int32 status vector [n] (read it after synchronising; 0 = decoded)."""
        if pb.n == 0:
            return torch.empty(0, dtype=torch.int32, device=self.device)
        if pb.H < 1 or pb.W < 1:
            raise ValueError("the first file of the batch is not decodable on the device; route the batch elsewhere")
        hb = header_bytes()
        n, H, W = pb.n, pb.H, pb.W
        channels_last = 0
        if out_f32 is not None:
            if tuple(out_f32.shape) != (n, 3, H, W) or out_f32.dtype != torch.float32 or out_f32.device != self.device:
                raise ValueError(f"out_f32 must be a float32 tensor of shape {(n, 3, H, W)} on {self.device}")
            if out_f32.is_contiguous():
                channels_last = 0
            elif out_f32.is_contiguous(memory_format=torch.channels_last):
                channels_last = 1                          # same values, [n, H, W, 3] in memory
            else:
                raise ValueError("out_f32 must be contiguous (NCHW) or torch.channels_last")
        if out_u8 is not None and (tuple(out_u8.shape) != (n, H, W, 3) or out_u8.dtype != torch.uint8 or
                                   not out_u8.is_contiguous() or out_u8.device != self.device):
            raise ValueError(f"out_u8 must be a contiguous uint8 tensor of shape {(n, H, W, 3)} on {self.device}")
        if out_f32 is None and out_u8 is None:
            raise ValueError("no output requested")
        st = stream or torch.cuda.current_stream(self.device)
        need = int(self.lib.dsmil_jpeg_workspace_bytes(n, H, W, pb.blob_bytes))
        with torch.cuda.device(self.device), torch.cuda.stream(st):
            self._blob = self._dev(self._blob, pb.blob_bytes + 16)
            self._hdr = self._dev(self._hdr, n * hb)
            self._ws = self._dev(self._ws, need)
            if self._status is None or self._status.numel() < n:
                self._status = torch.empty(max(n, 128), dtype=torch.int32, device=self.device)
            self._blob[:pb.blob_bytes].copy_(pb.blob[:pb.blob_bytes], non_blocking=True)
            self._hdr[:n * hb].copy_(pb.headers[:n * hb], non_blocking=True)
            ws_ptr = (self._ws.data_ptr() + 255) & ~255
            _lib.check(self.lib.dsmil_jpeg_decode_batch(
                self._blob.data_ptr(), pb.blob_bytes, self._hdr.data_ptr(), n, H, W,
                out_u8.data_ptr() if out_u8 is not None else None,
                out_f32.data_ptr() if out_f32 is not None else None, channels_last,
                self._status.data_ptr(), ws_ptr, self._ws.numel() - (ws_ptr - self._ws.data_ptr()), st.cuda_stream),
                "dsmil_jpeg_decode_batch")
        return self._status[:n]


def decode_files(files: Sequence[bytes], device, want_u8: bool = True, want_f32: bool = True
                 ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor], List[int]]:
    """One-shot convenience (tests, small tools): bytes of n same-size JPEG files -> (uint8 [n,H,W,3] or None,
    float32 [n,3,H,W] or None, per-file status list).  Raises if the batch is not device-decodable."""
    pb = parse_batch(files, pin=False)
    if pb.bad:
        raise ValueError(f"{pb.bad} of {pb.n} files are not decodable on the device (statuses {pb.statuses.tolist()})")
    dev = torch.device(device)
    dec = JpegBatchDecoder(dev)
    u8 = torch.empty(pb.n, pb.H, pb.W, 3, dtype=torch.uint8, device=dev) if want_u8 else None
    f32 = torch.empty(pb.n, 3, pb.H, pb.W, dtype=torch.float32, device=dev) if want_f32 else None
    status = dec.decode(pb, out_f32=f32, out_u8=u8)
    torch.cuda.synchronize(dev)
    return u8, f32, status.cpu().tolist()

"""The data formats either side of the hot path, host side.

    PointPairFile      the packed float32 `.bin` the reference trains from: [pairs, 2*point_num, 3], source = first
                       point_num rows, target = the rest (train_view.py:89-91, 369-370; written by
                       convert_seed_to_dataset.py:105-136)
    PairBatchLoader    the reference's batch walk (train_view.py:355-366: contiguous slices with wrap-around), but
                       through pinned staging buffers and a copy stream, double-buffered, so the H2D copy of batch k+1
                       overlaps the step on batch k.  The reference does a synchronous pageable `.cuda()` per step.
                       Batches are handed out as strided VIEWS [B,N,3] of one [B,2N,3] device buffer (batch stride
                       6N) — exactly what the reference passes on, and what the kernels accept without a copy.
    read_obj_vertices  `v x y z [r g b]` lines -> float32 [V,3] (+ colours); everything else is ignored, like
                       igl.read_triangle_mesh's vertex output (inference.py:52-53; sample OBJs carry 3 colour columns)
    write_obj_points   the text the reference writes (utils/drawer.py:81-113, inference.py:27-32): one
                       'v x y z[ r g b]' line per point, numbers formatted by Python's str() — byte-identical files
"""
from __future__ import annotations

import os
from typing import Iterator, Optional, Tuple

import numpy as np
import torch


class PointPairFile:
    """Read-only view of a `.bin` pair file (np.memmap; nothing is read until it is sliced)."""

    def __init__(self, path: str, point_num: int):
        size = os.path.getsize(path)
        rec = point_num * 2 * 3 * 4
        if point_num <= 0 or size == 0 or size % rec:
            raise ValueError(f"{path}: {size} bytes is not a whole number of [2*{point_num},3] float32 pairs")
        self.path, self.point_num = path, point_num
        self.pairs = np.memmap(path, dtype=np.float32, mode="r").reshape(-1, point_num * 2, 3)

    def __len__(self) -> int:
        return self.pairs.shape[0]

    def batch_indices(self, start_pos: int, batch_size: int) -> np.ndarray:
        """Pair indices of the batch starting at start_pos, wrapping around the end (train_view.py:357-366)."""
        return (start_pos + np.arange(batch_size)) % len(self)

    def read_batch(self, start_pos: int, batch_size: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        n = len(self)
        if batch_size > n:
            raise ValueError(f"batch_size {batch_size} exceeds the {n} pairs of {self.path}")
        start_pos %= n
        if out is None:
            out = np.empty((batch_size, self.point_num * 2, 3), np.float32)
        bottom = min(batch_size, n - start_pos)
        out[:bottom] = self.pairs[start_pos:start_pos + bottom]
        if bottom < batch_size:
            out[bottom:] = self.pairs[:batch_size - bottom]
        return out

    @staticmethod
    def write(path: str, source: np.ndarray, target: np.ndarray) -> None:
        """[P,N,3] + [P,N,3] -> the packed layout (convert_seed_to_dataset.py:105-136)."""
        src = np.asarray(source, np.float32)
        tgt = np.asarray(target, np.float32)
        if src.shape != tgt.shape or src.ndim != 3 or src.shape[-1] != 3:
            raise ValueError("source and target must both be [pairs, point_num, 3]")
        np.concatenate([src, tgt], 1).astype(np.float32).tofile(path)


class PairBatchLoader:
    """Iterate (source [B,N,3], target [B,N,3], start_pos) device views over a PointPairFile.

    Two pinned host buffers and two device buffers; batch k+1 is read from the file into pinned memory and copied
    on a side stream while the caller works on batch k.  The views returned for batch k stay valid until the NEXT
    call of __next__ (by then the step on batch k has been enqueued; that call marks the point on the consumer's
    stream after which batch k's device buffer may be refilled); the consumer's stream is made to wait for the copy."""

    def __init__(self, pairs: PointPairFile, batch_size: int, device: torch.device, start_pos: int = 0,
                 steps: Optional[int] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("PairBatchLoader needs a CUDA device (rma_net_b200 has no CPU path)")
        self.pairs, self.batch_size, self.device = pairs, int(batch_size), torch.device(device)
        self.pos, self.steps, self.done = int(start_pos), steps, 0
        shape = (self.batch_size, pairs.point_num * 2, 3)
        self._pin = [torch.empty(shape, dtype=torch.float32).pin_memory() for _ in range(2)]
        self._dev = [torch.empty(shape, dtype=torch.float32, device=self.device) for _ in range(2)]
        self._ready = [torch.cuda.Event() for _ in range(2)]
        self._consumed = [torch.cuda.Event() for _ in range(2)]
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self._slot = 0
        self._last: Optional[int] = None
        self._copied = [False, False]
        self._inflight: Optional[Tuple[int, int]] = None
        self._prefetch()

    def _prefetch(self) -> None:
        if self.steps is not None and self.done >= self.steps:
            self._inflight = None
            return
        s = self._slot
        if self._copied[s]:
            self._ready[s].synchronize()                     # the slot's previous H2D copy has left the pinned buffer
        self.pairs.read_batch(self.pos, self.batch_size, out=self._pin[s].numpy())
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._consumed[s])  # the kernels that read the device buffer have finished
            self._dev[s].copy_(self._pin[s], non_blocking=True)
            self._ready[s].record(self._copy_stream)
        self._copied[s] = True
        self._inflight = (s, self.pos % len(self.pairs))
        self.pos = (self.pos + self.batch_size) % len(self.pairs)
        self.done += 1
        self._slot ^= 1

    def __iter__(self) -> Iterator:
        return self

    def __next__(self):
        if self._inflight is None:
            raise StopIteration
        s, start = self._inflight
        cur = torch.cuda.current_stream(self.device)
        # The step on the batch handed out last has been enqueued by now: everything on the consumer's stream up to
        # here must finish before that batch's device buffer (the slot the prefetch below refills) is overwritten.
        if self._last is not None:
            self._consumed[self._last].record(cur)
        cur.wait_event(self._ready[s])
        batch = self._dev[s]
        self._prefetch()                                     # next batch: file read + H2D overlap the caller's step
        n = self.pairs.point_num
        self._last = s
        return batch[:, :n, :], batch[:, n:, :], start

    def release(self) -> None:
        """Optional earlier marker: call after the step on the batch returned last has been ENQUEUED to let the
        refill of its buffer start before the next __next__ (which records the same marker itself)."""
        if self._last is not None:
            self._consumed[self._last].record(torch.cuda.current_stream(self.device))


def read_obj_vertices(path: str, with_colors: bool = False):
    """-> float32 [V,3] (and float32 [V,3] colours or None)."""
    verts, cols = [], []
    with open(path) as f:
        for line in f:
            if not line.startswith("v ") and not line.startswith("v\t"):
                continue
            t = line.split()
            verts.append((float(t[1]), float(t[2]), float(t[3])))
            if with_colors:
                cols.append((float(t[4]), float(t[5]), float(t[6])) if len(t) >= 7 else (np.nan,) * 3)
    v = np.asarray(verts, np.float32).reshape(-1, 3)
    if not with_colors:
        return v
    c = np.asarray(cols, np.float32).reshape(-1, 3)
    return v, (None if c.size and np.isnan(c).all() else c)


def write_obj_points(points, path: str, rgb=None) -> None:
    """points [V,3] (numpy or tensor); rgb: None, one (r,g,b) for all points (inference.py:27-32,
    drawer.render_points_with_rgb) or [V,3] per-point colours (drawer.render_points_with_seg_info)."""
    if isinstance(points, torch.Tensor):
        points = points.detach().cpu().numpy()
    p = np.asarray(points).reshape(-1, 3)
    per_point = rgb is not None and np.ndim(rgb) == 2
    with open(path, "w") as f:
        for i in range(p.shape[0]):
            line = "v " + str(p[i][0]) + " " + str(p[i][1]) + " " + str(p[i][2])
            if rgb is not None:
                c = rgb[i] if per_point else rgb
                line += " " + str(c[0]) + " " + str(c[1]) + " " + str(c[2])
            f.write(line + "\n")

"""Data formats either side of the hot path: the packed `.bin` pair file with the
reference's wrap-around batch walk, OBJ vertex IO byte-compatible with the reference's writers."""
import os

import numpy as np
import pytest
import torch

from rma_net_b200 import data

REF = "/root/reference"


def _reference_batch(pairs_np, start_pos, batch_size):
    """train_view.py:355-366 restated on numpy (slices with wrap-around)."""
    n = pairs_np.shape[0]
    end_pos = start_pos + batch_size
    if end_pos <= n:
        return pairs_np[start_pos:end_pos]
    return np.concatenate([pairs_np[start_pos:], pairs_np[:end_pos - n]], 0)


def test_pair_file_round_trip_and_wraparound(tmp_path):
    rng = np.random.default_rng(0)
    src = rng.normal(size=(7, 50, 3)).astype(np.float32)
    tgt = rng.normal(size=(7, 50, 3)).astype(np.float32)
    path = str(tmp_path / "pairs.bin")
    data.PointPairFile.write(path, src, tgt)
    ref = np.fromfile(path, dtype=np.float32).reshape(-1, 50 * 2, 3)              # train_view.py:89
    assert np.array_equal(ref[:, :50], src) and np.array_equal(ref[:, 50:], tgt)
    f = data.PointPairFile(path, 50)
    assert len(f) == 7
    for start in (0, 3, 5, 6):
        got = f.read_batch(start, 3)
        assert np.array_equal(got, _reference_batch(ref, start, 3))
        assert np.array_equal(f.batch_indices(start, 3), (start + np.arange(3)) % 7)
    with pytest.raises(ValueError):
        data.PointPairFile(path, 49)
    with pytest.raises(ValueError):
        f.read_batch(0, 8)


def test_obj_round_trip_and_reference_text(tmp_path, golden_dir):
    g = np.load(os.path.join(golden_dir, "chamfer_samples3.npz"))
    pts = g["x"][0]
    p1, p2, p3 = (str(tmp_path / n) for n in ("a.obj", "b.obj", "c.obj"))
    data.write_obj_points(pts, p1)
    data.write_obj_points(torch.from_numpy(pts), p2, rgb=(0.9, 0.1, 0.1))
    seg = (np.arange(pts.shape[0]) % 3)[:, None] * np.array([[0.25, 0.5, 0.75]])
    data.write_obj_points(pts, p3, rgb=seg)
    assert np.array_equal(data.read_obj_vertices(p1), pts)                         # str(float32) round-trips exactly
    v, c = data.read_obj_vertices(p2, with_colors=True)
    assert np.array_equal(v, pts) and np.allclose(c, [0.9, 0.1, 0.1])
    v, c = data.read_obj_vertices(p3, with_colors=True)
    assert np.allclose(c, seg)
    assert data.read_obj_vertices(p1, with_colors=True)[1] is None
    # the reference's own writers, executed verbatim when the reference tree is present (build container only)
    if os.path.isdir(REF):
        ns = {}
        with open(os.path.join(REF, "utils", "drawer.py")) as fh:
            lines = fh.readlines()
        exec("".join(lines[80:86]), ns)                                            # render_points, drawer.py:81-86
        ns["render_points"](pts, str(tmp_path / "ref.obj"))
        assert open(str(tmp_path / "ref.obj")).read() == open(p1).read()
        # and its sample OBJ (v x y z r g b) parses to the vertex set the golden file was generated from
        sample = os.path.join(REF, "samples", "3", "source.obj")
        if os.path.exists(sample):
            assert np.array_equal(data.read_obj_vertices(sample), pts)


@pytest.mark.gpu
def test_pair_batch_loader_double_buffered(cuda, tmp_path):
    from rma_net_b200.model.loss import chamfer_dist
    rng = np.random.default_rng(1)
    P, N, B = 11, 300, 4
    src = rng.uniform(-1, 1, (P, N, 3)).astype(np.float32)
    tgt = rng.uniform(-1, 1, (P, N, 3)).astype(np.float32)
    path = str(tmp_path / "pairs.bin")
    data.PointPairFile.write(path, src, tgt)
    f = data.PointPairFile(path, N)
    loader = data.PairBatchLoader(f, B, cuda, start_pos=9, steps=7)
    seen = 0
    for s, t, start in loader:
        idx = f.batch_indices(start, B)
        assert not s.is_contiguous() and s.stride(0) == 6 * N                      # views, as train_view.py:369-370
        d1, d2 = chamfer_dist(s, t)                                                # strided views go straight in
        loader.release()
        assert torch.equal(s.cpu(), torch.from_numpy(src[idx])) and torch.equal(t.cpu(), torch.from_numpy(tgt[idx]))
        assert d1.shape == (B, N)
        seen += 1
    assert seen == 7
    assert start == (9 + 6 * B) % P

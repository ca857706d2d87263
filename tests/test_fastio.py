"""Frame I/O fast path: uncompressed TIFFs -- what the reference's write_img produces
(utils.py:23-30) -- are mapped and converted straight into the staging buffer; the result must be the array
cv2.imread(IMREAD_UNCHANGED) returns (reference utils.read_img), and everything else must fall back to it."""
import numpy as np
import pytest

from shinestacker_b200 import fastio, PyramidStack, DepthMapStack, synthetic
from shinestacker_b200.pyramid import read_img, stage_frame

cv2 = pytest.importorskip("cv2")
UNCOMPRESSED = [int(cv2.IMWRITE_TIFF_COMPRESSION), 1]


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
@pytest.mark.parametrize("shape", [(37, 53), (64, 64), (300, 407), (1, 5)])
def test_plain_tiff_equals_imread(tmp_path, dtype, shape):
    img = np.random.default_rng(shape[0]).integers(0, np.iinfo(dtype).max + 1, shape + (3,)).astype(dtype)
    p = str(tmp_path / "a.tif")
    cv2.imwrite(p, img, UNCOMPRESSED)
    t = fastio.open_plain_tiff(p)
    assert t is not None and t.shape == img.shape and t.dtype == img.dtype and t.ndim == 3 and t.nbytes == img.nbytes
    assert np.array_equal(t.to_array(), cv2.imread(p, cv2.IMREAD_UNCHANGED))
    assert np.array_equal(t.to_array(), img)
    view = np.zeros(img.nbytes + 64, np.uint8)          # a staging buffer larger than the frame
    stage_frame(fastio.read_frame(p, read_img), view)
    assert np.array_equal(view[:img.nbytes].view(dtype).reshape(img.shape), img)


def test_everything_else_falls_back_to_imread(tmp_path):
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, (40, 50, 3)).astype(np.uint8)
    cases = {}
    cv2.imwrite(str(tmp_path / "lzw.tif"), img)                                   # cv2's default: LZW
    cases["lzw.tif"] = cv2.IMREAD_UNCHANGED
    cv2.imwrite(str(tmp_path / "gray.tif"), img[..., 0], UNCOMPRESSED)
    cases["gray.tif"] = cv2.IMREAD_UNCHANGED
    cv2.imwrite(str(tmp_path / "rgba.tif"), np.dstack([img, img[..., :1]]), UNCOMPRESSED)
    cases["rgba.tif"] = cv2.IMREAD_UNCHANGED
    cv2.imwrite(str(tmp_path / "a.png"), img)
    cases["a.png"] = cv2.IMREAD_UNCHANGED
    cv2.imwrite(str(tmp_path / "a.jpg"), img)
    cases["a.jpg"] = cv2.IMREAD_COLOR
    for name, flag in cases.items():
        p = str(tmp_path / name)
        assert fastio.open_plain_tiff(p) is None, name
        got = fastio.read_frame(p, read_img)
        assert isinstance(got, np.ndarray) and np.array_equal(got, cv2.imread(p, flag)), name
    # damaged files: truncated pixel data, garbage header
    good = str(tmp_path / "good.tif")
    cv2.imwrite(good, img, UNCOMPRESSED)
    data = open(good, "rb").read()
    open(str(tmp_path / "trunc.tif"), "wb").write(data[:len(data) // 2])
    open(str(tmp_path / "junk.tif"), "wb").write(b"II*\\x00" + bytes(range(200)))
    assert fastio.open_plain_tiff(str(tmp_path / "trunc.tif")) is None
    assert fastio.open_plain_tiff(str(tmp_path / "junk.tif")) is None
    with pytest.raises(RuntimeError):                                             # reference utils.py:12-13
        fastio.read_frame(str(tmp_path / "missing.tif"), read_img)


@pytest.mark.parametrize("algo_cls", [PyramidStack, DepthMapStack])
@pytest.mark.parametrize("threads", [0, 3])
def test_stackers_stage_mapped_frames_in_order(tmp_path, monkeypatch, algo_cls, threads):
    frames = synthetic.make_noise_stack(5, 40, 56, np.uint16, 4)
    paths = []
    for i, f in enumerate(frames):
        p = str(tmp_path / f"f{i:03d}.tif")
        cv2.imwrite(p, f, UNCOMPRESSED if i != 2 else [])       # one compressed file in the middle: mixed paths
        paths.append(p)
    staged, kinds, swapped = [], [], []

    class Lib:
        def __getattr__(self, name):
            return lambda *a: 0

    monkeypatch.setattr(algo_cls, "_ensure_handle", lambda self, *a, **k: setattr(self, "_h", None))
    # worker threads stage every frame straight into a pinned slot; the upload reads the slot
    by_path = {}
    real_read = fastio.read_frame

    def spy(path, fallback):
        img = real_read(path, fallback)
        by_path[path] = type(img).__name__
        return img

    def push_slot(self, i, slot):
        img = slot[1][:frames[0].nbytes].view(np.uint16).reshape(frames[0].shape).copy()
        # PyramidStack stages mapped TIFFs as their raw R, G, B payload (ss_push_frame_rgb swaps on the device);
        # DepthMapStack converts to BGR while staging
        rgb = i in getattr(self, "_rgb_frames", ())
        staged.append(img[..., ::-1] if rgb else img)
        swapped.append(rgb)

    monkeypatch.setattr(fastio, "read_frame", spy)
    monkeypatch.setattr(algo_cls, "_staging",
                        lambda self, nbytes, count: [(None, np.zeros(nbytes, np.uint8)) for _ in range(2)])
    monkeypatch.setattr(algo_cls, "_push_slot", push_slot)
    monkeypatch.setattr(algo_cls, "_sync_copies", lambda self: None)
    if algo_cls is PyramidStack:
        monkeypatch.setattr(algo_cls, "_finish", lambda self, n, shape: np.zeros(tuple(shape) + (3,), self.dtype))
    else:
        monkeypatch.setattr(algo_cls, "_blend_and_finish", lambda self, n, shape, dtype, names=None: np.zeros(shape + (3,), dtype))
    algo = algo_cls(decode_threads=threads)
    algo._lib = Lib()
    out = algo.focus_stack(paths)
    assert out.shape == frames[0].shape
    kinds = [by_path[p] for p in paths]
    assert swapped == ([True, True, False, True, True] if algo_cls is PyramidStack else [False] * 5)
    assert kinds == ["PlainTiff", "PlainTiff", "ndarray", "PlainTiff", "PlainTiff"]
    for got, f in zip(staged, frames):
        assert np.array_equal(got, f)


@pytest.mark.parametrize("ext,dtype,params", [
    ("tif", np.uint8, UNCOMPRESSED), ("tif", np.uint16, UNCOMPRESSED), ("tif", np.uint8, []), ("tif", np.uint16, []),
    ("png", np.uint8, []), ("png", np.uint16, []), ("jpg", np.uint8, []), ("jpeg", np.uint8, [int(cv2.IMWRITE_JPEG_PROGRESSIVE), 1]),
])
def test_header_probe_agrees_with_read_img(tmp_path, ext, dtype, params):
    """The validation pass reads headers only; what it reports must be what read_img (reference utils.py:11-20) delivers."""
    img = np.random.default_rng(5).integers(0, np.iinfo(dtype).max + 1, (37, 61, 3)).astype(dtype)
    p = str(tmp_path / f"a.{ext}")
    cv2.imwrite(p, img, params)
    ref = read_img(p)
    assert fastio.probe_header(p) == (ref.shape[0], ref.shape[1], ref.shape[2], ref.dtype)


def test_header_probe_declines_what_it_cannot_settle(tmp_path):
    img = np.random.default_rng(6).integers(0, 256, (20, 30, 3)).astype(np.uint8)
    cv2.imwrite(str(tmp_path / "gray.tif"), img[..., 0])
    cv2.imwrite(str(tmp_path / "rgba.png"), np.dstack([img, img[..., :1]]))
    cv2.imwrite(str(tmp_path / "gray.png"), img[..., 0])
    (tmp_path / "junk.tif").write_bytes(b"not a tiff at all")
    (tmp_path / "empty.jpg").write_bytes(b"")
    for name in ("gray.tif", "rgba.png", "gray.png", "junk.tif", "empty.jpg", "missing.tif"):
        assert fastio.probe_header(str(tmp_path / name)) is None

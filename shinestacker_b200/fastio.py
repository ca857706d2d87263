"""Frame I/O at the boundary: a zero-decode path for the files the reference itself
writes between actions.

``write_img`` stores TIFFs uncompressed (reference src/shinestacker/algorithms/utils.py:23-30:
``IMWRITE_TIFF_COMPRESSION = 1``), so the aligned / balanced frames that ``FocusStack`` reads back
(``read_img``, utils.py:11-20 -> ``cv2.imread(IMREAD_UNCHANGED)``, ~100 ms for a 12.6 MP frame) are plain RGB
sample arrays behind a small header.  ``open_plain_tiff`` maps such a file and ``PlainTiff.into_bgr`` converts
straight from the page cache into the pinned staging buffer the upload reads (one pass, RGB -> BGR with cv2's
SIMD ``cvtColor``), producing exactly the array ``cv2.imread`` returns.  Anything else -- compression, tiles,
predictors, palettes, alpha, big-endian 16-bit, BigTIFF, non-contiguous strips, other formats -- returns ``None``
and the caller falls back to ``read_img``.
"""
import mmap
import os
import struct

import numpy as np

ENABLED = os.environ.get("SHINESTACK_B200_FASTIO", "1") != "0"     # A/B switch: "0" always decodes with cv2.imread

_TYPE_SIZE = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 16: 8}
_TYPE_FMT = {1: "B", 3: "H", 4: "I", 16: "Q"}


class PlainTiff:
    """An uncompressed, chunky 8/16-bit RGB TIFF mapped into memory; quacks enough like the ndarray that
    ``cv2.imread`` would return (shape / dtype / ndim / nbytes) for the validation code of the stackers."""

    def __init__(self, mm, offset, h, w, dtype, path=None):
        self._mm = mm
        self._path, self._offset = path, offset
        self.shape = (h, w, 3)
        self.dtype = np.dtype(dtype)
        self.ndim = 3
        self.nbytes = h * w * 3 * self.dtype.itemsize
        self._rgb = np.frombuffer(mm, dtype=self.dtype, count=h * w * 3, offset=offset).reshape(h, w, 3)

    def into_bgr(self, dst):
        """Writes the BGR image (what cv2.imread delivers) into ``dst`` (H x W x 3 array of the same dtype)."""
        import cv2  # pylint: disable=import-outside-toplevel
        cv2.cvtColor(self._rgb, cv2.COLOR_RGB2BGR, dst=dst)
        return dst

    def read_rgb_into(self, dst):
        """The file's R, G, B samples, unconverted, into the flat uint8 buffer ``dst`` (one read() from the page cache
        instead of touching the mapping page by page); the caller swaps the channels (on the device)."""
        with open(self._path, "rb", buffering=0) as f:
            f.seek(self._offset)
            view = memoryview(dst)[:self.nbytes]
            got = 0
            while got < self.nbytes:
                k = f.readinto(view[got:])
                if not k:
                    raise OSError(f"short read from {self._path}")
                got += k

    def to_array(self):
        return self.into_bgr(np.empty(self.shape, self.dtype))

    copy = to_array

    def close(self):
        self._rgb = None
        try:
            self._mm.close()
        except (BufferError, ValueError):
            pass


def open_plain_tiff(path):
    """``PlainTiff`` for a file of the simple layout described above, else ``None``."""
    try:
        size = os.path.getsize(path)
        if size < 16:
            return None
        with open(path, "rb") as f:
            head = f.read(8)
            if head[:4] != b"II*\x00":       # little-endian classic TIFF only
                return None
            (ifd,) = struct.unpack("<I", head[4:8])
            if ifd + 2 > size:
                return None
            f.seek(ifd)
            (n,) = struct.unpack("<H", f.read(2))
            raw = f.read(12 * n + 4)
            if len(raw) < 12 * n + 4:
                return None

            def values(typ, cnt, val4):
                if typ not in _TYPE_FMT:
                    return None
                nbytes = _TYPE_SIZE[typ] * cnt
                if nbytes <= 4:
                    data = val4[:nbytes]
                else:
                    (off,) = struct.unpack("<I", val4)
                    if off + nbytes > size:
                        return None
                    f.seek(off)
                    data = f.read(nbytes)
                return np.frombuffer(data, dtype="<" + {"B": "u1", "H": "u2", "I": "u4", "Q": "u8"}[_TYPE_FMT[typ]],
                                     count=cnt).astype(np.int64)

            tags = {}
            for i in range(n):
                tag, typ, cnt = struct.unpack("<HHI", raw[12 * i:12 * i + 8])
                tags[tag] = (typ, cnt, raw[12 * i + 8:12 * i + 12])
            if struct.unpack("<I", raw[12 * n:12 * n + 4])[0] != 0:
                return None                   # more than one image in the file

            def tag(t, default=None):
                if t not in tags:
                    return default
                return values(*tags[t])

            w, h = tag(256), tag(257)
            bits, comp, photo = tag(258), tag(259, np.array([1])), tag(262)
            spp, planar = tag(277, np.array([1])), tag(284, np.array([1]))
            fmt, pred, orient = tag(339, np.array([1])), tag(317, np.array([1])), tag(274, np.array([1]))
            offs, counts, rps = tag(273), tag(279), tag(278)
            if any(v is None for v in (w, h, bits, comp, photo, spp, planar, fmt, pred, orient, offs, counts)):
                return None
            if 322 in tags or 323 in tags or 338 in tags:       # tiles / extra samples
                return None
            w, h = int(w[0]), int(h[0])
            if int(comp[0]) != 1 or int(photo[0]) != 2 or int(spp[0]) != 3 or int(planar[0]) != 1:
                return None
            if int(pred[0]) != 1 or int(orient[0]) != 1 or np.any(fmt != 1) or len(bits) != 3:
                return None
            if not (np.all(bits == 8) or np.all(bits == 16)):
                return None
            bps = int(bits[0]) // 8
            total = h * w * 3 * bps
            rows = int(rps[0]) if rps is not None else h
            if rows < 1 or len(offs) != len(counts) or len(offs) != (h + rows - 1) // rows:
                return None
            if int(counts.sum()) != total or np.any(offs[1:] != offs[:-1] + counts[:-1]):
                return None                   # strips are not one contiguous block
            start = int(offs[0])
            if start + total > size or (bps == 2 and start % 2):
                return None
            mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
        return PlainTiff(mm, start, h, w, np.uint8 if bps == 1 else np.uint16, path)
    except (OSError, ValueError, struct.error):
        return None


def probe_header(path):
    """(height, width, channels, dtype) that ``read_img`` (reference utils.py:11-20) would deliver for ``path``, read
    from the file header only, or ``None`` when the header does not settle it (the caller then decodes the file).
    Covers what the reference's own pipeline produces and consumes: baseline / progressive JPEG without an EXIF
    block (cv2 rotates by the EXIF orientation), 8/16-bit RGB PNG, 8/16-bit RGB classic TIFF."""
    try:
        ext = path.split(".")[-1]
        with open(path, "rb") as f:
            head = f.read(32)
            if ext in ("tif", "tiff"):
                if head[:4] == b"II*\x00":
                    e = "<"
                elif head[:4] == b"MM\x00*":
                    e = ">"
                else:
                    return None
                (ifd,) = struct.unpack(e + "I", head[4:8])
                f.seek(ifd)
                (n,) = struct.unpack(e + "H", f.read(2))
                raw = f.read(12 * n)
                if len(raw) < 12 * n:
                    return None
                vals = {}
                for i in range(n):
                    tag, typ, cnt = struct.unpack(e + "HHI", raw[12 * i:12 * i + 8])
                    if tag not in (256, 257, 258, 262, 277, 339) or typ not in (3, 4):
                        continue
                    fmt, size = ("H", 2) if typ == 3 else ("I", 4)
                    if cnt * size <= 4:
                        data = raw[12 * i + 8:12 * i + 8 + cnt * size]
                    else:
                        (off,) = struct.unpack(e + "I", raw[12 * i + 8:12 * i + 12])
                        f.seek(off)
                        data = f.read(cnt * size)
                    vals[tag] = struct.unpack(e + fmt * cnt, data)
                if any(t not in vals for t in (256, 257, 258, 262)):
                    return None
                spp = vals.get(277, (1,))[0]
                bits = vals[258]
                if vals[262][0] != 2 or spp != 3 or len(bits) != 3 or any(v != 1 for v in vals.get(339, (1,))):
                    return None
                if all(b == 8 for b in bits):
                    dt = np.uint8
                elif all(b == 16 for b in bits):
                    dt = np.uint16
                else:
                    return None
                return int(vals[257][0]), int(vals[256][0]), 3, np.dtype(dt)
            if ext == "png":
                if head[:8] != b"\x89PNG\r\n\x1a\n" or head[12:16] != b"IHDR":
                    return None
                w, h, depth, ctype = struct.unpack(">IIBB", head[16:26])
                if ctype != 2 or depth not in (8, 16):
                    return None
                return int(h), int(w), 3, np.dtype(np.uint8 if depth == 8 else np.uint16)
            if ext in ("jpg", "jpeg"):
                if head[:2] != b"\xff\xd8":
                    return None
                f.seek(2)
                while True:
                    b = f.read(4)
                    if len(b) < 4 or b[0] != 0xFF:
                        return None
                    marker, (length,) = b[1], struct.unpack(">H", b[2:4])
                    if marker == 0xE1:
                        return None          # APP1 (EXIF): cv2.imread may rotate the image
                    if marker in (0xC0, 0xC1, 0xC2):
                        d = f.read(5)
                        h, w = struct.unpack(">HH", d[1:5])
                        return int(h), int(w), 3, np.dtype(np.uint8)     # default imread flags: 8-bit BGR
                    if marker == 0xDA or length < 2:
                        return None
                    f.seek(length - 2, 1)
    except (OSError, ValueError, struct.error, IndexError):
        return None
    return None


def read_frame(path, read_img):
    """What the stackers' decode-ahead workers call: the mapped file when the fast path applies (``.tif`` /
    ``.tiff`` only), otherwise ``read_img(path)`` (cv2.imread, as the reference)."""
    if not os.path.isfile(path):
        return read_img(path)               # raises like the reference (utils.py:12-13)
    if ENABLED and path.split(".")[-1] in ("tif", "tiff"):
        t = open_plain_tiff(path)
        if t is not None:
            return t
    return read_img(path)

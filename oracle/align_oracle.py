"""CPU oracle for the warp stage of AlignFrames -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module;
nothing under ``shinestacker_b200/`` does.

numpy restatement (integer arithmetic, no cv2 calls) of what ``align_images`` does with the estimated
transform (reference src/shinestacker/algorithms/align.py:216-251):

* ``cv2.warpAffine(img, m, (w, h), borderMode, borderValue)`` (align.py:239-241) and
  ``cv2.warpPerspective`` (align.py:231-233) with the default INTER_LINEAR: the matrix is inverted in
  float64, source coordinates are 1/32-pixel fixed point (affine: rint(M*x*1024) tables + rounding delta 16,
  >> 5; perspective: rint(32 * X/W)), the four taps are blended with weights (32-fx)(32-fy)... : uint8 in
  15-bit fixed point ((sum*32 + 2^14) >> 15), uint16 in float32 (v0*w0 + v1*w1 + v2*w2 + v3*w3, left to
  right, rint);
* the mask of ``BORDER_REPLICATE_BLUR`` (align.py:234-236,242-244,248): warp of a uint8 image of ones with
  a constant 0 border, i.e. mask != 0 <=> the in-image taps carry at least half of the weight;
* ``cv2.GaussianBlur(img_warp, (21, 21), sigmaX=border_blur)`` (align.py:249-250): cv2's fixed-point path,
  (sum_ij k_i k_j p_ij + half) >> shift with an integer kernel (8 fractional bits for uint8, 16 for uint16)
  that sums to exactly 1.0; the kernel itself is read out of the installed cv2 (:func:`fixed_point_kernel`,
  the host mirror does the same) -- and the paste ``img_warp[mask == 0] = blurred[mask == 0]`` (align.py:251).

Parity status: pinned bit-exact against cv2 4.13.0 (the library the reference calls) for warpAffine
uint8/uint16 and warpPerspective uint8, all three border modes, by ``tests/test_align_cpu.py``.  NOT pinned:
warpPerspective on uint16 -- cv2 4.13 routes that combination through different (float) kernels which,
with BORDER_REPLICATE, leave out-of-image pixels unwritten; :func:`warp` models it as float32 bilinear
interpolation with a replicated border (agrees with cv2 to 1 LSB on in-image pixels of smooth content).
"""
import numpy as np

BORDER_CONSTANT, BORDER_REPLICATE, BORDER_REPLICATE_BLUR = 0, 1, 2


def invert_affine(m):
    """cv2.warpAffine's in-place inversion of the 2x3 matrix (float64)."""
    M = np.array(m, np.float64).reshape(6).copy()
    D = M[0] * M[4] - M[1] * M[3]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[4] * D, M[0] * D
    M[0] = A11
    M[1] *= -D
    M[3] *= -D
    M[4] = A22
    b1 = -M[0] * M[2] - M[1] * M[5]
    b2 = -M[3] * M[2] - M[4] * M[5]
    M[2], M[5] = b1, b2
    return M


def invert_homography(m):
    """cv2.invert of a 3x3 float64 matrix (closed form via the adjugate, as cv2 does for n <= 3)."""
    S = np.array(m, np.float64).reshape(3, 3)
    d = S[0, 0] * (S[1, 1] * S[2, 2] - S[1, 2] * S[2, 1]) - S[0, 1] * (S[1, 0] * S[2, 2] - S[1, 2] * S[2, 0]) + \
        S[0, 2] * (S[1, 0] * S[2, 1] - S[1, 1] * S[2, 0])
    d = 1.0 / d
    return np.array([
        (S[1, 1] * S[2, 2] - S[1, 2] * S[2, 1]) * d, (S[0, 2] * S[2, 1] - S[0, 1] * S[2, 2]) * d,
        (S[0, 1] * S[1, 2] - S[0, 2] * S[1, 1]) * d, (S[1, 2] * S[2, 0] - S[1, 0] * S[2, 2]) * d,
        (S[0, 0] * S[2, 2] - S[0, 2] * S[2, 0]) * d, (S[0, 2] * S[1, 0] - S[0, 0] * S[1, 2]) * d,
        (S[1, 0] * S[2, 1] - S[1, 1] * S[2, 0]) * d, (S[0, 1] * S[2, 0] - S[0, 0] * S[2, 1]) * d,
        (S[0, 0] * S[1, 1] - S[0, 1] * S[1, 0]) * d])


def source_coords(m, h, w):
    """1/32-pixel fixed-point source coordinates (X, Y) of every destination pixel, int64 h x w."""
    m = np.asarray(m, np.float64)
    xs = np.arange(w, dtype=np.float64)
    ys = np.arange(h, dtype=np.float64)[:, None]
    if m.size == 6:
        Mi = invert_affine(m)
        ad = np.rint(Mi[0] * xs * 1024).astype(np.int64)
        bd = np.rint(Mi[3] * xs * 1024).astype(np.int64)
        X0 = np.rint((Mi[1] * ys + Mi[2]) * 1024).astype(np.int64) + 16
        Y0 = np.rint((Mi[4] * ys + Mi[5]) * 1024).astype(np.int64) + 16
        return (X0 + ad) >> 5, (Y0 + bd) >> 5
    Mi = invert_homography(m)
    X0 = Mi[0] * xs + (Mi[1] * ys + Mi[2])
    Y0 = Mi[3] * xs + (Mi[4] * ys + Mi[5])
    W = Mi[6] * xs + (Mi[7] * ys + Mi[8])
    with np.errstate(divide="ignore"):
        W = np.where(W != 0, 32.0 / W, 0.0)
    fX = np.clip(X0 * W, -2.0 ** 31, 2.0 ** 31 - 1)
    fY = np.clip(Y0 * W, -2.0 ** 31, 2.0 ** 31 - 1)
    return np.rint(fX).astype(np.int64), np.rint(fY).astype(np.int64)


def _fma32(a, b, c):
    """float32 fused multiply-add: the product of two float32 values is exact in float64."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def _float_perspective_u16(img, m, border, bval):
    """cv2 4.13 warpPerspective of uint16 images (found by probing cv2, tests/test_align_cpu.py pins it): quotient
    X/W and its fractional part in float64, the fraction as float32, vertical interpolation first, every step one
    fused multiply-add, round half to even.  Where one of the four taps lies outside the image cv2 writes the border
    value (BORDER_CONSTANT) or leaves the destination untouched (BORDER_REPLICATE); the latter is modelled as
    replication of the edge pixel."""
    h, w = img.shape[:2]
    f32 = np.float32
    Mi = invert_homography(m)
    xs = np.arange(w, dtype=np.float64)
    ys = np.arange(h, dtype=np.float64)[:, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        W = Mi[6] * xs + (Mi[7] * ys + Mi[8])
        X = (Mi[0] * xs + (Mi[1] * ys + Mi[2])) / W
        Y = (Mi[3] * xs + (Mi[4] * ys + Mi[5])) / W
    X = np.where(X > -1e9, np.where(X < 1e9, X, 1e9), -1e9)
    Y = np.where(Y > -1e9, np.where(Y < 1e9, Y, 1e9), -1e9)
    sx = np.clip(np.floor(X), -32768, 32767).astype(np.int64)
    sy = np.clip(np.floor(Y), -32768, 32767).astype(np.int64)
    fx = (X - np.floor(X)).astype(f32)[..., None]
    fy = (Y - np.floor(Y)).astype(f32)[..., None]

    def px(yy, xx):
        return img[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)].astype(f32)
    p00, p01, p10, p11 = px(sy, sx), px(sy, sx + 1), px(sy + 1, sx), px(sy + 1, sx + 1)
    fyb, fxb = np.broadcast_to(fy, p00.shape), np.broadcast_to(fx, p00.shape)
    left = _fma32(fyb, p10 - p00, p00)
    right = _fma32(fyb, p11 - p01, p01)
    out = np.clip(np.rint(_fma32(fxb, right - left, left)), 0, 65535).astype(np.uint16)
    if border == BORDER_CONSTANT:
        all_in = (sx >= 0) & (sx + 1 < w) & (sy >= 0) & (sy + 1 < h)
        out[~all_in] = bval.astype(np.uint16)
    return out


def float_perspective_all_taps_inside(m, h, w):
    """Mask of the destination pixels whose four taps lie in the image (where cv2's uint16 warpPerspective is defined
    for every border mode)."""
    Mi = invert_homography(m)
    xs = np.arange(w, dtype=np.float64)
    ys = np.arange(h, dtype=np.float64)[:, None]
    W = Mi[6] * xs + (Mi[7] * ys + Mi[8])
    sx = np.floor((Mi[0] * xs + (Mi[1] * ys + Mi[2])) / W)
    sy = np.floor((Mi[3] * xs + (Mi[4] * ys + Mi[5])) / W)
    return (sx >= 0) & (sx + 1 < w) & (sy >= 0) & (sy + 1 < h)


def warp(img, m, border=BORDER_REPLICATE, border_value=(0, 0, 0, 0), return_mask=False):
    """cv2.warpAffine (m 2x3) / cv2.warpPerspective (m 3x3) of an H x W x 3 uint8/uint16 image onto the same
    size, INTER_LINEAR.  ``border`` CONSTANT or REPLICATE (REPLICATE_BLUR warps like REPLICATE).  With
    ``return_mask`` also the uint8 mask that align_images derives from the warp of ones."""
    m = np.asarray(m, np.float64)
    h, w = img.shape[:2]
    const = border == BORDER_CONSTANT
    info = np.iinfo(img.dtype)
    bval = np.clip(np.rint(np.asarray(border_value[:3], np.float64)), info.min, info.max).astype(np.int64)
    X, Y = source_coords(m, h, w)
    sx = np.clip(X >> 5, -32768, 32767)
    sy = np.clip(Y >> 5, -32768, 32767)
    fx, fy = X & 31, Y & 31
    taps = [(sy, sx, (32 - fx) * (32 - fy)), (sy, sx + 1, fx * (32 - fy)),
            (sy + 1, sx, (32 - fx) * fy), (sy + 1, sx + 1, fx * fy)]
    inside_weight = np.zeros((h, w), np.int64)
    if img.dtype == np.uint8:
        acc = np.zeros((h, w, 3), np.int64)
    else:
        acc = np.zeros((h, w, 3), np.float32)
    for yy, xx, wt in taps:
        inside = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        inside_weight += np.where(inside, wt, 0)
        v = img[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)].astype(np.int64)
        if const:
            v[~inside] = bval
        if img.dtype == np.uint8:
            acc += v * wt[..., None]
        else:
            acc = acc + v.astype(np.float32) * (wt.astype(np.float32) / np.float32(1024))[..., None]
    if img.dtype == np.uint8:
        out = np.clip((acc * 32 + 16384) >> 15, 0, 255).astype(np.uint8)
    elif m.size == 9:
        out = _float_perspective_u16(img, m, border, bval.astype(np.float32))
    else:
        out = np.clip(np.rint(acc), 0, 65535).astype(np.uint16)
    if return_mask:
        return out, (inside_weight >= 512).astype(np.uint8)
    return out


def fixed_point_kernel(dtype, ksize, sigma):
    """The integer Gaussian kernel cv2.GaussianBlur uses for uint8 (8 fractional bits) / uint16 (16 bits),
    read out of cv2 itself with an impulse: all taps but the centre are returned exactly by a (ksize, 1) blur
    of a full-scale impulse, the centre is what is missing to 1.0 (cv2 normalises the kernel that way)."""
    import cv2  # pylint: disable=import-outside-toplevel
    dtype = np.dtype(dtype)
    one = 256 if dtype == np.uint8 else 65536
    imp = np.zeros((1, 4 * ksize + 1), dtype)
    c = 2 * ksize
    imp[0, c] = np.iinfo(dtype).max
    r = cv2.GaussianBlur(imp, (ksize, 1), sigmaX=sigma, sigmaY=sigma)[0].astype(np.int64)
    k = r[c - ksize // 2:c + ksize // 2 + 1].copy()
    k[ksize // 2] = one - (k.sum() - k[ksize // 2])
    return k


def gaussian_blur_fixed(img, kernel):
    """cv2.GaussianBlur on uint8/uint16 through its fixed-point path, REFLECT101."""
    k = [int(v) for v in kernel]
    r = len(k) // 2
    h, w = img.shape[:2]
    ap = np.pad(img.astype(np.int64), ((r, r), (r, r), (0, 0)), mode="reflect")
    rows = sum(k[j] * ap[:, j:j + w] for j in range(len(k)))
    tot = sum(k[i] * rows[i:i + h] for i in range(len(k)))
    shift = 16 if img.dtype == np.uint8 else 32
    return ((tot + (1 << (shift - 1))) >> shift).astype(img.dtype)


def warp_and_fill(img, m, border=BORDER_REPLICATE_BLUR, border_value=(0, 0, 0, 0), blur_kernel=None):
    """The image part of align_images after the transform is known (align.py:228-251)."""
    if border != BORDER_REPLICATE_BLUR:
        return warp(img, m, border, border_value)
    out, mask = warp(img, m, BORDER_REPLICATE, border_value, return_mask=True)
    blurred = gaussian_blur_fixed(out, blur_kernel)
    out[mask == 0] = blurred[mask == 0]
    return out

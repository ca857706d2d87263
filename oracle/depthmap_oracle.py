"""CPU oracle for the depth-map focus stacker -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module;
nothing under ``shinestacker_b200/`` does.

numpy restatement (no cv2 calls) of ``DepthMapStack.focus_stack`` of lucalista/shinestacker
(``src/shinestacker/algorithms/depth_map.py:62-123``) and of the cv2 4.13.0 primitives it calls:

====================================  =======================================================  ==========
reference call                        restated by                                              vs cv2 here
====================================  =======================================================  ==========
cv2.cvtColor(BGR2GRAY) u8/u16         :func:`bgr2gray` (fixed point, 15-bit)                   bit-exact
cv2.GaussianBlur(f32,(k,k),0)         :func:`gaussian_blur`                                    bit-exact
cv2.Laplacian(f32,CV_64F,ksize)       :func:`laplacian`                                        exact for u8 stacks, <= 1 ulp u16
cv2.Sobel(f32,CV_64F,..,ksize=3)      :func:`sobel_abs_sum`                                    bit-exact
cv2.bilateralFilter(f32,d,25,25)      :func:`bilateral`                                        <= 1 ulp
cv2.pyrDown / cv2.pyrUp (f32)         :func:`pyr_down` / :func:`pyr_up`                        <= 1 ulp
====================================  =======================================================  ==========

DepthMapStack blends with *soft* weights, so there is no integer map to pin; the parity contract
 is the fused image within 1 LSB of the reference.  Parity status: pinned by
live differential runs against the unmodified reference in the build container
(``tests/test_depthmap_oracle.py``) and by the golden vectors in ``tests/golden/depthmap`` (reference
outputs, generator ``tests/golden/make_golden_depthmap.py``); the reference's own test
(tests/test_0061_depth_map.py) only asserts shapes / dtypes.

One documented deviation: for ``map_type='average'`` the reference evaluates
``np.divide(energies, sum, where=sum != 0)`` without ``out=`` (depth_map.py:56), which leaves the weights
of pixels whose energy is zero in *every* frame uninitialised; this restatement (and the CUDA path) gives
those pixels weight 0.
"""
import numpy as np

f32 = np.float32

# reference src/shinestacker/config/constants.py:144-157
DM_ENERGY_LAPLACIAN = "laplacian"
DM_ENERGY_SOBEL = "sobel"
DM_MAP_AVERAGE = "average"
DM_MAP_MAX = "max"


def bgr2gray(img):
    """cv2.cvtColor(img, COLOR_BGR2GRAY) for uint8/uint16 (utils.img_bw, reference utils.py:45-46):
    (B*3735 + G*19235 + R*9798 + 2^14) >> 15."""
    b, g, r = (img[..., c].astype(np.int64) for c in range(3))
    return ((b * 3735 + g * 19235 + r * 9798 + 16384) >> 15).astype(img.dtype)


def _pad101(a, r):
    return np.pad(a, r, mode="reflect")


def gaussian_kernel(ksize):
    """cv2.getGaussianKernel(ksize, sigma<=0, CV_32F): fixed tables up to 9 taps, otherwise
    sigma = 0.3*((ksize-1)*0.5 - 1) + 0.8, normalised in float64 and rounded to float32."""
    small = {1: [1.0], 3: [0.25, 0.5, 0.25], 5: [0.0625, 0.25, 0.375, 0.25, 0.0625],
             7: [0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125],
             9: [x / 256.0 for x in (4, 13, 30, 51, 60, 51, 30, 13, 4)]}
    if ksize in small:
        return np.array(small[ksize], f32)
    sigma = 0.3 * ((ksize - 1) * 0.5 - 1) + 0.8
    x = np.arange(ksize) - (ksize - 1) * 0.5
    k = np.exp(-0.5 / (sigma * sigma) * x * x)
    return (k / k.sum()).astype(f32)


def _sym_filter_1d(a, k, axis):
    """cv2's symmetric / anti-symmetric small-kernel form along one axis (REFLECT101):
    k[c]*x0 + sum_j k[c+j]*(x[-j] +- x[+j]), accumulated in tap order from the centre outwards."""
    k = np.asarray(k, f32)
    r = len(k) // 2
    n = a.shape[axis]
    pad = [(0, 0), (0, 0)]
    pad[axis] = (r, r)
    ap = np.pad(a, pad, mode="reflect")

    def tap(o):
        sl = [slice(None), slice(None)]
        sl[axis] = slice(r + o, r + o + n)
        return ap[tuple(sl)]
    anti = r > 0 and k[0] == -k[-1] and k[r] == 0
    acc = tap(0) * k[r]
    for j in range(1, r + 1):
        pair = (tap(j) - tap(-j)) if anti else (tap(j) + tap(-j))
        acc = acc + pair * k[r + j]
    return acc.astype(f32)


def sep_filter(a, kx, ky):
    """cv2.sepFilter2D(a, CV_32F, kx, ky): rows first, then columns, float32 throughout."""
    return _sym_filter_1d(_sym_filter_1d(a.astype(f32), kx, 1), ky, 0)


def gaussian_blur(gray, ksize):
    """cv2.GaussianBlur(gray, (ksize, ksize), 0) on a float32 plane (depth_map.py:39)."""
    k = gaussian_kernel(ksize)
    return sep_filter(gray, k, k)


def _sobel_kernel(ksize, order):
    """cv2 getSobelKernels for ksize > 3: integer binomial construction (smoothing [1,1] ksize-order-1 times,
    then the difference [-1,1] `order` times)."""
    ker = [0] * (ksize + 1)
    ker[0] = 1
    for _ in range(ksize - order - 1):
        old = ker[0]
        for j in range(1, ksize + 1):
            new = ker[j] + ker[j - 1]
            ker[j - 1] = old
            old = new
    for _ in range(order):
        old = -ker[0]
        for j in range(1, ksize + 1):
            new = ker[j - 1] - ker[j]
            ker[j - 1] = old
            old = new
    return np.array(ker[:ksize], f32)


def deriv_kernels(ksize):
    """cv2.getDerivKernels(2, 0, ksize) -> (second-derivative kernel, smoothing kernel), the pair
    cv2.Laplacian builds for ksize > 3 (5: [1,0,-2,0,1] and [1,4,6,4,1])."""
    if ksize < 5 or ksize % 2 == 0 or ksize > 31:
        raise ValueError("laplacian ksize must be odd and in 5..31")
    return _sobel_kernel(ksize, 2), _sobel_kernel(ksize, 0)


def laplacian(img, ksize):
    """cv2.Laplacian(img_f32, CV_64F, ksize=ksize) (depth_map.py:40): computed in float32 (working depth
    of a float32 source), then widened: sepFilter(kd,ks) + sepFilter(ks,kd) for ksize 5/7."""
    img = img.astype(f32)
    if ksize in (1, 3):
        # the reference itself fails here: cv2 4.13 has no float32 -> float64 3x3 filter2D
        # ("Unsupported combination of source format (=5), and destination format (=6)")
        raise ValueError("cv2.Laplacian(float32, CV_64F, ksize<=3) is unsupported by the reference's cv2")
    kd, ks = deriv_kernels(ksize)
    return sep_filter(img, kd, ks) + sep_filter(img, ks, kd)


def sobel_abs_sum(gray):
    """|Sobel_x| + |Sobel_y|, ksize 3 (get_sobel_map, depth_map.py:28-35)."""
    g = gray.astype(f32)
    sx = sep_filter(g, [-1, 0, 1], [1, 2, 1])
    sy = sep_filter(g, [1, 2, 1], [-1, 0, 1])
    return np.abs(sx) + np.abs(sy)


def bilateral(e, d, sigma_color=25.0, sigma_space=25.0):
    """cv2.bilateralFilter(e_f32, d, 25, 25) (smooth_energy, depth_map.py:43-51): disc of radius d//2,
    weight = exp(-r^2/(2 ss^2)) * exp(-(v-v0)^2/(2 sc^2)), REFLECT101; unchanged when the plane is flat."""
    e = e.astype(f32)
    radius = max(d // 2, 1)
    mn, mx = float(e.min()), float(e.max())
    if abs(mn - mx) < np.finfo(f32).eps:
        return e.copy()
    gc = -0.5 / (sigma_color * sigma_color)
    gs = -0.5 / (sigma_space * sigma_space)
    ap = _pad101(e, radius)
    h, w = e.shape
    num = np.zeros_like(e)
    den = np.zeros_like(e)
    for i in range(-radius, radius + 1):
        for j in range(-radius, radius + 1):
            r2 = i * i + j * j
            if np.sqrt(float(r2)) > radius:
                continue
            sw = f32(np.exp(r2 * gs))
            val = ap[radius + i:radius + i + h, radius + j:radius + j + w]
            dlt = (val - e).astype(np.float64)
            wgt = (sw * np.exp(dlt * dlt * gc)).astype(f32)
            num = num + val * wgt
            den = den + wgt
    return num / den


def pyr_down(a):
    """cv2.pyrDown on a float32 plane: [1,4,6,4,1]^2/256, REFLECT101, output ceil(n/2)."""
    a = a.astype(f32)
    h, w = a.shape
    ho, wo = (h + 1) // 2, (w + 1) // 2
    ap = _pad101(a, 2)
    ix = 2 * np.arange(wo) + 2
    row = ap[:, ix] * f32(6) + (ap[:, ix - 1] + ap[:, ix + 1]) * f32(4) + ap[:, ix - 2] + ap[:, ix + 2]
    iy = 2 * np.arange(ho) + 2
    return (row[iy] * f32(6) + (row[iy - 1] + row[iy + 1]) * f32(4) + row[iy - 2] + row[iy + 2]) * f32(1 / 256)


def pyr_up(a, dh, dw):
    """cv2.pyrUp(a, dstsize=(dw, dh)) on a float32 plane (dw in {2w-1, 2w}): even outputs
    (l + 6c + r)/64-style, odd outputs 4(c + r); coarse index -1 -> 1 (reflect), n -> n-1 (replicate)."""
    a = a.astype(f32)
    h, w = a.shape

    def ext(x, axis):
        lo = np.take(x, [1 if x.shape[axis] > 1 else 0], axis=axis)
        hi = np.take(x, [x.shape[axis] - 1], axis=axis)
        return np.concatenate([lo, x, hi], axis=axis)
    ap = ext(a, 1)
    c, l, r = ap[:, 1:-1], ap[:, :-2], ap[:, 2:]
    row = np.empty((h, 2 * w), f32)
    row[:, 0::2] = l + c * f32(6) + r
    row[:, 1::2] = (c + r) * f32(4)
    rp = ext(row, 0)
    c, u, d = rp[1:-1], rp[:-2], rp[2:]
    out = np.empty((2 * h, 2 * w), f32)
    out[0::2] = (u + c * f32(6) + d) * f32(1 / 64)
    out[1::2] = ((c + d) * f32(4)) * f32(1 / 64)
    return out[:dh, :dw]


def energy_maps(frames, energy=DM_ENERGY_LAPLACIAN, kernel_size=5, blur_size=5):
    """Gray + focus energy of every frame, float32 (depth_map.py:67-86)."""
    out = []
    for img in frames:
        gray = bgr2gray(img).astype(f32)
        if energy == DM_ENERGY_SOBEL:
            out.append(sobel_abs_sum(gray))
        elif energy == DM_ENERGY_LAPLACIAN:
            out.append(np.abs(laplacian(gaussian_blur(gray, blur_size), kernel_size)))
        else:
            raise ValueError("energy")
    return np.array(out, dtype=f32)


def focus_map(energies, map_type=DM_MAP_AVERAGE, temperature=0.1):
    """get_focus_map (depth_map.py:53-62); frames are summed sequentially in float32."""
    n = energies.shape[0]
    if map_type == DM_MAP_AVERAGE:
        s = energies[0].copy()
        for i in range(1, n):
            s = s + energies[i]
        out = np.zeros_like(energies)
        np.divide(energies, s, out=out, where=s != 0)
        return out
    if map_type == DM_MAP_MAX:
        m = np.max(energies, axis=0)
        rel = np.exp((energies - m) / f32(temperature)).astype(f32)
        s = rel[0].copy()
        for i in range(1, n):
            s = s + rel[i]
        return rel / s
    raise ValueError("map_type")


def focus_stack(frames, map_type=DM_MAP_AVERAGE, energy=DM_ENERGY_LAPLACIAN, kernel_size=5, blur_size=5,
                smooth_size=15, temperature=0.1, levels=3, return_stages=False):
    """DepthMapStack.focus_stack (depth_map.py:64-123) on in-memory BGR uint8/uint16 frames, float-32."""
    frames = list(frames)
    dtype = frames[0].dtype
    e_raw = energy_maps(frames, energy, kernel_size, blur_size)
    mx = e_raw.max()
    e = e_raw / mx if mx > 0 else e_raw
    if smooth_size > 0:
        e = np.array([bilateral(x, smooth_size) for x in e], dtype=f32)
    weights = focus_map(e, map_type, temperature)
    blended = None
    for img, wgt in zip(frames, weights):
        gp_img = [[img[..., c].astype(f32) for c in range(3)]]
        gp_w = [wgt]
        for _ in range(levels - 1):
            gp_img.append([pyr_down(p) for p in gp_img[-1]])
            gp_w.append(pyr_down(gp_w[-1]))
        lp = [gp_img[-1]]
        for j in range(levels - 1, 0, -1):
            hh, ww = gp_img[j - 1][0].shape
            lp.append([fine - pyr_up(coarse, hh, ww) for fine, coarse in zip(gp_img[j - 1], gp_img[j])])
        cur = [[p * gp_w[levels - 1 - j] for p in lp[j]] for j in range(levels)]
        blended = cur if blended is None else [[b + c for b, c in zip(bl, cl)] for bl, cl in zip(blended, cur)]
    result = blended[0]
    for j in range(1, levels):
        hh, ww = blended[j][0].shape
        result = [pyr_up(p, hh, ww) + q for p, q in zip(result, blended[j])]
    res = np.stack(result, axis=-1)
    n_values = 255 if dtype == np.uint8 else 65535
    out = np.clip(np.absolute(res), 0, n_values).astype(dtype)
    if return_stages:
        return out, dict(energy=e_raw, smoothed=e, weights=weights, result=res)
    return out

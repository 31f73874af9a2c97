"""CPU restatement (test infrastructure only) of cv2.warpPerspective(INTER_CUBIC, BORDER_REPLICATE) on uint8 images —
the arithmetic under get_rotate_crop_image (rapid_doc/utils/ocr_utils.py:494-537).  OpenCV is a third-party dependency of the
reference (opencv-python, unpinned); the algorithm restated here is imgwarp.cpp's: the inverse map is evaluated in double
per 32x32-ish block (block origin first, then + M*x1, exactly in that order), scaled by INTER_TAB_SIZE = 32 / W, rounded
half-to-even to a 5-bit fixed-point coordinate; remapBicubic then applies a 4x4 int16 weight table (a = -0.75 cubic,
product of two float 1-D tables, scaled by 2^15, rounded, and corrected so that every 16 weights sum to 32768) and rounds
with (sum + 2^14) >> 15.  Pinned bit-for-bit against cv2 itself on random quads by tests/test_oracle.py (cv2 is installed).
"""
import numpy as np

INTER_BITS = 5
TAB = 1 << INTER_BITS
COEF_BITS = 15
ONE = 1 << COEF_BITS


def cubic_tab_1d():
    """initInterTab1D(INTER_CUBIC): float32 arithmetic as in interpolateCubic."""
    t = np.zeros((TAB, 4), np.float32)
    A = np.float32(-0.75)
    scale = np.float32(1.0) / np.float32(TAB)
    for i in range(TAB):
        x = np.float32(i) * scale
        x1 = x + np.float32(1)
        c0 = ((A * x1 - np.float32(5) * A) * x1 + np.float32(8) * A) * x1 - np.float32(4) * A
        c1 = ((A + np.float32(2)) * x - (A + np.float32(3))) * x * x + np.float32(1)
        xm = np.float32(1) - x
        c2 = ((A + np.float32(2)) * xm - (A + np.float32(3))) * xm * xm + np.float32(1)
        c3 = np.float32(1) - c0 - c1 - c2
        t[i] = [c0, c1, c2, c3]
    return t


def cubic_tab_2d():
    """initInterTab2D(INTER_CUBIC, fixpt=true) -> int16 [32*32][4][4]; index = fy*32 + fx."""
    t1 = cubic_tab_1d()
    out = np.zeros((TAB * TAB, 4, 4), np.int32)
    for i in range(TAB):
        for j in range(TAB):
            v = (t1[i][:, None] * t1[j][None, :]).astype(np.float32)              # vy * vx in float
            it = np.rint(v * np.float32(ONE)).astype(np.int64)                       # saturate_cast<short>(v * 32768): round half even
            it = np.clip(it, -32768, 32767).astype(np.int32)
            isum = int(it.sum())
            if isum != ONE:
                diff = isum - ONE
                Mk = mk = (2, 2)
                for k1 in (2, 3):
                    for k2 in (2, 3):
                        if it[k1, k2] < it[mk]:
                            mk = (k1, k2)
                        elif it[k1, k2] > it[Mk]:
                            Mk = (k1, k2)
                if diff < 0:
                    it[Mk] -= diff
                else:
                    it[mk] -= diff
            out[i * TAB + j] = it
    return out.astype(np.int16)


_TAB2D = None


def warp_perspective_cubic_replicate(src, Minv, dsize):
    """src [H,W,C] uint8, Minv = inverse (dst->src) 3x3 float64 as cv::warpPerspective holds it after invert(); dsize = (w, h)."""
    global _TAB2D
    if _TAB2D is None:
        _TAB2D = cubic_tab_2d()
    H, W, C = src.shape
    dw, dh = int(dsize[0]), int(dsize[1])
    M = np.asarray(Minv, np.float64).reshape(9)
    BLOCK = 32
    bh0 = min(BLOCK // 2, dh)
    bw0 = min(BLOCK * BLOCK // bh0, dw)
    bh0 = min(BLOCK * BLOCK // bw0, dh)
    ys, xs = np.arange(dh), np.arange(dw)
    xb = (xs // bw0) * bw0                       # block origin of every column
    x1 = (xs - xb).astype(np.float64)
    yy = ys.astype(np.float64)[:, None]
    xbf = xb.astype(np.float64)[None, :]
    X0 = M[0] * xbf + M[1] * yy + M[2]           # (a*x + b*y) + c, evaluated left to right, no fused multiply-add
    Y0 = M[3] * xbf + M[4] * yy + M[5]
    W0 = M[6] * xbf + M[7] * yy + M[8]
    Wd = W0 + M[6] * x1[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        Wi = np.where(Wd != 0, TAB / Wd, 0.0)
    lim_lo, lim_hi = float(np.iinfo(np.int32).min), float(np.iinfo(np.int32).max)
    fX = np.maximum(lim_lo, np.minimum(lim_hi, (X0 + M[0] * x1[None, :]) * Wi))
    fY = np.maximum(lim_lo, np.minimum(lim_hi, (Y0 + M[3] * x1[None, :]) * Wi))
    X = np.rint(fX).astype(np.int64)             # cvRound: half to even
    Y = np.rint(fY).astype(np.int64)
    sx = np.clip(X >> INTER_BITS, -32768, 32767) - 1
    sy = np.clip(Y >> INTER_BITS, -32768, 32767) - 1
    a = (Y & (TAB - 1)) * TAB + (X & (TAB - 1))
    w = _TAB2D[a].astype(np.int64)               # [dh, dw, 4, 4]
    out = np.zeros((dh, dw, C), np.int64)
    for ky in range(4):
        iy = np.clip(sy + ky, 0, H - 1)
        for kx in range(4):
            ix = np.clip(sx + kx, 0, W - 1)
            out += src[iy, ix].astype(np.int64) * w[:, :, ky, kx][..., None]
    return np.clip((out + (1 << (COEF_BITS - 1))) >> COEF_BITS, 0, 255).astype(np.uint8)


def get_rotate_crop_image(img, points):
    """rapid_doc/utils/ocr_utils.py:494-537 with cv2.warpPerspective replaced by the restatement above
    (the tiny 3x3 solves stay on cv2: getPerspectiveTransform + cv::invert's closed form)."""
    import cv2
    points = np.asarray(points, np.float32)
    cw = int(max(np.linalg.norm(points[0] - points[1]), np.linalg.norm(points[2] - points[3])))
    ch = int(max(np.linalg.norm(points[0] - points[3]), np.linalg.norm(points[1] - points[2])))
    std = np.float32([[0, 0], [cw, 0], [cw, ch], [0, ch]])
    M = cv2.getPerspectiveTransform(points, std)
    Minv = cv2.invert(M)[1]
    dst = warp_perspective_cubic_replicate(img, Minv, (cw, ch))
    if dst.shape[0] * 1.0 / dst.shape[1] >= 2:
        dst = np.rot90(dst)
    return dst

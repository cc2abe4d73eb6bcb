"""NumPy restatement of the coverage rules of ``bsg_render`` (csrc/render.cu).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference draws its frames with pygame (`_render_frame`, e.g. /root/reference/bluesky_gym/envs/horizontal_cr_env.py:277-395)
and pygame is not in the image, so pygame's own rasteriser cannot be the checker: parity of the *draw calls* is pinned to the
reference by tests/golden/ref_render.npz (make_render_golden.py records them from the reference's code), and this file
restates how a draw call covers pixels -- the rules written in include/bsg.h -- in float64, with a mask of the pixels whose
decision lies within ``band`` pixels of a primitive's edge (where float32 on the device may decide the other way).
"""
import numpy as np

LINE, RING, RECT, EDGE, EDGE_END = 1, 2, 3, 4, 5


def paint(prims, width, height, background, band=1e-3):
    """prims [n, 9] = (kind, x0, y0, x1, y1, w, r, g, b), later rows over earlier ones.
    Returns (rgb [height, width, 3] uint8, unsure [height, width] bool)."""
    y, x = np.meshgrid(np.arange(height) + 0.5, np.arange(width) + 0.5, indexing="ij")
    img = np.empty((height, width, 3), dtype=np.uint8)
    img[:] = np.asarray(background, dtype=np.uint8)
    unsure = np.zeros((height, width), dtype=bool)
    parity = np.zeros((height, width), dtype=bool)
    for p in np.asarray(prims, dtype=np.float64):
        kind, x0, y0, x1, y1, w = int(p[0]), *p[1:6]
        hit = None
        if kind == LINE:
            dx, dy = x1 - x0, y1 - y0
            l2 = dx * dx + dy * dy
            t = np.clip(((x - x0) * dx + (y - y0) * dy) / l2, 0.0, 1.0) if l2 > 0 else np.zeros_like(x)
            d = np.hypot(x - (x0 + t * dx), y - (y0 + t * dy))
            hw = 0.5 * max(w, 1.0)
            hit = d <= hw
            near = np.abs(d - hw) < band
        elif kind == RING:
            d = np.hypot(x - x0, y - y0)
            r, ri = x1, max(x1 - w, 0.0)
            hit = (d <= r) & ((w <= 0) | (d >= ri))
            near = (np.abs(d - r) < band) | ((w > 0) & (np.abs(d - ri) < band))
        elif kind == RECT:
            hit = (x >= x0) & (x < x1) & (y >= y0) & (y < y1)
            near = np.zeros_like(hit)
            for edge, c in ((x0, x), (x1, x), (y0, y), (y1, y)):
                near |= np.abs(c - edge) < band
        elif kind in (EDGE, EDGE_END):
            straddle = (y0 > y) != (y1 > y)
            with np.errstate(divide="ignore", invalid="ignore"):
                xc = (x1 - x0) * (y - y0) / (y1 - y0) + x0
            parity ^= straddle & (x < xc)
            unsure |= (straddle & (np.abs(x - xc) < band)) | (np.abs(y - y0) < band) | (np.abs(y - y1) < band)
            if kind == EDGE_END:
                hit, near = parity.copy(), np.zeros_like(parity)
                parity[:] = False
        if hit is not None:
            img[hit] = p[6:9].astype(np.uint8)
            unsure |= near
    return img, unsure

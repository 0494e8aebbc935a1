"""Host emulation of the data movement of sepfilter_vwalk_kernel (kornia_b200/csrc/sepfilter_vwalk.cuh; with TW = 64 the
geometry of ssim_vwalk_kernel, ssim_vwalk.cuh, whose five planes move exactly like this one): the TMA boxes
(zero fill outside the image), the horizontal patch, the prologue / tile sequence of a band segment, the carried rows,
the vertical patch on the row-filtered rows and the thread -> (row, quad) maps of the row pass and of the carry -- phase
by phase as the barriers order them, against a plain padded separable filter.  A design check that runs without a GPU
(the kernel itself has not run on hardware yet: DESIGN.md section 9); arithmetic is float64 here, so it checks indices,
not rounding.   python tools/emulate_vwalk.py"""
import itertools

import numpy as np

TH, XPAD = 32, 8


def border_index(q, n, border):
    if 0 <= q < n:
        return q
    if border == "constant":
        return -1
    if border == "replicate":
        return 0 if q < 0 else n - 1
    return -q if q < 0 else 2 * (n - 1) - q  # reflect


def tma_box(img, x, y, w, h):
    H, W = img.shape
    box = np.zeros((h, w))
    ys, xs = np.arange(y, y + h), np.arange(x, x + w)
    my, mx = (ys >= 0) & (ys < H), (xs >= 0) & (xs < W)
    box[np.ix_(my, mx)] = img[np.ix_(ys[my], xs[mx])]
    return box


def emulate(img, kx, ky, border, split, TW=128):
    """TW = 128: sepfilter_vwalk_kernel (32 quads x 8 rows per sweep, 4 sweeps); TW = 64: ssim_vwalk_kernel (16 x 16, 2 sweeps)."""
    H, W = img.shape
    BW, QUADS = TW + 2 * XPAD, TW // 4
    RPS = 256 // QUADS
    RY = TH * (TW // 2) // 256
    K = len(kx)
    HALO, CARRY = (K - 1) // 2, K - 1
    MH = TH + CARRY
    COL0 = XPAD - HALO
    bands, tiles_y = -(-W // TW), -(-H // TH)
    out = np.full((H, W), np.nan)
    for band in range(bands):
        x0, ox = band * TW, band * TW - XPAD
        nl, nr = max(-ox, 0), max(ox + BW - W, 0)
        cuts = sorted(set([0, tiles_y] + [c for c in split if 0 < c < tiles_y]))
        for t0, t1 in zip(cuts[:-1], cuts[1:]):       # segments of this band (as the Segments scheduler may cut them)
            mid = np.full((MH, TW), np.nan)           # garbage between segments
            for t in range(t0 - 1, t1):
                pro = t < t0
                rows = CARRY if pro else TH
                tile = np.full((TH, BW), np.nan)
                tile[:rows] = tma_box(img, ox, (t0 * TH - HALO) if pro else (t * TH + HALO), BW, rows)
                if border != "constant" and nl + nr > 0:
                    ncols = nl + nr
                    src = tile.copy()  # reads are of in-image columns, never written by the patch
                    for e in range(rows * ncols):
                        r, k = divmod(e, ncols)
                        c = k if k < nl else BW - nr + (k - nl)
                        sc = border_index(ox + c, W, border) - ox
                        if 0 <= sc < BW:
                            tile[r, c] = src[r, sc]
                # row pass: thread (rq, rr), sweeps it
                base = 0 if pro else CARRY
                for rr, rq, it in itertools.product(range(RPS), range(QUADS), range(TH // RPS)):
                    row = rr + it * RPS
                    if row < rows:
                        for o in range(4):
                            mid[base + row, 4 * rq + o] = sum(kx[j] * tile[row, COL0 + 4 * rq + o + j] for j in range(K))
                if pro:
                    continue
                y0, gy0 = t * TH, t * TH - HALO
                if border != "constant" and (gy0 < 0 or gy0 + MH > H):
                    src = mid.copy()
                    for mr in range(MH):
                        gy = gy0 + mr
                        if 0 <= gy < H:
                            continue
                        sr = border_index(gy, H, border) - gy0
                        if 0 <= sr < MH:
                            mid[mr] = src[sr]
                # column pass: thread (cp, yb)
                for yb, cp in itertools.product(range(TH // RY), range(TW // 2)):
                    for o in range(RY):
                        y, x = y0 + yb * RY + o, x0 + 2 * cp
                        if y < H and x < W:
                            out[y, x:x + 2] = sum(ky[i] * mid[yb * RY + o + i, 2 * cp:2 * cp + 2] for i in range(K))
                if t + 1 < t1:
                    nxt = mid.copy()
                    for rr, rq, it in itertools.product(range(RPS), range(QUADS), range(TH // RPS)):
                        R = CARRY + rr + RPS * it
                        if R >= TH:
                            nxt[R - TH, 4 * rq:4 * rq + 4] = mid[R, 4 * rq:4 * rq + 4]
                    nxt[CARRY:] = np.nan  # overwritten by the next row pass: must not be relied upon
                    mid = nxt
    return out


def reference(img, kx, ky, border):
    K = len(kx)
    h = (K - 1) // 2
    mode = {"constant": "constant", "reflect": "reflect", "replicate": "edge"}[border]
    pad = np.pad(img, h, mode=mode)
    H, W = img.shape
    mid = sum(kx[j] * pad[:, j:j + W] for j in range(K))
    return sum(ky[i] * mid[i:i + H] for i in range(K))


def main():
    rng = np.random.default_rng(0)
    cases = [(70, 132), (32, 128), (33, 4), (97, 260), (6, 8), (64, 388)]
    for (H, W), K, border, split in itertools.product(cases, (3, 11, 17), ("constant", "reflect", "replicate"), ((), (1,), (2, 3))):
        if border != "constant" and (H <= K // 2 or W <= K // 2):
            continue
        img = rng.random((H, W))
        kx, ky = rng.random(K), rng.random(K)
        want = reference(img, kx, ky, border)
        for tw in (128, 64):
            if tw == 64 and (K > 11 or border != "reflect"):
                continue  # the SSIM kernel: windows up to 11, reflect only
            got = emulate(img, kx, ky, border, split, tw)
            err = np.nanmax(np.abs(got - want)) if not np.isnan(got).any() else float("nan")
            assert err < 1e-12, (H, W, K, border, split, tw, err)
    print("emulation of the band walk agrees with the padded separable filter on every case")


if __name__ == "__main__":
    main()

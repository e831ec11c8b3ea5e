"""Measures the one numeric output the reference publishes of its own solver: the true-scale deformed shape in
/root/reference/examples/linkedin-logo/output.png (readme.md:28-30; drawn by scripts/plot.py:143-147 at
(x + ux, y + uy), no magnification, beside the undeformed mesh).

Nothing of the picture is copied: this script reads it where it lies, finds the gridlines of the two panels
(matplotlib's seaborn-v0_8 style: white lines every 100 units on a (234,234,242) background; the tick labels, read
by eye, say the first vertical line is x = 0 and the first horizontal line from the top is y = 100), and writes the
x-intervals the model covers along horizontal lines and the y-intervals along vertical lines, every 25 units, for
both panels, into tests/golden/reference_linkedin_picture.json — a few hundred numbers with a resolution of one
pixel = 0.73 units — and, for the solved panel, the stress colour on a 12.5-unit grid (red minus blue of the "coolwarm"
face colour, a monotone function of the plotted stress).  tests/test_reference_picture.py compares the oracle's (CPU) and the library's (GPU) solution of
the same example with them.

    python tests/golden/measure_reference_picture.py        # needs /root/reference and Pillow
"""
import json
from pathlib import Path

import numpy as np
from PIL import Image

PICTURE = Path("/root/reference/examples/linkedin-logo/output.png")
OUT = Path(__file__).resolve().parent / "reference_linkedin_picture.json"
BACKGROUND = np.array([234, 234, 242])
STEP = 25
MIN_RUN = 3            # pixels: shorter runs of model / gap are anti-aliasing, not geometry


def runs_of(mask):
    """[(first, last)] index runs of True."""
    idx = np.flatnonzero(mask)
    if not len(idx):
        return []
    cut = np.flatnonzero(np.diff(idx) > 1)
    starts = np.concatenate([[idx[0]], idx[cut + 1]])
    ends = np.concatenate([idx[cut], [idx[-1]]])
    return list(zip(starts.tolist(), ends.tolist()))


def ranges(mask, min_gap):
    """runs of True, merged across gaps shorter than min_gap."""
    out = []
    for a, b in runs_of(mask):
        if out and a - out[-1][1] < min_gap:
            out[-1] = (out[-1][0], b)
        else:
            out.append((a, b))
    return out


def model_intervals(mask_line, to_value):
    """Model-covered intervals along one pixel line, in data units; runs / gaps below MIN_RUN pixels are dropped."""
    rs = [r for r in ranges(mask_line, MIN_RUN) if r[1] - r[0] + 1 >= MIN_RUN]
    return [[round(float(to_value(a)), 2), round(float(to_value(b + 1)), 2)] for a, b in rs]


def main():
    im = np.array(Image.open(PICTURE).convert("RGB")).astype(int)
    H, W, _ = im.shape
    is_bg = np.abs(im - BACKGROUND).sum(2) < 12
    is_white = im.min(2) >= 245
    model = ~(is_bg | is_white)
    row_rng = ranges(is_bg.sum(1) > 0.05 * W, 50)
    col_rng = ranges(is_bg.sum(0) > 0.05 * H, 50)
    assert len(row_rng) == 1 and len(col_rng) == 2, (row_rng, col_rng)
    r0, r1 = row_rng[0]
    result = {"source": "examples/linkedin-logo/output.png of kyle-tennison/Magnetite (the reference's own run of its example)",
              "made_by": "tests/golden/measure_reference_picture.py", "picture_size": [W, H], "step": STEP, "panels": {}}
    for name, (c0, c1) in zip(("solved", "initial"), col_rng):
        # gridlines: white runs in the background strip under / left of the model
        # (centres averaged over a band of 8 pixel lines: a line is 2-3 pixels wide and not pixel-aligned)
        band_v = [[0.5 * (a + b) + c0 for a, b in runs_of(is_white[r, c0:c1 + 1]) if b - a < 6] for r in range(r1 - 11, r1 - 3)]
        band_h = [[0.5 * (a + b) + r0 for a, b in runs_of(is_white[r0:r1 + 1, c]) if b - a < 6] for c in range(c0 + 3, c0 + 11)]
        assert len({len(v) for v in band_v}) == 1 and len({len(h) for h in band_h}) == 1
        vx = [round(float(v), 3) for v in np.mean(band_v, axis=0)]
        hy = [round(float(h), 3) for h in np.mean(band_h, axis=0)]
        px_per_unit_x = (vx[-1] - vx[0]) / (100.0 * (len(vx) - 1))
        px_per_unit_y = (hy[-1] - hy[0]) / (100.0 * (len(hy) - 1))
        assert abs(px_per_unit_x - px_per_unit_y) < 0.01 * px_per_unit_x          # set_aspect("equal")
        assert max(abs(np.diff(vx) - 100 * px_per_unit_x)) < 1.5 and max(abs(np.diff(hy) - 100 * px_per_unit_y)) < 1.5
        x_of = lambda px, vx=vx, s=px_per_unit_x: (px - vx[0] - 0.5) / s          # pixel edge -> x (first line: x = 0)
        y_of = lambda px, hy=hy, s=px_per_unit_y: 100.0 - (px - hy[0] - 0.5) / s  # pixel edge -> y (first line: y = 100)
        col_of = lambda x, vx=vx, s=px_per_unit_x: int(round(vx[0] + x * s))
        row_of = lambda y, hy=hy, s=px_per_unit_y: int(round(hy[0] + (100.0 - y) * s))
        panel = {"pixels_per_unit": round(px_per_unit_x, 4), "gridlines_x_px": vx, "gridlines_y_px": hy,
                 "along_y": {}, "along_x": {}}
        sub = model[:, c0:c1 + 1]
        # "at": the coordinate of the centre of the pixel line that was read (the nominal one to within half a pixel)
        for Y in range(-650, 176, STEP):
            r = row_of(Y)
            if r0 < r < r1:
                iv = model_intervals(sub[r], lambda p: x_of(p + c0))
                if iv:
                    panel["along_y"][str(Y)] = {"at": round(float(y_of(r + 0.5)), 3), "intervals": iv}
        for X in range(0, 651, STEP):
            c = col_of(X)
            if c0 < c < c1:
                iv = model_intervals(model[r0:r1 + 1, c], lambda p: y_of(p + r0))
                if iv:
                    panel["along_x"][str(X)] = {"at": round(float(x_of(c + 0.5)), 3),
                                                "intervals": [[b, a] for a, b in iv][::-1]}      # ascending y
        if name == "solved":
            # the stress colours (scripts/plot.py:136-141,154-158: cmap "coolwarm" over [min stress, max stress], faces
            # drawn with alpha 0.7): on a 12.5-unit grid, where the 7x7 pixel patch around the point is all model, the
            # median colour of its brighter half (the darker half is the black mesh lines), un-blended from the
            # background, as red minus blue — a monotone function of the colormap parameter, hence of the stress
            colour = []
            for Y in np.arange(-637.5, 160.0, 12.5):
                for X in np.arange(0.0, 640.0, 12.5):
                    r, c = row_of(Y), col_of(X)
                    if not (r0 + 4 <= r <= r1 - 4 and c0 + 4 <= c <= c1 - 4) or not model[r - 3:r + 4, c - 3:c + 4].all():
                        continue
                    patch = im[r - 3:r + 4, c - 3:c + 4].reshape(-1, 3).astype(float)
                    lum = patch.sum(1)
                    rgb = (np.median(patch[lum >= np.percentile(lum, 50)], axis=0) - 0.3 * BACKGROUND) / 0.7
                    colour.append([round(float(x_of(c + 0.5)), 2), round(float(y_of(r + 0.5)), 2), round(float(rgb[0] - rgb[2]), 1)])
            panel["red_minus_blue"] = colour
        result["panels"][name] = panel
    OUT.write_text(json.dumps(result, separators=(",", ":")) + "\n")
    s, i = result["panels"]["solved"], result["panels"]["initial"]
    print(f"{len(s['red_minus_blue'])} colour samples")
    print(f"{OUT.name}: {len(s['along_y'])} + {len(s['along_x'])} lines of the solved model, "
          f"{len(i['along_y'])} + {len(i['along_x'])} of the initial one, {s['pixels_per_unit']} px per unit")


if __name__ == "__main__":
    main()

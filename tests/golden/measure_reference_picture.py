"""Measures the numeric outputs the reference publishes of its own solver: the TRUE-SCALE deformed shapes in
/root/reference/examples/linkedin-logo/output.png (readme.md:28-30), /root/reference/media/tensilve-results.png
(the tensile example) and /root/reference/examples/cover-eample/output.png (readme.md:1), all drawn by
scripts/plot.py:143-147 at (x + ux, y + uy), no magnification, the first two beside the undeformed mesh.

Nothing of the pictures is copied: this script reads them where they lie, finds the gridlines of the two panels
(matplotlib's seaborn-v0_8 style: white lines on a (234,234,242) background; which values the first lines carry is
read off the tick labels by eye and written into PICTURES below — the undeformed panel, whose geometry is known,
checks it), and writes the intervals the model covers along horizontal and vertical lines, for both panels, into
tests/golden/reference_<name>_picture.json: a few hundred numbers with a resolution of one pixel — and the face
colour of the solved model on a grid (a monotone function of the plotted element stress).  tests/test_reference_picture.py compares the oracle's (CPU) and the library's (GPU)
solutions of the same examples with them.

    python tests/golden/measure_reference_picture.py        # needs /root/reference and Pillow
"""
import json
from pathlib import Path

import numpy as np
from PIL import Image

REFERENCE = Path("/root/reference")
HERE = Path(__file__).resolve().parent
BACKGROUND = np.array([234, 234, 242])
MIN_RUN = 3            # pixels: shorter runs of model / gap are anti-aliasing, not geometry

# x0 / dx: value of the first vertical gridline from the left and the spacing; y0 / dy: first horizontal gridline from
# the TOP and the spacing (tick labels, read by eye); lines_*: where the model is cut; colour_step: grid of the colour samples
PICTURES = {
    "linkedin": dict(path="examples/linkedin-logo/output.png", x0=0.0, dx=100.0, y0=100.0, dy=100.0,
                     lines_y=np.arange(-650.0, 176.0, 25.0), lines_x=np.arange(0.0, 651.0, 25.0), colour_step=12.5),
    "tensile": dict(path="media/tensilve-results.png", x0=-10.0, dx=5.0, y0=4.0, dy=2.0,
                    lines_y=np.arange(-4.75, 4.76, 0.25), lines_x=np.arange(-11.5, 14.51, 0.5), colour_step=0.4),
    # The cover picture is cropped to the solved panel, tick labels cut away.  Its gridlines are 412.4 px apart in x and
    # 82.5 px in y; the clamped bottom edge and the pulled top band keep ux = 0 (input.json), so the model stays 479.66
    # units wide there, which is 1978 px: 4.124 px per unit, i.e. lines every 100 units in x and (equal aspect) every 20
    # in y.  The line through the model's left edge is x = 0; the clamped bottom edge (y = -91.05) lies 11 units under
    # the lowest line, so the lines are y = 0, -20, ... -80 from the top.
    "cover": dict(path="examples/cover-eample/output.png", x0=0.0, dx=100.0, y0=0.0, dy=20.0, panels=("solved",),
                  lines_y=np.arange(-90.0, 10.1, 2.5), lines_x=np.arange(5.0, 480.0, 10.0), colour_step=4.0),
}


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
    return [[round(float(to_value(a)), 4), round(float(to_value(b + 1)), 4)] for a, b in rs]


def measure(name, cfg):
    im = np.array(Image.open(REFERENCE / cfg["path"]).convert("RGB")).astype(int)
    H, W, _ = im.shape
    is_bg = np.abs(im - BACKGROUND).sum(2) < 12
    is_white = im.min(2) >= 245
    model = ~(is_bg | is_white)
    row_rng = ranges(is_bg.sum(1) > 0.05 * W, 50)
    col_rng = ranges(is_bg.sum(0) > 0.05 * H, 50)
    boxes = [(r, c) for r in row_rng for c in col_rng]                  # "Solved Model" is the first panel (left / top)
    panels = cfg.get("panels", ("solved", "initial"))
    assert len(boxes) == len(panels), (row_rng, col_rng)
    result = {"source": f"{cfg['path']} of kyle-tennison/Magnetite (the reference's own run of its example)",
              "made_by": "tests/golden/measure_reference_picture.py", "picture_size": [W, H], "panels": {}}
    for panel_name, ((r0, r1), (c0, c1)) in zip(panels, boxes):
        # gridlines: white runs in the background strip under / left of the model
        # (centres averaged over a band of 8 pixel lines: a line is 2-3 pixels wide and not pixel-aligned)
        band_v = [[0.5 * (a + b) + c0 for a, b in runs_of(is_white[r, c0:c1 + 1]) if b - a < 6] for r in range(r1 - 11, r1 - 3)]
        band_h = [[0.5 * (a + b) + r0 for a, b in runs_of(is_white[r0:r1 + 1, c]) if b - a < 6] for c in range(c0 + 3, c0 + 11)]
        assert len({len(v) for v in band_v}) == 1 and len({len(h) for h in band_h}) == 1
        vx = [round(float(v), 3) for v in np.mean(band_v, axis=0)]
        hy = [round(float(h), 3) for h in np.mean(band_h, axis=0)]
        sx = (vx[-1] - vx[0]) / (cfg["dx"] * (len(vx) - 1))             # pixels per unit
        sy = (hy[-1] - hy[0]) / (cfg["dy"] * (len(hy) - 1))
        assert max(abs(np.diff(vx) - cfg["dx"] * sx)) < 1.5 and max(abs(np.diff(hy) - cfg["dy"] * sy)) < 1.5
        x_of = lambda px: cfg["x0"] + (px - vx[0] - 0.5) / sx             # pixel EDGE coordinate -> x
        y_of = lambda px: cfg["y0"] - (px - hy[0] - 0.5) / sy             # pixel EDGE coordinate -> y
        col_of = lambda x: int(round(vx[0] + (x - cfg["x0"]) * sx))
        row_of = lambda y: int(round(hy[0] + (cfg["y0"] - y) * sy))
        panel = {"pixels_per_unit_x": round(sx, 4), "pixels_per_unit_y": round(sy, 4), "gridlines_x_px": vx,
                 "gridlines_y_px": hy, "along_y": [], "along_x": []}
        # "at": the coordinate of the centre of the pixel line that was read (the nominal one to within half a pixel)
        for Y in cfg["lines_y"]:
            r = row_of(Y)
            if r0 < r < r1:
                iv = model_intervals(model[r, c0:c1 + 1], lambda p: x_of(p + c0))
                if iv:
                    panel["along_y"].append({"at": round(float(y_of(r + 0.5)), 4), "intervals": iv})
        for X in cfg["lines_x"]:
            c = col_of(X)
            if c0 < c < c1:
                iv = model_intervals(model[r0:r1 + 1, c], lambda p: y_of(p + r0))
                if iv:
                    panel["along_x"].append({"at": round(float(x_of(c + 0.5)), 4),
                                             "intervals": [[b, a] for a, b in iv][::-1]})      # ascending y
        if panel_name == "solved" and cfg["colour_step"]:
            # the stress colours (scripts/plot.py:136-141,154-158: the --cmap colormap over [min stress, max stress],
            # faces drawn with alpha 0.7): on a grid, where the 7x7 pixel patch around the point is all model, the
            # median colour of its brighter half (the darker half is the black mesh lines), un-blended from the
            # background.  The default "coolwarm" of the linkedin picture is monotone in red minus blue, the dark-to-
            # bright maps of the other two in red — monotone functions of the plotted stress.
            colour, step = [], cfg["colour_step"]
            for Y in np.arange(y_of(r1) - y_of(r1) % step, y_of(r0), step):
                for X in np.arange(x_of(c0) - x_of(c0) % step, x_of(c1), step):
                    r, c = row_of(Y), col_of(X)
                    if not (r0 + 4 <= r <= r1 - 4 and c0 + 4 <= c <= c1 - 4) or not model[r - 3:r + 4, c - 3:c + 4].all():
                        continue
                    patch = im[r - 3:r + 4, c - 3:c + 4].reshape(-1, 3).astype(float)
                    lum = patch.sum(1)
                    rgb = (np.median(patch[lum >= np.percentile(lum, 50)], axis=0) - 0.3 * BACKGROUND) / 0.7
                    colour.append([round(float(x_of(c + 0.5)), 3), round(float(y_of(r + 0.5)), 3)] + [round(float(v), 1) for v in rgb])
            panel["face_rgb"] = colour
        result["panels"][panel_name] = panel
    out = HERE / f"reference_{name}_picture.json"
    out.write_text(json.dumps(result, separators=(",", ":")) + "\n")
    s, i = result["panels"]["solved"], result["panels"].get("initial", {"along_y": [], "along_x": []})
    print(f"{out.name}: {len(s['along_y'])} + {len(s['along_x'])} lines of the solved model, "
          f"{len(i['along_y'])} + {len(i['along_x'])} of the initial one, {s['pixels_per_unit_x']} x {s['pixels_per_unit_y']} "
          f"px per unit, {len(s.get('face_rgb', []))} colour samples")


if __name__ == "__main__":
    for name, cfg in PICTURES.items():
        measure(name, cfg)

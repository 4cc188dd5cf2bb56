"""Golden vectors for row f3 (explanation output path), generated with the REAL Pillow in the build container.

    python -m oracle.make_golden_vis        # writes tests/golden/vis_upsample.npz

Each case holds uint8 maps and what ``Image.fromarray(m, 'L').resize((W, H), Image.BILINEAR)`` (test.py:35) returns
for them, plus the attention ratios of test.py:43.  The first case takes its maps from the reference's own vis branch
(``vis`` of tests/golden/head_s90_spc3_n81.npz, dumped by oracle/make_golden.py from slot_attention.py:68-80).
TEST INFRASTRUCTURE ONLY.
"""
import json
import os

import numpy as np
import PIL
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")

# name: (count, h, w, out_h, out_w)
CASES = {
    "ref_vis_9x9_to_260": (4, 9, 9, 260, 260),          # the reference's native geometry (train.py:39)
    "rand_7x7_to_224": (4, 7, 7, 224, 224),             # BASELINE geometry
    "rand_7x7_to_200x300": (2, 7, 7, 200, 300),         # non-square image
    "rand_9x7_to_33x65": (2, 9, 7, 33, 65),             # non-square map, odd sizes
    "rand_1x1_to_5x5": (1, 1, 1, 5, 5),
    "rand_7x7_same": (2, 7, 7, 7, 7),                   # no resampling needed: identity
    "rand_14x14_to_7x7": (2, 14, 14, 7, 7),             # down-scaling (support 2)
    "extremes_7x7_to_64": (3, 7, 7, 64, 64),            # all 0, all 255, checkerboard 0/255
}


def main():
    rng = np.random.RandomState(20260117)
    out = {}
    for name, (c, h, w, oh, ow) in CASES.items():
        if name.startswith("ref_vis"):
            maps = np.load(os.path.join(GOLDEN, "head_s90_spc3_n81.npz"))["vis"][:c].astype(np.uint8)
        elif name.startswith("extremes"):
            maps = np.zeros((3, h, w), np.uint8)
            maps[1] = 255
            maps[2] = ((np.add.outer(np.arange(h), np.arange(w)) % 2) * 255).astype(np.uint8)
        else:
            maps = rng.randint(0, 256, (c, h, w)).astype(np.uint8)
        assert maps.shape == (c, h, w)
        heat = np.stack([np.array(Image.fromarray(m, mode="L").resize((ow, oh), resample=Image.BILINEAR), dtype=np.uint8)
                         for m in maps])
        ratios = np.array([float(m.sum()) / float(h * w * 255) for m in maps], np.float64)
        out[name + ".maps"], out[name + ".heat"], out[name + ".ratios"] = maps, heat, ratios
    out["meta"] = np.frombuffer(json.dumps({"pillow": PIL.__version__, "cases": CASES}).encode(), np.uint8)
    np.savez_compressed(os.path.join(GOLDEN, "vis_upsample.npz"), **out)
    print("wrote vis_upsample.npz, Pillow", PIL.__version__)


if __name__ == "__main__":
    main()

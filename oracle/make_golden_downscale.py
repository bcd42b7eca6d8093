"""Generate tests/golden/downscale.pt by running the REAL reference `Downscale` class (k-space truncation + cv2.INTER_CUBIC
+ round + clip, /root/reference/src/acdc_preprocess.py:102-180) on seeded synthetic cine frames.
Run in the build container only:   python -m oracle.make_golden_downscale

acdc_preprocess.py imports nibabel at module level (absent here); it is stubbed with an empty module - the Downscale class
itself needs numpy and cv2 only.  The reference file is executed in place, never copied.

Precision: the reference pins numpy 1.16 (env.yml), whose fftn always computes in double precision; numpy >= 2 computes a
float32 input in single precision, where low-pass values that land within float32 round-off of x.5 round differently.  The
frames are therefore handed to Downscale as float64 - the arithmetic the reference's own environment performs."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

from oracle.load_reference import REFERENCE_ROOT

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "downscale.pt")


def load_downscale():
    sys.modules.setdefault("nibabel", types.ModuleType("nibabel"))
    spec = importlib.util.spec_from_file_location("ref_acdc_preprocess", os.path.join(REFERENCE_ROOT, "src", "acdc_preprocess.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.Downscale


def frames(h, w, n, seed):
    """integer-valued frames in [0, 255] (what the preprocessing feeds Downscale, acdc_preprocess.py:39-40): smooth blobs +
    texture + a few saturated / zero regions"""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    out = []
    for _ in range(n):
        img = np.zeros((h, w))
        for _ in range(6):
            cy, cx, s = rng.uniform(0, h), rng.uniform(0, w), rng.uniform(3, h / 3)
            img += rng.uniform(30, 200) * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * s * s))
        img += rng.normal(0, 12, (h, w))
        img[: h // 8, : w // 8] = 255
        out.append(np.clip(np.round(img), 0, 255).astype(np.float32))
    return np.stack(out)


def main():
    Downscale = load_downscale()
    cases = []
    for (h, w, r, seed) in [(128, 128, 2, 1), (128, 128, 4, 2), (126, 126, 3, 3), (256, 256, 4, 4), (96, 160, 2, 5), (100, 76, 4, 6)]:
        hr = frames(h, w, 3, seed)
        ds = Downscale(r)
        lr = np.stack([ds(f[..., None].astype(np.float64))[0][..., 0] for f in hr]).astype(np.float32)
        cases.append({"r": r, "hr": torch.from_numpy(hr).to(torch.uint8), "lr": torch.from_numpy(lr).to(torch.uint8)})
        assert (lr == np.round(lr)).all() and lr.min() >= 0 and lr.max() <= 255
    torch.save(cases, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()

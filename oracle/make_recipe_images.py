"""TEST INFRASTRUCTURE (oracle/): packs the reference's test inputs (Test/in_images) and golden outputs
(Test/reference_out_images) that oracle/goldens.py's recipes use into oracle/_ref/recipe_images.npz.

Run where /root/reference exists (the build container; __graft_entry__.build() calls it).  The pack is git-ignored
(oracle/_ref/) -- reference material never enters this repository's history -- but it travels to the GPU box with the
compiled reference, so the GPU tests can run the reference's own recipes through the CUDA engine and compare with the
reference's goldens (tests/test_gpu_golden_recipes.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import goldens  # noqa: E402


def main():
    if not (os.path.isdir(goldens.IN_DIR) and os.path.isdir(goldens.GOLD_DIR)):
        print("make_recipe_images: /root/reference not present, nothing written")
        return 0
    d = {}
    for n in goldens.INPUT_NAMES:
        d["in/" + n] = goldens.load_png(n)
    for n in goldens.CASES:
        d["gold/" + n] = goldens.load_golden(n)
    os.makedirs(os.path.dirname(goldens.PACK), exist_ok=True)
    np.savez_compressed(goldens.PACK, **d)
    print("wrote %s (%d arrays, %.1f MB)" % (goldens.PACK, len(d), os.path.getsize(goldens.PACK) / 1e6))
    # BASELINE.json configs[0] names a real image of Test/in_images: bench.py's "cfg1_brick" sub-record reads this copy
    # (git-ignored like the pack; bench.py never touches oracle/ outside its CPU-baseline legs)
    bdir = os.path.join(ROOT, "baseline", "_ref")
    os.makedirs(bdir, exist_ok=True)
    np.save(os.path.join(bdir, "brick_512.npy"), d["in/brick"])
    return 0


if __name__ == "__main__":
    sys.exit(main())

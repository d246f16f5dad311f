"""Generates tests/golden/golden_hashes.json and tests/golden/ref_vectors.npz from the COMPILED REFERENCE
(oracle/_ref/libref_mt_1t.so, built by oracle/build_ref.sh from /root/reference).  Run in the build container:

    python tests/golden/make_golden_hashes.py

golden_hashes.json : SHA-256 of the reference's output for every case of oracle/goldens.py (each of which equals
                     the reference's own golden PPM), plus whether it matched the golden bit-exactly.
ref_vectors.npz    : small synthetic cases (inputs are regenerated from seeds; outputs stored) run through the
                     compiled reference -- all 9 ordering modes, tiling, the 4 simple-API formats, map channels,
                     alpha, imageSynth2 -- so that the restatement can be checked against the real reference on
                     machines where /root/reference does not exist.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import goldens, refdriver as R  # noqa: E402
from tests.golden import cases  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    lib = R.load("ref_mt_1t")
    outputs, exact = {}, {}
    for name, (is_exact, fn) in goldens.CASES.items():
        out = fn(lib)
        n, mx, _ = goldens.check(lib, name)
        outputs[name] = hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest()
        exact[name] = {"differing_pixels_vs_golden": n, "max_abs_diff": mx, "expected_exact": is_exact}
        print(name, n, mx, flush=True)
    json.dump({"generator": "oracle/_ref/libref_mt_1t.so (compiled reference + GRand MT19937 shim)",
               "outputs": outputs, "vs_reference_golden": exact}, open(os.path.join(HERE, "golden_hashes.json"), "w"),
              indent=1, sort_keys=True)
    vec = {}
    for name, case in cases.all_cases().items():
        vec[name] = cases.run_case(lib, case)
        print(name, vec[name].shape, flush=True)
    np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), **vec)


if __name__ == "__main__":
    main()

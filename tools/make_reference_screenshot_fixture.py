#!/usr/bin/env python
"""Turn the reference's committed screenshot of tests/tie_fighter into a test fixture.

    python tools/make_reference_screenshot_fixture.py [/root/reference]

`tests/tie_fighter/download (10).png` (600x600 RGBA8) is the only output OF THE REFERENCE ITSELF in
the tree whose scene file still matches (the other three screenshots are stale).
The fixture is the decoded RGB array; tests/test_reference_screenshot.py compares the oracle and
the CUDA path with it (the GPU box has no /root/reference, so the decoded array travels instead).
"""
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out = os.path.join(ROOT, "tests", "golden", "reference_screenshot_tie_fighter_600.npz")
    arrays = {}
    for key, name in (("shot10", "download (10).png"), ("shot11", "download (11).png")):
        img = np.asarray(Image.open(os.path.join(ref, "tests", "tie_fighter", name)).convert("RGB"))
        assert img.shape == (600, 600, 3)
        arrays[key] = img
    np.savez_compressed(out, **arrays)
    print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()

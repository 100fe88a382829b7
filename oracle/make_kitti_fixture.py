"""tests/golden/kitti_test_pair.npz: the reference's own test pair (reference/left_test.png, right_test.png, the images
`python inference.py --left_img reference/left_test.png` runs on; BASELINE.json configs[0]) decoded exactly as inference.py:90-91
does (cv2.imread(..., IMREAD_UNCHANGED): uint8 HWC BGR, 375 x 1242) and stored losslessly, so that the GPU box, where
/root/reference does not exist, can run configs[0].  TEST INFRASTRUCTURE.  Run here:  python -m oracle.make_kitti_fixture
"""
import os

import cv2
import numpy as np

REF = "/root/reference/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "kitti_test_pair.npz")

if __name__ == "__main__":
    left = cv2.imread(os.path.join(REF, "left_test.png"), cv2.IMREAD_UNCHANGED)
    right = cv2.imread(os.path.join(REF, "right_test.png"), cv2.IMREAD_UNCHANGED)
    assert left.dtype == np.uint8 and left.shape == right.shape == (375, 1242, 3), (left.shape, right.shape, left.dtype)
    np.savez_compressed(OUT, left_bgr=left, right_bgr=right)
    print(f"wrote {OUT}: {os.path.getsize(OUT) / 1e6:.2f} MB")

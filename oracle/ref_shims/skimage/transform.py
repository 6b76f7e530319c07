"""Stand-in for the two skimage.transform names OOPAO/tools/tools.py:210-217 uses, backed by the
restatement of the scikit-image 0.18.3 algorithm in oracle/warp018.py (PARITY UNPINNED there)."""
import os
import sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.warp018 import warp_translate  # noqa: E402

KERNEL = os.environ.get("ORACLE_WARP_KERNEL", "lagrange018")


class SimilarityTransform:
    def __init__(self, translation=(0, 0), _matrix=None):
        if _matrix is None:
            _matrix = np.eye(3)
            _matrix[0, 2] = translation[0]
            _matrix[1, 2] = translation[1]
        self.params = _matrix

    @property
    def inverse(self):
        return SimilarityTransform(_matrix=np.linalg.inv(self.params))


def warp(image, inverse_map, order=3):
    assert order == 3
    m = inverse_map.params
    assert m[0, 0] == 1 and m[1, 1] == 1 and m[0, 1] == 0 and m[1, 0] == 0
    return warp_translate(image, -m[0, 2], -m[1, 2], kernel=KERNEL)

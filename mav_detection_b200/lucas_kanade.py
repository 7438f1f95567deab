"""LucasKanade placeholder.  The sparse Shi-Tomasi / pyramidal-LK tracker of
/root/reference/src/lucas_kanade.py:9-63 is unused by Processor.run_detection (out of
scope); FocusOfExpansion only reads `old_frame.shape` and `total_num_corners` from it
(focus_of_expansion.py:24-27).  The constructor keeps the reference's draw from the global legacy NumPy
generator (lucas_kanade.py:32) so that the random stream seen by get_FOE_dense is unchanged."""
from __future__ import annotations

import numpy as np


class LucasKanade:
    def __init__(self, old_frame: np.ndarray) -> None:
        self.old_frame = old_frame
        self.num_corners = 2000
        self.minimum_num_corners = self.num_corners // 3
        self.total_num_corners = self.num_corners + self.minimum_num_corners
        self.corners = np.zeros((self.total_num_corners, 2), dtype=np.uint)
        self.num_features = 0
        self.features = []
        self.color = np.random.randint(0, 255, (self.total_num_corners, 3))

    def get_features(self, frame: np.ndarray):
        raise NotImplementedError('the sparse Lucas-Kanade tracker is outside the rebuilt hot path (SURVEY.md §2)')

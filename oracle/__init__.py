"""CPU oracle for the mav-detection hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in ``mav_detection_b200/`` may import this package.  The only callers
allowed are ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` (where it is the thing *compared
against*, never the thing shipped).

Contents
--------
farneback_np   stage-wise NumPy restatement of OpenCV's CPU
               ``calcOpticalFlowFarneback`` (third-party: opencv-python,
               unpinned in /root/reference/requirements.txt:4; the call site is
               /root/reference/src/farneback.py:76-80).  Pinned against
               cv2 4.13.0 run in this container (tests/test_oracle_farneback.py
               and the committed vectors under tests/golden/).
detect_np      NumPy restatement of derotate / get_FOE_dense / ransac / get_phi /
               the mask block / bbox / tpr-fpr
               (/root/reference/src/detector.py:70-117,
               focus_of_expansion.py:32-86,150-184, processor.py:306-362,
               im_helpers.py:55-84,150-159,244-252, utils.py:183-197).
               Pinned against the reference modules imported in this container
               (tests/golden/make_golden.py writes the vectors).
vis_np         the HSV visualisation Farneback.process() returns
               (/root/reference/src/farneback.py:83-99): the reference's own cv2
               statement sequence, plus a NumPy restatement of the arithmetic cv2
               4.13.0 performs, pinned bit-exactly against it
               (tests/test_oracle_vis.py).
ccl_np         8-connected component labelling with raster-first-appearance
               label numbering.  NOT a reference feature: parity
               for labels is "unpinned by the reference"; the oracle is pinned
               against cv2.connectedComponentsWithStats after canonicalisation.
"""

"""Regenerates tests/golden/flo_ref.* and vis_ref.npz from the REFERENCE ITSELF, run in the build container
(/root/reference does not exist on the GPU box):
    python tests/golden/make_golden_vis_flo.py

* flo_ref.flo / flo_ref.npz : a flow field written by the reference's utils.write_flow and read back by its
                              utils.read_flow (/root/reference/src/utils.py:204-257) — the on-disk format of
                              the Dataset.get_flow_uv seam (/root/reference/src/datasets/dataset.py:205-212).
* vis_ref.npz               : the reference's im_helpers.to_rgb / apply_colormap (/root/reference/src/im_helpers.py:
                              103-135,162-201) applied as at processor.py:324-325,364-376, and the mask overlay of
                              processor.py:385-392 evaluated verbatim (cv2.addWeighted), on small synthetic inputs.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402


def main():
    import cv2
    detector, focus_of_expansion, im_helpers, utils = import_reference()
    rng = np.random.default_rng(41)

    # ---- .flo ----
    h, w = 23, 37
    flow = rng.normal(0, 3, (h, w, 2)).astype(np.float32)
    flow[0, 0] = (1e-9, -1e9)
    path = os.path.join(HERE, 'flo_ref.flo')
    utils.write_flow(path, flow)
    back = utils.read_flow(path)
    assert back.dtype == np.float32 and np.array_equal(back, flow)
    np.savez_compressed(os.path.join(HERE, 'flo_ref.npz'), flow=flow, read_back=back)

    # ---- visualisation ----
    H, W = 48, 64
    phi64 = rng.uniform(0, 180, (H, W))
    phi64[0, :8] = [0.0, 180.0, 90.0, 0.35294117647058826, 0.3529411764705882, 179.9, 1e-12, 45.0]
    phi32 = phi64.astype(np.float32)
    out = {}
    for name, phi in (('f64', phi64), ('f32', phi32)):
        rgb = im_helpers.to_rgb(phi, max_value=180.0)                       # processor.py:324
        out['phi_rgb_' + name] = rgb
        out['phi_jet_' + name] = im_helpers.apply_colormap(rgb.copy())      # processor.py:376
        out['phi_jet_max_' + name] = im_helpers.apply_colormap(rgb.copy(), max_value=180.0)   # processor.py:325
    frame = rng.integers(0, 256, (H, W, 3)).astype(np.uint8)
    frame[1, :6] = [[0, 0, 0], [255, 255, 255], [2, 3, 7], [12, 13, 17], [250, 1, 128], [128, 128, 128]]
    fixed = rng.random((H, W)) < 0.2
    fixed[1, :6] = True
    # processor.py:364 and :385-392 verbatim
    result_img = im_helpers.to_rgb(255 * fixed)
    mask_rgb = np.copy(frame)
    mask_rgb[fixed, 0] = 150
    mask_rgb[fixed, 1] = 0
    mask_rgb[fixed, 2] = 150
    alpha = 0.2
    mask_vis = cv2.addWeighted(frame, alpha, mask_rgb, 1.0 - alpha, 0.0)
    empty_img = im_helpers.to_rgb(255 * np.zeros((H, W), bool))
    lut = cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(1, 256), cv2.COLORMAP_JET).reshape(256, 3)
    np.savez_compressed(os.path.join(HERE, 'vis_ref.npz'), phi64=phi64, phi32=phi32, frame=frame, fixed=fixed,
                        result_img=result_img, empty_img=empty_img, mask_vis=mask_vis, jet_lut=lut, **out)
    print('written; jet lut head', lut[:3].tolist(), 'cv2', cv2.__version__)


if __name__ == '__main__':
    main()

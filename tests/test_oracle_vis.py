"""CPU: the visualisation restatement (oracle/vis_np.py) against the reference's own statement sequence run
with cv2 (farneback.py:83-99)."""
import numpy as np
import pytest

cv2 = pytest.importorskip('cv2')


def _flows():
    rng = np.random.default_rng(4)
    ys, xs = np.mgrid[0:120, 0:160].astype(np.float32)
    radial = np.stack([(xs - 60) * 0.05, (ys - 50) * 0.05], -1).astype(np.float32)
    noisy = (radial + rng.normal(0, 0.3, radial.shape)).astype(np.float32)
    axes = np.zeros((16, 64, 2), np.float32)     # widths are multiples of 64: see the note in oracle/vis_np.py
    axes[0, :8, 0] = np.linspace(-3, 3, 8)
    axes[1, :8, 1] = np.linspace(-3, 3, 8)
    axes[2:, :, :] = rng.normal(0, 2, (14, 64, 2))
    return {'radial': radial, 'noisy': noisy, 'axes': axes, 'zero': np.zeros((20, 64, 2), np.float32),
            'const': np.full((20, 64, 2), 1.5, np.float32)}


@pytest.mark.parametrize('name', ['radial', 'noisy', 'axes', 'zero', 'const'])
def test_restatement_equals_cv2_sequence(name):
    from oracle import vis_np
    flow = _flows()[name]
    ref, ref_invalid = vis_np.process_visualisation_cv2(flow, flow.shape[:2] + (3,))
    got, invalid = vis_np.process_visualisation(flow)
    assert invalid == ref_invalid
    assert got.shape == ref.shape and got.dtype == ref.dtype
    assert np.array_equal(got, ref), int((got != ref).any(-1).sum())


def test_hsv2bgr_exhaustive_s255():
    from oracle import vis_np
    H, V = np.meshgrid(np.arange(256), np.arange(256), indexing='ij')
    hsv = np.stack([H, np.full_like(H, 255), V], -1).astype(np.uint8)
    assert np.array_equal(vis_np.hsv2bgr_s255(hsv[..., 0], hsv[..., 2]), cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR))

"""GPU crop + resize (hfr_crop_resize_u8) must equal cv2.resize on the same crops bit for bit - the reference's own
per-face path (facial_analysis.py:267 + :95) - and feed the network the same pixels."""
import numpy as np
import pytest
import torch

import hse_facerec_tf_b200 as hfr

cv2 = pytest.importorskip("cv2")
pytestmark = pytest.mark.gpu


def test_crop_resize_bit_exact_vs_cv2():
    rs = np.random.RandomState(1)
    frames = rs.randint(0, 256, (3, 588, 784, 3)).astype(np.uint8)
    boxes, fidx = [], []
    for _ in range(64):
        x1, y1 = rs.randint(0, 700), rs.randint(0, 500)
        boxes.append([x1, y1, min(784, x1 + rs.randint(2, 300)), min(588, y1 + rs.randint(2, 300))])
        fidx.append(rs.randint(0, 3))
    boxes += [[0, 0, 784, 588], [783, 587, 784, 588], [100, 100, 324, 324]]   # whole frame, 1x1 crop, same size
    fidx += [0, 1, 2]
    for size in (224, 192):
        out = hfr.crop_resize(frames, boxes, size, frame_index=fidx).cpu().numpy()
        for i, (b, f) in enumerate(zip(boxes, fidx)):
            ref = cv2.resize(frames[f][b[1]:b[3], b[0]:b[2]], (size, size))
            assert np.array_equal(out[i], ref), (i, b, np.abs(out[i].astype(int) - ref.astype(int)).max())
    with pytest.raises(ValueError):
        hfr.crop_resize(frames, [[10, 10, 10, 20]], 224)
    with pytest.raises(ValueError):
        hfr.crop_resize(frames, [[0, 0, 800, 20]], 224)
    assert hfr.crop_resize(frames, np.zeros((0, 4)), 224).shape == (0, 224, 224, 3)


def test_process_boxes_equals_per_face_reference_path(age_gender_pb, golden_dir):
    """FacialImageProcessing.process_boxes (one GPU batch) == the reference's per-face loop through age_gender_fun."""
    rs = np.random.RandomState(2)
    crops = np.load(f"{golden_dir}/face_crops_u8.npz")["c224"]
    frame = rs.randint(0, 256, (588, 784, 3)).astype(np.uint8)
    frame[50:274, 60:284] = crops[0]
    frame[300:524, 400:624] = crops[2]
    dets = [[70, 60, 274, 264, 0.99], [410, 310, 614, 514, 0.98], [5, 5, 5, 40, 0.3]]
    fp = hfr.FacialImageProcessing(age_gender_pb, precision="tf32")
    bboxes, ages, genders, feats = fp.process_boxes(frame, dets)
    assert bboxes == [[60, 50, 284, 274], [400, 300, 624, 524]] and len(ages) == 2
    for (x1, y1, x2, y2), a, g, f in zip(bboxes, ages, genders, feats):
        ra, rg, rf = fp.age_gender_fun(frame[y1:y2, x1:x2, :])        # cv2.resize on the host, as the reference does
        assert abs(a - ra) < 1e-4 and abs(g[0] - rg[0]) < 1e-6
        np.testing.assert_allclose(f, rf, rtol=0, atol=1e-6)

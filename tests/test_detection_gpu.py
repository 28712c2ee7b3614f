"""MTCNN detector (scope row 8f-4) on the GPU against the oracle (oracle/mtcnn.py evaluating the same mtcnn.pb on the CPU):
the three networks one by one, then the whole cascade on the reference's test image."""
import cv2
import numpy as np
import pytest

import hse_facerec_tf_b200 as hfr
from oracle.mtcnn import MtcnnOracle, detect_faces

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nets(golden_dir):
    return hfr.MTCNN(f"{golden_dir}/mtcnn.pb"), MtcnnOracle(f"{golden_dir}/mtcnn.pb")


def _image(golden_dir):
    return cv2.cvtColor(cv2.imread(f"{golden_dir}/test_image.jpg"), cv2.COLOR_BGR2RGB)


@pytest.mark.parametrize("hw", [(12, 12), (13, 17), (100, 75), (417, 313)])
def test_pnet_matches_the_oracle(nets, golden_dir, hw):
    """P-Net is fully convolutional: odd and even sizes exercise the SAME-padded 2x2 pool (ceil) and the VALID convs."""
    gpu, ora = nets
    img = _image(golden_dir)
    x = (cv2.resize(img, (hw[1], hw[0]), interpolation=cv2.INTER_AREA) - 127.5) * 0.0078125
    x = np.transpose(x[None], (0, 2, 1, 3)).astype(np.float32)
    got, want = gpu.pnet(x), ora.pnet(x)
    assert len(got) == 2
    for g, w in zip(got, want):
        assert g.shape == w.shape
        np.testing.assert_allclose(g, w, rtol=0, atol=2e-5)


def test_rnet_onet_match_the_oracle(nets, golden_dir):
    gpu, ora = nets
    img = _image(golden_dir)
    rs = np.random.RandomState(0)
    for size, fun_g, fun_o, nout in ((24, gpu.rnet, ora.rnet, 2), (48, gpu.onet, ora.onet, 3)):
        crops = []
        for _ in range(37):
            y, x = rs.randint(0, img.shape[0] - 90), rs.randint(0, img.shape[1] - 90)
            s = rs.randint(20, 90)
            crops.append(cv2.resize(img[y:y + s, x:x + s], (size, size), interpolation=cv2.INTER_AREA))
        xb = np.transpose((np.stack(crops) - 127.5) * 0.0078125, (0, 2, 1, 3)).astype(np.float32)
        got, want = fun_g(xb), fun_o(xb)
        assert len(got) == nout
        for g, w in zip(got, want):
            assert g.shape == w.shape
            np.testing.assert_allclose(g, w, rtol=0, atol=5e-5)
    with pytest.raises(ValueError):
        gpu.rnet(np.zeros((2, 20, 20, 3), np.float32))         # R-Net's dense layer needs a 24 x 24 crop


def test_cascade_matches_the_oracle_and_feeds_process_image(nets, golden_dir, age_gender_pb):
    gpu, ora = nets
    img = _image(golden_dir)
    for view, minsize in ((img, 32), (img[100:400, 200:700], 20)):
        b0, p0 = detect_faces(ora, view, minsize)
        b1, p1 = gpu.detect_faces(view, minsize)
        assert b1.shape == b0.shape and len(b1) >= 1            # same faces (fp32 GPU vs fp32 CPU: thresholds not straddled)
        np.testing.assert_allclose(b1, b0, rtol=0, atol=0.05)   # pixels
        np.testing.assert_allclose(p1, p0, rtol=0, atol=0.05)
    assert len(gpu.detect_faces(img, 32)[0]) == 4               # the notebook's four faces
    # the detector as the upstream of the age/gender path: the reference's process_image flow end to end
    fp = hfr.FacialImageProcessing(False, True, 32, model_file=age_gender_pb, precision="tf32", detector=gpu)
    bboxes, points, ages, genders, feats = fp.process_image(np.ascontiguousarray(img[..., ::-1]))
    assert len(bboxes) == len(ages) == len(genders) == len(feats) == 4 and points.shape == (10, 4)
    assert all(1.0 <= a <= 100.0 for a in ages) and all(f.shape == (1024,) for f in feats)

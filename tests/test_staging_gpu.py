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
    fp = hfr.FacialImageProcessing(model_file=age_gender_pb, precision="tf32")
    bboxes, ages, genders, feats = fp.process_boxes(frame, dets)
    assert bboxes == [[60, 50, 284, 274], [400, 300, 624, 524]] and len(ages) == 2
    for (x1, y1, x2, y2), a, g, f in zip(bboxes, ages, genders, feats):
        ra, rg, rf = fp.age_gender_fun(frame[y1:y2, x1:x2, :])        # cv2.resize on the host, as the reference does
        assert abs(a - ra) < 1e-4 and abs(g[0] - rg[0]) < 1e-6
        np.testing.assert_allclose(f, rf, rtol=0, atol=1e-6)


def test_resize_pil_bit_exact_vs_pillow():
    """hfr_resize_pil_u8 == Image.resize(BILINEAR) (what scipy.misc.imresize called, facerec_test.py:84,93) on images of
    mixed sizes in one call: strong and mild reductions, enlargements, single-axis passes, identity, and crops."""
    from PIL import Image
    rs = np.random.RandomState(4)
    shapes = [(250, 250), (128, 128), (377, 512), (1000, 700), (60, 45), (33, 1200), (5, 7), (192, 300), (300, 192),
              (192, 192), (2000, 1500), (224, 224)]
    imgs = [rs.randint(0, 256, (h, w, 3)).astype(np.uint8) for h, w in shapes]
    for oh, ow in ((192, 192), (224, 224), (250, 250), (96, 160)):
        out = hfr.resize_pil(imgs, (oh, ow)).cpu().numpy()
        assert out.shape == (len(imgs), oh, ow, 3)
        for i, img in enumerate(imgs):
            ref = np.asarray(Image.fromarray(img).resize((ow, oh), resample=Image.BILINEAR))
            assert np.array_equal(out[i], ref), (shapes[i], oh, ow, np.abs(out[i].astype(int) - ref.astype(int)).max())
    # batched tensor input + crop window (the reference's img[dh:-dh, dw:-dw] between its two resizes)
    stack = rs.randint(0, 256, (5, 250, 250, 3)).astype(np.uint8)
    out = hfr.resize_pil(torch.from_numpy(stack).cuda(), 192, crop=(61, 61, 128, 128)).cpu().numpy()
    for i in range(5):
        ref = np.asarray(Image.fromarray(np.ascontiguousarray(stack[i, 61:-61, 61:-61])).resize((192, 192), Image.BILINEAR))
        assert np.array_equal(out[i], ref)
    assert hfr.resize_pil([], 192).shape == (0, 192, 192, 3)
    with pytest.raises(ValueError):
        hfr.resize_pil([np.zeros((10, 10), np.uint8)], 192)
    with pytest.raises(ValueError):
        hfr.resize_pil(stack, 192, crop=(200, 200, 128, 128))
    with pytest.raises(ValueError):                      # 64x reduction: outside the supported range, said loudly
        hfr.resize_pil([np.zeros((8192, 64, 3), np.uint8)], (128, 64))


def test_extract_files_equals_per_file_reference_path(age_gender_pb, tmp_path):
    """TensorFlowInference.extract_files (GPU resize, batched) == [extract_features(f) for f in files] (host Pillow
    resize, batch 1) - with and without the crop_center chain of facerec_test.py:81-89."""
    from PIL import Image
    rs = np.random.RandomState(6)
    paths = []
    for i, (h, w) in enumerate([(250, 250), (300, 200), (181, 233), (640, 480), (97, 131)]):
        p = tmp_path / f"img{i}.png"
        Image.fromarray(rs.randint(0, 256, (h, w, 3)).astype(np.uint8)).save(p)
        paths.append(str(p))
    tfi = hfr.TensorFlowInference(age_gender_pb, "input_1:0", "global_pooling/Mean:0", precision="tf32", input_hw=192)
    for crop_center in (False, True):
        want = np.stack([tfi.extract_features(p, crop_center=crop_center) for p in paths])
        got = tfi.extract_files(paths, crop_center=crop_center, batch=3)
        np.testing.assert_array_equal(got, want)

"""Host logic and oracle of the input-staging row: box expansion as process_image does it, and the numpy restatement of
cv2's uint8 bilinear resize pinned against cv2.resize itself."""
import numpy as np
import pytest

from hse_facerec_tf_b200.staging import expand_and_clamp_boxes
from oracle.resize import resize_linear_u8

cv2 = pytest.importorskip("cv2")


def test_resize_restatement_is_bit_exact_vs_cv2():
    rs = np.random.RandomState(0)
    for _ in range(25):
        h, w = rs.randint(5, 400), rs.randint(5, 400)
        img = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
        for ow, oh in ((224, 224), (192, 192), (160, 96)):
            assert np.array_equal(resize_linear_u8(img, ow, oh), cv2.resize(img, (ow, oh)))
    same = rs.randint(0, 256, (224, 224, 3)).astype(np.uint8)
    assert np.array_equal(resize_linear_u8(same, 224, 224), same)


def test_box_expansion_matches_process_image():
    # facial_analysis.py:236-263 on a 588x784 frame (height 588, width 784 - the size of the reference's test image)
    boxes = [[100.7, 50.2, 180.9, 140.0, 0.99], [2, 3, 40, 60, 0.9], [700, 500, 790, 600, 0.8], [10, 10, 10, 50, 0.5]]
    got = expand_and_clamp_boxes(boxes, img_h=588, img_w=784)
    assert got == [[90, 40, 190, 150], [0, 0, 50, 70], [690, 490, 784, 588]]   # degenerate 4th box dropped


def test_pil_bilinear_restatement_is_bit_exact_vs_pillow():
    """scipy.misc.imresize(img, size, interp='bilinear') (facerec_test.py:84,93) is Pillow's BILINEAR resample; the
    oracle restatement must equal Pillow itself for shrinking (antialiased), enlarging, one-axis-only and identity."""
    from PIL import Image
    from oracle.resize import pil_resize_bilinear_u8
    rs = np.random.RandomState(3)
    shapes = [(250, 250), (128, 128), (377, 512), (640, 480), (60, 45), (33, 900), (5, 7), (192, 300), (300, 192)]
    for h, w in shapes:
        img = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
        for oh, ow in ((192, 192), (224, 224), (250, 250)):
            ref = np.asarray(Image.fromarray(img).resize((ow, oh), resample=Image.BILINEAR))
            assert np.array_equal(pil_resize_bilinear_u8(img, oh, ow), ref), (h, w, oh, ow)
    # the reference's crop_center chain: 250x250, centre 128x128, network size
    img = rs.randint(0, 256, (301, 411, 3)).astype(np.uint8)
    big = pil_resize_bilinear_u8(img, 250, 250)
    want = np.asarray(Image.fromarray(np.asarray(Image.fromarray(img).resize((250, 250), Image.BILINEAR))[61:-61, 61:-61])
                      .resize((192, 192), Image.BILINEAR))
    assert np.array_equal(pil_resize_bilinear_u8(big[61:-61, 61:-61], 192, 192), want)

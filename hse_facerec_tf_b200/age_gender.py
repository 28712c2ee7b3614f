"""Drop-in for the age/gender path of age_gender_identity/facial_analysis.py.

  FacialImageProcessing(print_stat, mtcnn_detector, minsize)  facial_analysis.py:35-71   (same positional arguments)
  FacialImageProcessing.load_age_gender -> age_gender_fun     facial_analysis.py:83-130
  FacialImageProcessing.is_male                               facial_analysis.py:76-81
  FacialImageProcessing.process_image                         facial_analysis.py:225-294 (face loop as one GPU batch)

Detection (MTCNN / LBP cascade, facial_analysis.py:210-223, 334-604) is upstream of the hot path: `detect_faces` calls
the `detector` the object was given (any callable img_rgb -> (boxes, points)), or boxes can be passed to process_image.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, lib
from .model import HfrModel, _stream_ptr

AGE_GENDER_OUTPUTS = ["age_pred/Softmax:0", "gender_pred/Sigmoid:0", "global_pooling/Mean:0"]


# the model file names the reference hard-codes next to facial_analysis.py (facial_analysis.py:45) and ships (README)
DEFAULT_MODEL_FILES = ("age_gender_tf2_224_deep-03-0.13-0.97_new.pb", "age_gender_tf2_new-01-0.14-0.92.pb",
                       "age_gender_tf2_new-01-0.14-0.92_quantized.pb")


def _default_model_file():
    import os
    env = os.environ.get("HFR_AGE_GENDER_MODEL")
    if env:
        return env
    for d in (os.getcwd(), os.path.join(os.getcwd(), "age_gender_identity")):
        for name in DEFAULT_MODEL_FILES:
            if os.path.exists(os.path.join(d, name)):
                return os.path.join(d, name)
    raise FileNotFoundError("no age/gender model found: pass model_file=..., set HFR_AGE_GENDER_MODEL, or run next to one "
                            "of " + ", ".join(DEFAULT_MODEL_FILES))


class FacialImageProcessing:
    def __init__(self, print_stat=False, mtcnn_detector=True, minsize=32, *, model_file=None, detector=None,
                 device="cuda:0", precision="bf16"):
        """Positional arguments as in the reference (facial_analysis.py:37): FacialImageProcessing(True),
        FacialImageProcessing(print_stat=False, minsize=112) keep working (facial_analysis.py:609,
        process_photos.py:385, utkface_test.py:24).  model_file: the frozen age/gender graph (default: the reference's own
        file names, looked up in the working directory).  detector: callable img_rgb -> (bounding_boxes, points)."""
        self.mtcnn_detector = mtcnn_detector
        self.print_stat = print_stat
        self.minsize = minsize
        self.detector = detector
        self.model = HfrModel(model_file or _default_model_file(), "input_1:0", AGE_GENDER_OUTPUTS, device=device,
                              precision=precision)
        self.age_gender_fun = self.load_age_gender()

    def detect_faces(self, img):
        """facial_analysis.py:210-223.  The detector is upstream of the hot path: it is whatever callable was passed as
        `detector` (e.g. the MTCNN of this package's `detection` module when present, or OpenCV's cascade)."""
        if self.detector is None:
            raise NotImplementedError("no face detector configured: pass detector=callable(img_rgb)->(boxes, points), "
                                      "or give the boxes to process_image(draw, bounding_boxes=...)")
        return self.detector(img)

    def process_image(self, draw, bounding_boxes=None, points=None):
        """facial_analysis.py:225-294: BGR frame in (cv2 convention), (bboxes, points, ages, genders, facial_features)
        out - the per-face loop runs as ONE batch on the GPU (process_boxes).  bounding_boxes given: skip detection."""
        img = np.ascontiguousarray(np.asarray(draw)[..., ::-1])         # cv2.cvtColor(draw, cv2.COLOR_BGR2RGB)
        if bounding_boxes is None:
            bounding_boxes, points = self.detect_faces(img)
        bboxes, ages, genders, feats = self.process_boxes(img, bounding_boxes)
        return bboxes, ([] if points is None else points), ages, genders, feats

    def close(self):
        self.model.close()

    @staticmethod
    def is_male(gender_preds):
        return gender_preds >= 0.6

    # -- batched -------------------------------------------------------------------------------------
    def age_gender_batch(self, x, graph=False):
        """x: [B,H,W,3] RGB uint8 crops already resized to the network size (CUDA tensor or numpy).
        Returns (age [B], gender [B,1], feat [B,1024], age_probs [B,100]); tensors for tensor input, numpy otherwise."""
        as_numpy = not isinstance(x, torch.Tensor)
        if as_numpy:
            x = torch.from_numpy(np.ascontiguousarray(x)).to(self.model.device)
        probs, gender, feat = self.model.forward(x, True, True, graph=graph)
        age = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
        dev = x.device.index or 0
        check(lib.hfr_age_gender_post(probs.data_ptr(), x.shape[0], probs.shape[1], age.data_ptr(), dev,
                                      _stream_ptr(x.device)))
        if as_numpy:
            return age.cpu().numpy(), gender.cpu().numpy(), feat.cpu().numpy(), probs.cpu().numpy()
        return age, gender, feat, probs

    def process_boxes(self, img_rgb, bounding_boxes, graph=False):
        """The face loop of process_image (facial_analysis.py:233-293) as ONE batch on the GPU: expand every detected box
        by 10 px, clamp, crop, resize (cv2-exact) and run the network.  Returns (bboxes, ages, genders, features) with the
        reference's list layout; detection itself (MTCNN) is upstream and not part of this class."""
        from .staging import crop_resize, expand_and_clamp_boxes
        img_h, img_w, _ = img_rgb.shape
        bboxes = expand_and_clamp_boxes(bounding_boxes, img_h, img_w)
        if not bboxes:
            return [], [], [], []
        crops = crop_resize(img_rgb, bboxes, (self.model.h, self.model.w))
        age, gender, feat, _ = self.age_gender_batch(crops, graph=graph)
        age, gender, feat = age.cpu().numpy(), gender.cpu().numpy(), feat.cpu().numpy()
        return bboxes, [float(a) for a in age], [g for g in gender], [f for f in feat]

    # -- reference-compatible closure -------------------------------------------------------------------
    def load_age_gender(self, sess=None, graph=None):
        w, h = self.model.w, self.model.h

        def age_gender_fun(img):
            import cv2
            resized_image = cv2.resize(img, (w, h))
            age, gender, feat, probs = self.age_gender_batch(np.ascontiguousarray(resized_image)[None])
            if self.print_stat:
                idx = probs[0].argsort()[::-1][:2]
                print('gender', gender[0])
                print('age', age[0])
                print(idx, probs[0][idx], probs[0][idx] / probs[0][idx].sum())
            return float(age[0]), gender[0], feat[0]
        return age_gender_fun

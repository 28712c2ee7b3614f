"""Drop-in extractor entry points of the reference, backed by libhfr.so instead of TensorFlow / Keras.

  TensorFlowInference      facerec_test.py:50-125    (same constructor arguments, attributes and methods)
  extract_keras_features   facerec_test.py:128-147
  FeatureExtractor         age_gender_identity/facial_clustering_test.py:288-319

plus the batched entry `extract_batch` the throughput metric is quoted on.  Image decoding / resizing stays on the host
(it is upstream of the "image batch in" boundary); colour flip and mean subtraction run fused in the first GPU kernel.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .model import HfrModel


def _imread_rgb(path) -> np.ndarray:
    from PIL import Image
    with Image.open(path) as im:
        return np.asarray(im.convert("RGB"))


def _imresize_bilinear(img: np.ndarray, size) -> np.ndarray:
    """scipy.misc.imresize(img, size, interp='bilinear') (removed from SciPy >= 1.3): PIL resize, uint8 result."""
    from PIL import Image
    return np.asarray(Image.fromarray(img).resize((int(size[1]), int(size[0])), resample=Image.BILINEAR))


class TensorFlowInference:
    def __init__(self, frozen_graph_filename, input_tensor, output_tensor, learning_phase_tensor=None, convert2BGR=True,
                 imageNetUtilsMean=True, additional_input_value=0, *, device="cuda:0", precision="bf16", input_hw=0):
        self.model = HfrModel(frozen_graph_filename, input_tensor, [output_tensor],
                              learning_phase_tensor=learning_phase_tensor,
                              additional_input_value=additional_input_value, input_hw=input_hw, device=device,
                              precision=precision)
        self.w, self.h = self.model.w, self.model.h
        self.convert2BGR = convert2BGR
        self.imageNetUtilsMean = imageNetUtilsMean
        self.additional_input_value = additional_input_value

    # -- reference-compatible per-image API -------------------------------------------------------
    def _load_resized_u8(self, img_filepath, crop_center):
        img = _imread_rgb(img_filepath)
        if crop_center:                                   # facerec_test.py:81-89
            orig_w, orig_h = 250, 250
            img = _imresize_bilinear(img, (orig_w, orig_h))
            w1, h1 = 128, 128
            dw, dh = (orig_w - w1) // 2, (orig_h - h1) // 2
            img = img[dh:-dh, dw:-dw]
        # (rows, cols) = (h, w) here and in extract_files.  The reference passes (w, h) to imresize (facerec_test.py:93),
        # which only works - and is identical - for the square placeholders all of its networks have
        return _imresize_bilinear(img, (self.h, self.w))

    def preprocess_image(self, img_filepath, crop_center):
        """Host restatement of facerec_test.py:80-112 (returns the float array the reference would feed)."""
        x = self._load_resized_u8(img_filepath, crop_center).astype(float)
        if self.convert2BGR:
            x = x[..., ::-1]
            if self.imageNetUtilsMean:
                x[..., 0] -= 103.939
                x[..., 1] -= 116.779
                x[..., 2] -= 123.68
            else:
                x[..., 0] -= 91.4953
                x[..., 1] -= 103.8827
                x[..., 2] -= 131.0912
        else:
            x /= 127.5
            x -= 1.
        return x

    def extract_features(self, img_filepath, crop_center=False):
        u8 = self._load_resized_u8(img_filepath, crop_center)
        (out,) = self.model.forward_host(u8[None], self.convert2BGR, self.imageNetUtilsMean)
        return out.reshape(-1)

    def close_session(self):
        self.model.close()

    # -- batched API ---------------------------------------------------------------------------------
    def extract_batch(self, x, l2norm=False, graph=False):
        """x: [B,H,W,3] RGB uint8 crops (or float32 already pre-processed); torch CUDA tensor -> torch CUDA tensor
        [B,D]; numpy array -> numpy array (copies included)."""
        if isinstance(x, torch.Tensor):
            (out,) = self.model.forward(x, self.convert2BGR, self.imageNetUtilsMean, l2norm=l2norm, graph=graph)
            return out
        (out,) = self.model.forward_host(np.asarray(x), self.convert2BGR, self.imageNetUtilsMean, l2norm=l2norm,
                                         graph=graph)
        return out


    def extract_files(self, img_filepaths, crop_center=False, l2norm=False, batch=64):
        """[self.extract_features(f, crop_center) for f in img_filepaths] (facerec_test.py:394) in batches: files are
        decoded on the host, resized (Pillow-exact) and embedded on the GPU.  Returns float32 [n, D]."""
        from .staging import load_resized_batch
        outs = []
        for i in range(0, len(img_filepaths), batch):
            x = load_resized_batch(img_filepaths[i:i + batch], (self.h, self.w), crop_center, self.model.device)
            outs.append(self.extract_batch(x, l2norm=l2norm, graph=False).cpu().numpy())
        d = self.model.out_dims[0]
        return np.concatenate(outs) if outs else np.empty((0, d), np.float32)

    def extract_stream(self, batches, l2norm=False, depth=2):
        """The dataset loop of facerec_test.py:394 with batches instead of single files: an iterable of host batches
        ([B,H,W,3] uint8 RGB crops, or float32 already pre-processed) -> a generator of [B,D] float32 arrays, `depth`
        batches in flight so that uploads and downloads overlap the compute.  A yielded array is valid only until the
        next item is requested (its pinned buffer is resubmitted then) - copy it (np.vstack / .copy()) to keep it."""
        for (out,) in self.model.stream_host(batches, depth=depth, convert2BGR=self.convert2BGR,
                                             imageNetUtilsMean=self.imageNetUtilsMean, l2norm=l2norm, graph=True):
            yield out


def extract_keras_features(model, img_filepath, crop_center):
    """facerec_test.py:128-147 with `model` a TensorFlowInference built from the Keras .h5/.pb (caffe-mode
    preprocess_input == convert2BGR + ImageNet mean).  Keras' load_img(target_size=...) resizes with PIL NEAREST
    (its default interpolation='nearest'); the bare img.resize((w, h)) of the crop_center branch used Pillow's default
    filter, which was NEAREST in the Pillow releases of the reference's era (it became BICUBIC in Pillow 7): the filter
    is passed explicitly so the result does not depend on the installed Pillow."""
    from PIL import Image
    w, h = model.w, model.h
    with Image.open(img_filepath) as im:
        im = im.convert("RGB")
        if crop_center:
            orig_w, orig_h = 250, 250
            im = im.resize((orig_w, orig_h), Image.NEAREST)
            w1, h1 = 128, 128
            dw, dh = (orig_w - w1) / 2, (orig_h - h1) / 2
            im = im.crop((dw, dh, orig_w - dw, orig_h - dh)).resize((w, h), Image.NEAREST)
        else:
            im = im.resize((w, h), Image.NEAREST)
        u8 = np.asarray(im)
    (out,) = model.model.forward_host(u8[None], True, True)
    return out.reshape(-1)


class FeatureExtractor:
    """facial_clustering_test.py:288-319.  vggmodel=None: the age/gender MobileNet used as an embedder
    (`global_pooling/Mean:0`, facial_clustering_test.py:291); vggmodel='resnet50': the VGGFace2 ResNet-50, which the
    reference takes from keras_vggface and this build loads from its frozen twin `models/vgg2_resnet.pb`
    (facerec_test.py:213: BGR + VGGFace2 mean).  'vgg16' is a competitor baseline outside the hot path."""

    def __init__(self, vggmodel=None, model_path=None, **kw):
        if vggmodel is None:
            self.tfInference = TensorFlowInference(model_path or "age_gender_tf2_new-01-0.14-0.92.pb",
                                                   input_tensor="input_1:0", output_tensor="global_pooling/Mean:0", **kw)
        elif vggmodel == "resnet50":
            self.tfInference = TensorFlowInference(model_path or os.path.join("models", "vgg2_resnet.pb"),
                                                   input_tensor="input:0", output_tensor="pool5_7x7_s1:0",
                                                   convert2BGR=True, imageNetUtilsMean=False, **kw)
        else:
            raise ValueError(f"vggmodel={vggmodel!r}: only None (age/gender MobileNet) and 'resnet50' are on the hot path")

    def extract_features(self, image_path):
        return self.tfInference.extract_features(image_path)

    def close(self):
        if self.tfInference is not None:
            self.tfInference.close_session()

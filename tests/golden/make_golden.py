"""Generates the committed fixtures under tests/golden/ from the read-only reference mount.

Run once in the build container (needs /root/reference, cv2):   python tests/golden/make_golden.py

Outputs
  age_gender_quantized.pb   byte copy of the reference's only shipped model graph
                            (age_gender_identity/age_gender_tf2_new-01-0.14-0.92_quantized.pb, Apache-2.0,
                            sha256 2a917c62...dee3) - a weight artefact, not source code.  The GPU box has no
                            /root/reference, so the model the library is a drop-in for must travel with the tests.
  face_crops_u8.npz         RGB uint8 crops of age_gender_identity/test_image.jpg resized with cv2.resize
                            (INTER_LINEAR, as facial_analysis.py:95) to 224 and 192, plus shifted/flipped variants.
  mtcnn.pb, test_image.jpg  byte copies of age_gender_identity/mtcnn.pb (the detector's weights, the file the MTCNN drop-in
                            exists to load) and of the demo photo the notebook runs on - inputs of the detector tests.
  kat_oracle.npz            oracle outputs (f32-dequant, fp32 compute) for zeros / seeded noise / crops at 224 and
                            192 - lets the GPU tests check against values produced in THIS container as well.
"""
import hashlib
import os
import shutil
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.tfnet import GraphOracle, preprocess_rgb_u8  # noqa: E402

REF = "/root/reference/age_gender_identity"
PB = os.path.join(REF, "age_gender_tf2_new-01-0.14-0.92_quantized.pb")
SHA = "2a917c62e3a6cdd2d1556a86ae564dacb3704243048659d5d4d217933239dee3"


def copy_detector_fixtures():
    for name in ("mtcnn.pb", "test_image.jpg"):
        shutil.copyfile(os.path.join(REF, name), os.path.join(HERE, name))


def crops(size):
    img = cv2.cvtColor(cv2.imread(os.path.join(REF, "test_image.jpg")), cv2.COLOR_BGR2RGB)
    boxes = [(203, 285, 581, 663), (195, 290, 570, 670), (210, 280, 590, 655), (150, 330, 520, 720)]
    out = []
    for (t, b, l, r) in boxes:
        c = cv2.resize(img[t:b, l:r], (size, size))
        out.append(c)
        out.append(c[:, ::-1].copy())
    # a few other regions of the photo (other faces / background texture)
    for (t, l) in [(60, 60), (120, 300), (300, 100), (380, 420)]:
        out.append(cv2.resize(img[t:t + 160, l:l + 160], (size, size)))
    return np.stack(out).astype(np.uint8)


def main():
    data = open(PB, "rb").read()
    assert hashlib.sha256(data).hexdigest() == SHA
    shutil.copyfile(PB, os.path.join(HERE, "age_gender_quantized.pb"))
    copy_detector_fixtures()
    c224, c192 = crops(224), crops(192)
    np.savez_compressed(os.path.join(HERE, "face_crops_u8.npz"), c224=c224, c192=c192)

    g = GraphOracle(PB)
    outs = ["age_pred/Softmax:0", "gender_pred/Sigmoid:0", "global_pooling/Mean:0"]
    kat = {}
    for size, cr in ((224, c224), (192, c192)):
        noise = np.random.RandomState(0).randint(0, 256, (1, size, size, 3)).astype(np.uint8)
        batch = np.concatenate([noise, cr[:4]])
        x = np.concatenate([np.zeros((1, size, size, 3), np.float32), preprocess_rgb_u8(batch)])
        if size == 224:
            a, ge, f = g.run(outs, {"input_1:0": x})
            kat["age_probs_224"], kat["gender_224"] = a, ge
        else:
            (f,) = g.run(outs[2:], {"input_1:0": x})
        kat[f"emb_{size}"] = f
    np.savez_compressed(os.path.join(HERE, "kat_oracle.npz"), **kat)
    for k, v in kat.items():
        print(k, v.shape, float(np.abs(v).sum()))


if __name__ == "__main__":
    main()

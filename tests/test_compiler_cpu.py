"""Graph compiler (C++ host code behind hfr_model_load, device=-1) checked without a GPU: the fused plan, executed on
the CPU with the folded weights the library exports, must reproduce the oracle's evaluation of the original graph."""
import ctypes
import os
import re

import numpy as np
import pytest

import hse_facerec_tf_b200 as hfr
from hse_facerec_tf_b200 import _lib
from oracle.tfnet import GraphOracle, preprocess_rgb_u8
from tests.helpers import cosine, run_plan_cpu

OUTS = ["age_pred/Softmax:0", "gender_pred/Sigmoid:0", "global_pooling/Mean:0"]


def test_abi_exports_every_declared_symbol():
    header = open(os.path.join(os.path.dirname(_lib.__file__), "..", "include", "hfr.h")).read()
    declared = set(re.findall(r"\b(hfr_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/hfr.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.hfr_version() >= 100


def test_plan_of_real_graph(age_gender_pb):
    m = hfr.HfrModel(age_gender_pb, "input_1:0", OUTS, device=None)
    assert (m.h, m.w, m.c) == (224, 224, 3) and m.out_dims == [100, 1, 1024]
    layers = m.plan()["layers"]
    kinds = [L["kind"] for L in layers]
    assert kinds == ["stem"] + ["dw", "pw"] * 13 + ["gap", "fc", "fc", "fc"]
    assert [L["stride"] for L in layers if L["kind"] == "dw"] == [1, 2, 1, 2, 1, 2, 1, 1, 1, 1, 1, 2, 1]
    assert all(L["act"] == "relu6" for L in layers[:27])
    assert [L["act"] for L in layers[27:]] == ["none", "relu", "softmax", "sigmoid"]
    # TF SAME at stride 2 on even sizes: 0 before, 1 after
    assert layers[0]["pad"] == [0, 1, 0, 1] and layers[3]["pad"] == [0, 1, 0, 1] and layers[1]["pad"] == [1, 1, 1, 1]
    assert [L["cout"] for L in layers if L["kind"] == "pw"] == [64, 128, 128, 256, 256] + [512] * 6 + [1024, 1024]


@pytest.mark.parametrize("size", [224, 192])
def test_compiled_plan_matches_oracle(age_gender_pb, golden_dir, size):
    crops = np.load(f"{golden_dir}/face_crops_u8.npz")[f"c{size}"][:2]
    noise = np.random.RandomState(0).randint(0, 256, (1, size, size, 3)).astype(np.uint8)
    x = preprocess_rgb_u8(np.concatenate([noise, crops]))
    outs = OUTS if size == 224 else OUTS[2:]
    m = hfr.HfrModel(age_gender_pb, "input_1:0", outs, device=None, input_hw=0 if size == 224 else size)
    got, _ = run_plan_cpu(m, x)
    ref = GraphOracle(age_gender_pb).run(outs, {"input_1:0": x})
    for g, r in zip(got, ref):
        np.testing.assert_allclose(g, r.reshape(g.shape), rtol=2e-3, atol=2e-5)
    assert cosine(got[-1], ref[-1].reshape(got[-1].shape)).min() > 0.999999


def test_error_mapping(age_gender_pb, tmp_path):
    with pytest.raises(KeyError):
        hfr.HfrModel(age_gender_pb, "input_1:0", ["no_such_tensor:0"], device=None)
    with pytest.raises(KeyError):
        hfr.HfrModel(age_gender_pb, "no_such_input:0", OUTS, device=None)
    with pytest.raises(FileNotFoundError):
        hfr.HfrModel(str(tmp_path / "missing.pb"), "input_1:0", OUTS, device=None)
    bad = tmp_path / "bad.pb"
    bad.write_bytes(b"\xff\xff\xff\xff not a graphdef")
    with pytest.raises(ValueError):
        hfr.HfrModel(str(bad), "input_1:0", OUTS, device=None)
    m = hfr.HfrModel(age_gender_pb, "input_1:0", OUTS, device=None)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(hfr.HfrError):   # no CPU fallback: a GPU handle cannot be created here
            hfr.HfrModel(age_gender_pb, "input_1:0", OUTS, device="cuda:0")


def test_unfolded_keras_mobilenet_with_learning_phase(tmp_path):
    """The format of the missing models/vgg2_mobilenet.pb (facerec_test.py:212): un-folded BN arithmetic under
    Switch/Merge on the learning-phase placeholder, Reshape((1,1,1024)) output."""
    from hse_facerec_tf_b200.synth import write_mobilenet_pb
    pb = write_mobilenet_pb(str(tmp_path / "vgg2_mobilenet.pb"), seed=3, input_hw=96)
    args = dict(learning_phase_tensor="conv1_bn/keras_learning_phase:0", device=None)
    m = hfr.HfrModel(pb, "input_1:0", ["reshape_1/Reshape:0"], **args)
    assert (m.h, m.w) == (96, 96) and m.out_dims == [1024]
    kinds = [L["kind"] for L in m.plan()["layers"]]
    assert kinds == ["stem"] + ["dw", "pw"] * 13 + ["gap"]
    x = preprocess_rgb_u8(np.random.RandomState(1).randint(0, 256, (2, 96, 96, 3)).astype(np.uint8))
    (ref,) = GraphOracle(pb).run(["reshape_1/Reshape:0"], {"input_1:0": x, "conv1_bn/keras_learning_phase:0": False})
    got, _ = run_plan_cpu(m, x)
    np.testing.assert_allclose(got[0], ref.reshape(2, -1), rtol=1e-3, atol=1e-4)
    with pytest.raises(ValueError):      # phase placeholder not named -> the conditional cannot be resolved
        hfr.HfrModel(pb, "input_1:0", ["reshape_1/Reshape:0"], device=None)
    # feeding the phase with 1 selects the (here: stand-in) training branch, as TF would
    m1 = hfr.HfrModel(pb, "input_1:0", ["reshape_1/Reshape:0"], additional_input_value=1, **args)
    got1, _ = run_plan_cpu(m1, x)
    assert np.abs(got1[0] - got[0]).max() > 1.0


def test_resnet50_caffe_style_plan(tmp_path):
    """The format assumed for the missing models/vgg2_resnet.pb (facerec_test.py:213): Pad+VALID stem, explicit Pad in
    front of the ceil-mode max pool, FusedBatchNorm, stride on the 1x1 reduce, residual Add -> Relu, 7x7 AvgPool."""
    from collections import Counter
    from hse_facerec_tf_b200.synth import write_resnet50_pb
    pb = write_resnet50_pb(str(tmp_path / "vgg2_resnet.pb"), seed=5)
    m = hfr.HfrModel(pb, "input:0", ["pool5_7x7_s1:0"], device=None)
    assert m.out_dims == [2048]
    layers = m.plan()["layers"]
    assert Counter(L["kind"] for L in layers) == {"pw": 36, "conv": 16, "subsample": 3, "stem": 1, "maxpool": 1, "gap": 1}
    assert layers[0]["pad"] == [3, 3, 3, 3] and layers[0]["k"] == [7, 7] and layers[0]["hw_out"] == [112, 112]
    assert layers[1]["kind"] == "maxpool" and layers[1]["explicit_zero_pad"] and layers[1]["hw_out"] == [56, 56]
    assert sum(1 for L in layers if L["in2"] >= 0) == 16          # one fused residual add per bottleneck
    x = preprocess_rgb_u8(np.random.RandomState(2).randint(0, 256, (1, 224, 224, 3)).astype(np.uint8), True, False)
    (ref,) = GraphOracle(pb).run(["pool5_7x7_s1:0"], {"input:0": x})
    got, _ = run_plan_cpu(m, x)
    np.testing.assert_allclose(got[0], ref.reshape(1, -1), rtol=2e-3, atol=2e-4)


@pytest.mark.parametrize("full_model", [True, False])
def test_keras_h5_weights_file(tmp_path, full_model):
    """models/vgg2_mobilenet.h5 (facerec_test.py:326-334): hand-written HDF5 reader (no libhdf5 / h5py in this image).
    The fixture is written by hse_facerec_tf_b200.synth.H5Writer in the layout Keras 2.x + h5py produce by default
    (superblock v0, v1 object headers, symbol-table groups, /model_weights/<layer>/<layer>/<weight>:0); the oracle
    evaluates a .pb twin built from the same arrays."""
    from hse_facerec_tf_b200.synth import mobilenet_weights, write_keras_mobilenet_h5, write_mobilenet_pb_from_weights
    w = mobilenet_weights(seed=11, heads=full_model)
    h5 = write_keras_mobilenet_h5(str(tmp_path / "vgg2_mobilenet.h5"), w, full_model=full_model)
    pb = write_mobilenet_pb_from_weights(str(tmp_path / "twin.pb"), w, input_hw=96)
    outs = ["reshape_1/Reshape:0"] + (["age_pred/Softmax:0", "gender_pred/Sigmoid:0"] if full_model else [])
    m = hfr.HfrModel(h5, None, outs, device=None, input_hw=96)
    assert (m.h, m.w) == (96, 96) and m.out_dims[0] == 1024
    x = preprocess_rgb_u8(np.random.RandomState(4).randint(0, 256, (2, 96, 96, 3)).astype(np.uint8))
    ref = GraphOracle(pb).run(outs, {"input_1:0": x})
    got, _ = run_plan_cpu(m, x)
    for g, r in zip(got, ref):
        np.testing.assert_allclose(g, r.reshape(g.shape), rtol=1e-3, atol=1e-5)
    # default size is the reference's sz=192 (facerec_test.py:325); default output is reshape_1
    m192 = hfr.HfrModel(h5, None, [], device=None)
    assert (m192.h, m192.w) == (192, 192) and m192.out_dims == [1024]


def test_keras_h5_errors(tmp_path):
    from hse_facerec_tf_b200.synth import H5Writer, mobilenet_weights, write_keras_mobilenet_h5
    h5 = write_keras_mobilenet_h5(str(tmp_path / "ok.h5"), mobilenet_weights(seed=2))
    data = open(h5, "rb").read()
    (tmp_path / "trunc.h5").write_bytes(data[: len(data) // 3])
    with pytest.raises(ValueError):
        hfr.HfrModel(str(tmp_path / "trunc.h5"), None, [], device=None)
    other = H5Writer()
    other.dataset("model_weights/dense_1/dense_1/kernel:0", np.zeros((4, 4), np.float32))
    other.save(str(tmp_path / "other.h5"))
    with pytest.raises(ValueError, match="conv1"):
        hfr.HfrModel(str(tmp_path / "other.h5"), None, [], device=None)
    with pytest.raises(KeyError):
        hfr.HfrModel(h5, None, ["age_pred/Softmax:0"], device=None)      # this file has no heads

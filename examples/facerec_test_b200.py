"""The flow of the reference's facerec_test.py (__main__, lines 296-442) with this package in place of TensorFlow and
scikit-learn's k-NN: extract features for a directory tree `DATASET/<identity>/<image>` -> .npz cache {x, y} -> L2
normalise -> the reference's classifier list through classifier_tester (StratifiedShuffleSplit + cross_validate).

    python examples/facerec_test_b200.py --dataset /data/lfw --model models/vgg2_mobilenet.pb \\
        --input input_1:0 --output reshape_1/Reshape:0 --phase conv1_bn/keras_learning_phase:0

Without --dataset a small synthetic tree is generated (random crops of the repo's golden faces with per-identity colour
shifts), and without --model the repo's only shipped network (the age/gender MobileNet, whose body is the VGGFace2
identity network) is used at 192x192 - enough to exercise every call on a B200 box.
Only the lines that touch the framework differ from the reference; they are marked `# <-`.
"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np
from sklearn import model_selection, preprocessing
from sklearn.pipeline import Pipeline

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hse_facerec_tf_b200 as hfr  # noqa: E402                                   # <- instead of tensorflow / sklearn.neighbors


def is_image(path):
    return path.lower().endswith((".jpg", ".jpeg", ".png", ".bmp"))


def get_files(db_dir):
    """[identity, identity/file] for every image one level below db_dir (the layout facerec_test.py:150-153 walks)."""
    pairs = []
    for identity in sorted(os.listdir(db_dir)):
        folder = os.path.join(db_dir, identity)
        if os.path.isdir(folder):
            pairs += [[identity, os.path.join(identity, f)] for f in sorted(os.listdir(folder)) if is_image(f)]
    return pairs


def classifier_tester(classifier, x, y):
    """Same protocol as facerec_test.py:200-207: one stratified 50/50 split (seed 0), accuracy through cross_validate -
    which clones the estimator, so it exercises get_params/set_params of the drop-in classes."""
    split = model_selection.StratifiedShuffleSplit(n_splits=1, test_size=0.5, random_state=0)
    result = model_selection.cross_validate(classifier, x, y, scoring="accuracy", cv=split)
    accuracy = 100.0 * result["test_score"]
    print(f"  accuracy {accuracy.mean():.2f} % (+- {accuracy.std():.2f}), predict time {result['score_time'].sum() * 1e3:.1f} ms")


def synthetic_dataset(root, n_ids=24, per_id=6):
    from PIL import Image
    crops = np.load(os.path.join(ROOT, "tests", "golden", "face_crops_u8.npz"))["c224"]
    rs = np.random.RandomState(0)
    for i in range(n_ids):
        d = os.path.join(root, f"id{i:03d}")
        os.makedirs(d)
        base = crops[i % len(crops)].astype(np.int32) + rs.randint(-60, 60, (1, 1, 3))
        for j in range(per_id):
            y0, x0 = rs.randint(0, 24, 2)
            img = np.clip(base[y0:y0 + 200, x0:x0 + 200] + rs.randint(-12, 12, (200, 200, 3)), 0, 255).astype(np.uint8)
            Image.fromarray(img).save(os.path.join(d, f"{j}.png"))
    return root


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dataset")
    ap.add_argument("--model", default=os.path.join(ROOT, "tests", "golden", "age_gender_quantized.pb"))
    ap.add_argument("--input", default="input_1:0")
    ap.add_argument("--output", default="global_pooling/Mean:0")
    ap.add_argument("--phase", default=None)
    ap.add_argument("--input-hw", type=int, default=192)
    ap.add_argument("--crop-center", action="store_true")
    ap.add_argument("--features-file", default=None)
    args = ap.parse_args()
    tmp = None
    if not args.dataset:
        tmp = tempfile.TemporaryDirectory()
        args.dataset = synthetic_dataset(tmp.name)
    features_file = args.features_file or os.path.join(tempfile.gettempdir(), "hfr_example_features.npz")

    if not os.path.exists(features_file) or tmp is not None:
        tfInference = hfr.TensorFlowInference(args.model, input_tensor=args.input, output_tensor=args.output,   # <- same ctor
                                              learning_phase_tensor=args.phase, convert2BGR=True, imageNetUtilsMean=True,
                                              input_hw=args.input_hw)
        dirs_and_files = np.array(get_files(args.dataset))
        dirs, files = dirs_and_files[:, 0], dirs_and_files[:, 1]
        label_enc = preprocessing.LabelEncoder()
        label_enc.fit(dirs)
        y = label_enc.transform(dirs)
        start_time = time.time()
        # the reference: X = np.array([tfInference.extract_features(os.path.join(DATASET_PATH, f), crop_center) for f in files])
        X = tfInference.extract_files([os.path.join(args.dataset, f) for f in files], crop_center=args.crop_center)  # <- batched
        tfInference.close_session()
        print("--- %s seconds ---" % (time.time() - start_time))
        print("X.shape=", X.shape)
        np.savez(features_file, x=X, y=y)

    data = np.load(features_file)
    X, y = data["x"], data["y"]
    X_norm = hfr.normalize(X, norm="l2")                                          # <- preprocessing.normalize
    labels, counts = np.unique(y, return_counts=True)          # identities with a single image cannot be split 50/50
    keep = np.isin(y, labels[counts > 1])
    y = preprocessing.LabelEncoder().fit_transform(y[keep])
    X_norm = X_norm[keep]
    print("after loading: num_classes=", len(np.unique(y)), " X_norm shape:", X_norm.shape)

    pca_components = min(128, X_norm.shape[0] // 2 - 1)
    KNN = hfr.KNeighborsClassifier                                                # <- sklearn.neighbors.KNeighborsClassifier
    classifiers = [
        ["k-NN+PCA", Pipeline(steps=[("pca", hfr.PCA(n_components=pca_components)), ("classifier", KNN(n_neighbors=1, p=2))])],
        ["k-NN", KNN(n_neighbors=1, p=2)],
        ["3-NN", KNN(n_neighbors=3, p=2)],
    ]
    for cls_name, classifier in classifiers:
        print(cls_name)
        classifier_tester(classifier, X_norm, y)
    if tmp is not None:
        tmp.cleanup()


if __name__ == "__main__":
    main()

"""hse_facerec_tf_b200: B200-native (sm_100a) back end for the inference hot path of av-savchenko/HSE_FaceRec_tf.

Drop-in entry points (reference file:line in each module's docstring):
  TensorFlowInference, extract_keras_features, FeatureExtractor   - embedding extraction
  FacialImageProcessing (load_age_gender / age_gender_fun / is_male) - age / gender / identity features
  normalize, KNeighborsClassifier                                 - L2 normalisation + 1-NN identification
Importing this package loads hse_facerec_tf_b200/libhfr.so; there is no CPU fallback.
"""
from ._lib import HfrError, LIB_PATH, lib  # noqa: F401
from .model import HfrModel  # noqa: F401
from .extractor import FeatureExtractor, TensorFlowInference, extract_keras_features  # noqa: F401
from .age_gender import FacialImageProcessing  # noqa: F401
from .classifier import KNeighborsClassifier  # noqa: F401
from .preprocessing import normalize  # noqa: F401
from . import parallel  # noqa: F401
from .decomposition import PCA  # noqa: F401
from .clustering import album_distance_matrix, pairwise_distances  # noqa: F401
from .detection import MTCNN  # noqa: F401
from .staging import crop_resize, expand_and_clamp_boxes, load_resized_batch, resize_pil  # noqa: F401


def launch_count() -> int:
    """Kernels launched by libhfr.so in this process."""
    return int(lib.hfr_launch_count())

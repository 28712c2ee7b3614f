import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
AGE_GENDER_PB = os.path.join(GOLDEN, "age_gender_quantized.pb")


def _ensure_library():
    """A fresh checkout has no libhfr.so (built artefacts are git-ignored): build it (nvcc cross-compiles without a
    GPU) before any test imports the package.  build.py is loaded by path - importing the package would dlopen the
    library it is about to build."""
    if os.path.exists(os.path.join(ROOT, "hse_facerec_tf_b200", "libhfr.so")):
        return           # present (built here, or shipped to the GPU box with the snapshot): never rebuild behind a test run
    import importlib.util
    spec = importlib.util.spec_from_file_location("_hfr_build", os.path.join(ROOT, "hse_facerec_tf_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    _ensure_library()


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def age_gender_pb():
    return AGE_GENDER_PB


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN

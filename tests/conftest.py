import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _load_ref(name):
    path = os.path.join(ROOT, "oracle", "_ref", name + ".so")
    if not os.path.exists(path):
        return None
    import torch  # noqa: F401  (the reference modules link against libtorch)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def ref_a():
    """The UNMODIFIED reference module A (pointnet2._ext) compiled by oracle/build_ref.py, or skip."""
    m = _load_ref("gbref_pointnet2_ext")
    if m is None:
        pytest.skip("oracle/_ref/gbref_pointnet2_ext.so not built")
    return m


@pytest.fixture(scope="session")
def ref_b():
    m = _load_ref("gbref_pointnet2_batch")
    if m is None:
        pytest.skip("oracle/_ref/gbref_pointnet2_batch.so not built")
    return m


@pytest.fixture(scope="session")
def ref_c():
    m = _load_ref("gbref_knn")
    if m is None:
        pytest.skip("oracle/_ref/gbref_knn.so not built")
    return m


@pytest.fixture(scope="session")
def dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def np_rng(seed):
    return np.random.default_rng(seed)

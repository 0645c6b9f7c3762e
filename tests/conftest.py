import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _loader import load_dogm_b200, load_oracle  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def orc():
    return load_oracle()


@pytest.fixture(scope="session")
def dogm_b200():
    return load_dogm_b200()


@pytest.fixture(scope="session")
def gpu(dogm_b200):
    """The product binding, with the CUDA library loaded and a device present (gpu tests only)."""
    dogm_b200.load_library()
    if dogm_b200.device_count() < 1:
        pytest.fail("gpu-marked test started without a CUDA device: there is no CPU fallback")
    return dogm_b200


DEMO = dict(  # reference demo defaults, demo/main.cpp:24-34
    persistence_prob=0.99,
    stddev_process_noise_position=0.1,
    stddev_process_noise_velocity=1.0,
    birth_prob=0.02,
    stddev_velocity=30.0,
    init_max_velocity=30.0,
    freespace_discount=0.01,
)


def make_params(mod, size, resolution, n, b, **over):
    kw = dict(DEMO)
    kw.update(over)
    return mod.Params(
        size,
        resolution,
        n,
        b,
        kw["persistence_prob"],
        kw["stddev_process_noise_position"],
        kw["stddev_process_noise_velocity"],
        kw["birth_prob"],
        kw["stddev_velocity"],
        kw["init_max_velocity"],
        kw["freespace_discount"],
    )


def synthetic_meas(dtype, gs, rng, n_blobs=6, free_level=0.3):
    """A measurement grid with a few occupied blobs on a free background (likelihood = p_A = 1 as in the reference)."""
    occ = np.zeros((gs, gs), np.float32)
    for _ in range(n_blobs):
        cx, cy = rng.integers(2, gs - 2, size=2)
        w, h = rng.integers(1, max(2, gs // 10), size=2)
        occ[max(0, cy - h) : cy + h, max(0, cx - w) : cx + w] = rng.uniform(0.5, 0.95)
    fre = np.where(occ > 0, 0.0, free_level).astype(np.float32)
    meas = np.zeros(gs * gs, dtype=dtype)
    meas["occ_mass"] = occ.ravel()
    meas["free_mass"] = fre.ravel()
    meas["likelihood"] = 1.0
    meas["p_A"] = 1.0
    return meas


def cycle_noise(rng, n, b, p):
    """Noise buffers of one cycle in the shapes dogm_set_noise / oracle_set_noise take."""
    pn = rng.standard_normal((n, 4)).astype(np.float32)
    pn[:, :2] *= np.float32(p.stddev_process_noise_position)
    pn[:, 2:] *= np.float32(p.stddev_process_noise_velocity)
    bn = (rng.standard_normal((b, 2)) * p.stddev_velocity).astype(np.float32)
    iv = rng.uniform(-p.init_max_velocity, p.init_max_velocity, size=(n, 2)).astype(np.float32)
    ru = np.sort(rng.uniform(0.0, 1.0, size=n).astype(np.float32))
    ru = np.minimum(ru, np.float32(1.0) - np.float32(2.0 ** -24))
    return pn, bn, iv, ru

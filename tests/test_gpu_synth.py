"""GPU: the CUDA synthetic generator is bit-identical to the numpy one (any shard anywhere)."""
import numpy as np
import pytest
from numpy.testing import assert_array_equal

pytestmark = pytest.mark.gpu


def test_device_generator_matches_numpy():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from enspara_b200 import synth
    for A, n, first in ((22, 300, 0), (500, 200, 123456), (264, 64, 9_999_999)):
        want = synth.trajectory(n, A, seed=3, first_frame=first)
        got = synth.device_trajectory_aos(n, A, seed=3, first_frame=first).cpu().numpy()
        assert_array_equal(got, want)
    want = synth.features(1000, 64, seed=5, first_row=777)
    got = synth.device_features(1000, 64, seed=5, first_row=777).X.cpu().numpy()
    assert_array_equal(got, want)


def test_device_trajectory_equals_host_upload():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from enspara_b200 import synth
    from enspara_b200.device import DeviceTrajectory
    a = synth.device_trajectory(1000, 50, seed=1, first_frame=40, chunk_frames=300)
    b = DeviceTrajectory.from_host(synth.trajectory(1000, 50, seed=1, first_frame=40))
    assert torch.equal(a.xyz, b.xyz) and torch.equal(a.traces, b.traces)

"""Host-side argument checks of the input feeder (no GPU needed): bad arguments are rejected before any CUDA object is
created, and without a CUDA device the feeder fails loudly instead of falling back to host tensors."""
import pytest
import torch

import chest_x_ray_vit_b200 as pkg


def test_device_feeder_rejects_bad_arguments():
    with pytest.raises(ValueError, match="depth"):
        pkg.data.DeviceFeeder([], depth=1)
    with pytest.raises(ValueError, match="probability"):
        pkg.data.DeviceFeeder([], hflip_p=1.5)
    with pytest.raises(ValueError, match="probability"):
        pkg.data.DeviceFeeder([], hflip_p=-0.1)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour on a machine without a GPU")
def test_device_feeder_needs_a_cuda_device():
    with pytest.raises((RuntimeError, AssertionError)):
        pkg.data.DeviceFeeder([{"pixel_values": torch.zeros(2, 16, 16, dtype=torch.uint8)}], device="cuda:0")

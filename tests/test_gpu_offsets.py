"""-m gpu: the device-built neighbour-offset table equals the oracle's (and hence the reference's qsort order)."""
import ctypes as C

import numpy as np
import pytest

from oracle import refdriver as R
from resynthesizer_b200 import api

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dims", [(7, 5, 9, 4), (64, 48, 32, 32), (301, 200, 256, 256), (1024, 1024, 256, 256)])
def test_device_offsets_equal_oracle(built_oracle, built_lib, dims):
    tw, th, cw, ch = dims
    port = R.load_port()
    w, h = min(tw, cw), min(th, ch)
    n = (2 * w - 1) * (2 * h - 1)
    want = np.zeros((n, 2), np.int32)
    assert port.port_offsets(tw, th, cw, ch, want.ctypes.data, n) == n
    got = api.device_sorted_offsets(tw, th, cw, ch)
    assert got.shape == want.shape and (got == want).all()

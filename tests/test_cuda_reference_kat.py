"""The reference's own unit tests (tests/reference_kat.py) against the CUDA product: each engine
method of core.py is executed on the device, one call at a time, through inv_debug_phase."""
import pytest

import reference_kat as K

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", K.ALL_KATS)
def test_cuda_kat(name):
    from engine_facade import CudaEngine
    getattr(K, name)(lambda w, h: CudaEngine(w, h))

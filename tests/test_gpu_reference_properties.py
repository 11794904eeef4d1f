"""The reference's 89 unit tests (shapes, thresholds: tests/reference_properties.py) run against
the CUDA path through the C ABI, on the inputs for which the oracle satisfies them."""
import pytest

import reference_properties as props

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("check", props.ALL_CHECKS, ids=lambda f: f.__name__)
@pytest.mark.parametrize("scalar,dim", props.CASES)
def test_reference_property_on_device(check, scalar, dim):
    from rusty_compression_b200 import api
    props.run_check(check, api, scalar, dim)


def test_compression_error_semantics():
    """Quirk Q3 and the assert! on tol (src/qr.rs:188, 196-199; src/svd.rs:88, 97-100)."""
    import numpy as np
    from rusty_compression_b200 import api
    mat = np.eye(6)
    with pytest.raises(api.CompressionError):
        api.QR.compute_from(mat).compress(api.ADAPTIVE(1e-3))
    with pytest.raises(api.CompressionError):
        api.SVD.compute_from(mat).compress(api.ADAPTIVE(1e-3))
    with pytest.raises(api.CompressionError):
        api.LQ.compute_from(mat).compress(api.ADAPTIVE(1e-3))
    with pytest.raises(AssertionError):
        api.QR.compute_from(mat).compress(api.ADAPTIVE(1.5))
    # rank clamps (src/qr.rs:172-174)
    assert api.QR.compute_from(mat).compress(api.RANK(100)).rank() == 6

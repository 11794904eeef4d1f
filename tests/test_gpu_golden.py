"""The CUDA path (through the C ABI) against the frozen fixtures of tests/golden/: index vectors bit-exact wherever
the pivot gap exceeds 1e-6 (every step validated in double precision, no skip on a tie -- see golden_common.py),
singular values / residuals / ID errors within 1e-10 (f64, c64) and 1e-4 (f32, c32)."""
import numpy as np
import pytest

import golden_common as gc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from rusty_compression_b200 import api as a
    return a


@pytest.mark.parametrize("name", gc.CASES)
def test_device_reproduces_golden_case(api, name):
    g = gc.load(name)
    ctx = api.default_context()
    ctx.reset_counters()

    def make_stream(blocks):
        return {"omega": blocks[0]}

    def adaptive(a, tol, s, blocks):
        return api.sample_range_adaptive(a, tol, s, omega_blocks=blocks)

    def conj_matmat(a, q):      # the device's own A^H q: the factor b it pivots is built from this product
        return api.DeviceMatrix.from_numpy(a).conj_matmat(q).to_numpy()

    gc.run_case(api, name, g, make_stream, adaptive, conj_matmat=conj_matmat)
    assert ctx.counter("kernel_launches") > 100, "CUDA path did not run"


def test_device_config2_toy(api):
    g = dict(np.load(gc.GOLDEN + "/config2_toy_f64.npz"))
    from oracle import reference_path as ref
    q = api.sample_range_power_iteration(g["a"], 16, 4, 2, omega=g["omega"])
    svd = api.SVD.compute_from_range_estimate(q, g["a"])
    gc.close(svd.s_f64(), g["rsvd_s"], 1e-10, "config-2 toy singular values")
    gc.close(ref.range_residual(g["a"], q), g["residual"], 1e-10, "config-2 toy residual")


def test_device_permutation_known_answers_of_the_reference(api):
    """src/permutation.rs:192-239 verbatim, on the device gather kernels."""
    k = gc.permutation_answers()
    mat, perm, vec = np.array(k["matrix"]), np.array(k["perm"]), np.array(k["vector"])
    for mode in ("COL", "COLINV", "ROW", "ROWINV"):
        assert np.array_equal(api.apply_permutation_matrix(mat, perm, mode), np.array(k[mode])), mode
    for mode in ("NOINV", "INV"):
        assert np.array_equal(api.apply_permutation_vector(vec, perm, mode), np.array(k[mode])), mode

"""The oracle against the frozen fixtures of tests/golden/ (CPU).  Guards the restatement and the
LAPACK it runs on against drift; the same fixtures are used on the B200 by tests/test_gpu_golden.py."""
import numpy as np
import pytest

import golden_common as gc
from oracle import reference_path as ref


@pytest.mark.parametrize("name", gc.CASES)
def test_oracle_reproduces_golden_case(name):
    g = gc.load(name)

    def make_stream(blocks):
        return {"omega_stream": ref.OmegaStream(g["a"].dtype, blocks=blocks)}

    def adaptive(a, tol, s, blocks):
        nb = blocks.shape[1] // s
        stream = ref.OmegaStream(a.dtype, blocks=[blocks[:, i * s:(i + 1) * s] for i in range(nb)])
        return ref.sample_range_adaptive(a, tol, s, stream)

    gc.run_case(ref, name, g, make_stream, adaptive, oracle_side=True)


def test_oracle_config2_toy():
    g = dict(np.load(gc.GOLDEN + "/config2_toy_f64.npz"))
    q = ref.sample_range_power_iteration(g["a"], 16, 4, 2, ref.OmegaStream(np.float64, blocks=[g["omega"]]))
    svd = ref.SVD.compute_from_range_estimate(q, g["a"])
    gc.close(svd.s, g["rsvd_s"], 1e-10, "config-2 toy singular values")
    gc.close(ref.range_residual(g["a"], q), g["residual"], 1e-10, "config-2 toy residual")
    # the rSVD recovers the generator's leading spectrum
    gc.close(svd.s[:8], g["sigma"][:8], 1e-6, "leading spectrum")


def test_permutation_known_answers_of_the_reference():
    """src/permutation.rs:192-239 verbatim."""
    k = gc.permutation_answers()
    mat, perm, vec = np.array(k["matrix"]), np.array(k["perm"]), np.array(k["vector"])
    for mode in ("COL", "COLINV", "ROW", "ROWINV"):
        assert np.array_equal(ref.apply_permutation_matrix(mat, perm, mode), np.array(k[mode])), mode
    for mode in ("NOINV", "INV"):
        assert np.array_equal(ref.apply_permutation_vector(vec, perm, mode), np.array(k[mode])), mode

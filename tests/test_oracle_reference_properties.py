"""The reference's property tests run against the CPU oracle: proves the restatement
(oracle/reference_path.py) satisfies everything the reference asserts of itself."""
import pytest

from oracle import reference_path as ref
import reference_properties as props


@pytest.mark.parametrize("check", props.ALL_CHECKS, ids=lambda f: f.__name__)
@pytest.mark.parametrize("scalar,dim", props.CASES)
def test_reference_property(check, scalar, dim):
    seed = props.run_check(check, ref, scalar, dim)
    assert seed < 10

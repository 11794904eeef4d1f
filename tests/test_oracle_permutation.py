"""The reference's only exact known-answer tests (src/permutation.rs:192-239), verbatim."""
import numpy as np

from oracle import reference_path as ref


def test_matrix_permutation():
    mat = np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0], [7.0, 8.0, 9.0]])
    right_row = np.array([[7.0, 8.0, 9.0], [1.0, 2.0, 3.0], [4.0, 5.0, 6.0]])
    left_row = np.array([[4.0, 5.0, 6.0], [7.0, 8.0, 9.0], [1.0, 2.0, 3.0]])
    right_col = np.array([[3.0, 1.0, 2.0], [6.0, 4.0, 5.0], [9.0, 7.0, 8.0]])
    left_col = np.array([[2.0, 3.0, 1.0], [5.0, 6.0, 4.0], [8.0, 9.0, 7.0]])
    perm = np.array([2, 0, 1])
    assert np.array_equal(right_col, ref.apply_permutation_matrix(mat, perm, "COL"))
    assert np.array_equal(left_col, ref.apply_permutation_matrix(mat, perm, "COLINV"))
    assert np.array_equal(right_row, ref.apply_permutation_matrix(mat, perm, "ROW"))
    assert np.array_equal(left_row, ref.apply_permutation_matrix(mat, perm, "ROWINV"))


def test_vector_permutation():
    vec = np.array([1.0, 2.0, 3.0])
    perm = np.array([2, 0, 1])
    assert np.array_equal(np.array([3.0, 1.0, 2.0]), ref.apply_permutation_vector(vec, perm, "NOINV"))
    assert np.array_equal(np.array([2.0, 3.0, 1.0]), ref.apply_permutation_vector(vec, perm, "INV"))


def test_invert():
    perm = np.array([3, 0, 2, 1])
    inv = ref.invert_permutation_vector(perm)
    assert np.array_equal(inv[perm], np.arange(4))

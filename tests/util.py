"""Parity helpers.  Tolerances come from BASELINE.json's north_star: fp32 mode 1e-5
relative (absolute floor: values are cosines / probabilities of O(1), so the floor is the
tolerance itself, |d| <= tol * max(1, |ref|)); bf16-vault mode 1e-2; top-k row ids exact
wherever the neighbouring score gap exceeds the tolerance."""
import numpy as np

FP32_TOL = 1e-5
BF16_TOL = 1e-2


def assert_close(got, ref, tol=FP32_TOL, what=""):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    both_nan = np.isnan(got) & np.isnan(ref)
    err = np.where(both_nan, 0.0, np.abs(got - ref))
    bound = tol * np.maximum(1.0, np.where(both_nan, 0.0, np.abs(ref)))
    bad = ~(err <= bound)
    assert not bad.any(), f"{what}: {bad.sum()} values off, max err {np.nanmax(err):.3e} (tol {tol})"


def assert_topk(rows, scores, ref_rows, ref_scores, tol=FP32_TOL, what=""):
    """scores must agree rank by rank; rows must agree except inside near-ties (gap <= 2*tol)."""
    rows, ref_rows = np.asarray(rows), np.asarray(ref_rows)
    assert_close(scores, ref_scores, tol, what + " scores")
    diff = rows != ref_rows
    if diff.any():
        rs = np.asarray(ref_scores, np.float64)
        for q, j in zip(*np.nonzero(diff)):
            near = [abs(rs[q, j] - rs[q, jj]) for jj in (j - 1, j + 1) if 0 <= jj < rs.shape[1]]
            # the row may also have swapped with one just outside the top-k: accept if the reported
            # score is within tol of the reference score at this rank (already checked) AND some
            # neighbour is within 2*tol, or it is the last rank
            assert (near and min(near) <= 2 * tol) or j == rs.shape[1] - 1, \
                f"{what}: row mismatch at query {q} rank {j}: {rows[q, j]} vs {ref_rows[q, j]} with gap {near}"

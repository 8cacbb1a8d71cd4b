"""The arithmetic behind the hi-only filter of fp32 stores, checked in numpy on the CPU:
  * bf16_contraction_eps (csrc/common.cuh), restated here, really bounds |q.x - bf16(q).bf16(x)| -- on random rows,
    on rows built to be as bad as Cauchy-Schwarz allows, on rows of very different norms;
  * the set the kernels re-score -- every row whose approximate distance is within 2 eps of the approximate k-th
    best -- always contains the exact top-k (ties included), which is what makes the result exact without an
    a-posteriori guard (csrc/kernels.h, RefineFilterArgs)."""
import numpy as np
import pytest

from oracle.exact_search import round_to_bf16
from tests.conftest import unit_rows

GUARD_REL = 1.2e-4          # kernels.h: kGuardRel


def contraction_eps(q, x_max_norm2, x_lo_max2, l2=False):
    """common.cuh: bf16_contraction_eps, per query (fp32 arithmetic like the device)."""
    q = q.astype(np.float32)
    qn = np.sqrt(np.sum(q * q, axis=1, dtype=np.float32))
    ql = np.sqrt(np.sum((q - round_to_bf16(q)) ** 2, axis=1, dtype=np.float32))
    xn, xl = np.float32(np.sqrt(x_max_norm2)), np.float32(np.sqrt(x_lo_max2))
    eps = np.float32(GUARD_REL) * qn * xn + np.float32(1.01) * (ql * xn + (qn + ql) * xl)
    return 2.0 * eps if l2 else eps


def store_bounds(x):
    return float(np.max(np.sum(x * x, axis=1))), float(np.max(np.sum((x - round_to_bf16(x)) ** 2, axis=1)))


def approx_and_exact_dots(q, x):
    exact = q.astype(np.float64) @ x.astype(np.float64).T
    approx = (round_to_bf16(q).astype(np.float32) @ round_to_bf16(x).astype(np.float32).T).astype(np.float64)   # fp32 accumulate
    return approx, exact


@pytest.mark.parametrize("dim", [16, 384, 768])
def test_eps_bounds_the_contraction_error(dim):
    rng = np.random.default_rng(dim)
    x = unit_rows(4000, dim, dim + 1)
    x[::7] *= (10.0 ** rng.uniform(-2, 2, size=(len(x[::7]), 1))).astype(np.float32)      # norms over four decades
    q = rng.standard_normal((64, dim)).astype(np.float32) * (10.0 ** rng.uniform(-1, 1, size=(64, 1))).astype(np.float32)
    # adversarial rows: aligned with what the rounding of a query drops, and rows whose own rounding error is
    # aligned with a query -- the two cases in which Cauchy-Schwarz is tight
    q_lo = q - round_to_bf16(q)
    x[1:9] = (q_lo[:8] / np.linalg.norm(q_lo[:8], axis=1, keepdims=True)).astype(np.float32)
    approx, exact = approx_and_exact_dots(q, x)
    eps = contraction_eps(q, *store_bounds(x))
    err = np.abs(approx - exact)
    assert np.all(err <= eps[:, None]), float(np.max(err / eps[:, None]))
    assert np.max(err / eps[:, None]) > 0.05          # ... and is not vacuous: the aligned rows come within 20x


@pytest.mark.parametrize("space,k", [("cosine", 10), ("l2", 16), ("cosine", 1)])
def test_rows_within_two_eps_of_the_approximate_kth_best_contain_the_exact_top_k(space, k):
    rng = np.random.default_rng(k)
    n, dim, B = 6000, 256, 48
    centre = unit_rows(1, dim, 3)[0]
    x = np.vstack([unit_rows(n // 2, dim, 4),                                   # spread out
                   (centre + 2e-2 * rng.standard_normal((n // 2, dim))).astype(np.float32)])      # and a dense cluster
    x[100:140] = x[99]                                                          # exact duplicates: ties at every rank
    q = np.vstack([unit_rows(B // 2, dim, 5), (centre + 2e-2 * rng.standard_normal((B // 2, dim))).astype(np.float32)])
    q[0] = x[99]
    if space == "cosine":
        x = (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)
        q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    approx_dot, exact_dot = approx_and_exact_dots(q, x)
    if space == "l2":
        qn, xn = np.sum(q.astype(np.float64) ** 2, 1)[:, None], np.sum(x.astype(np.float64) ** 2, 1)[None, :]
        approx, exact = np.maximum(qn + xn - 2 * approx_dot, 0.0), qn + xn - 2 * exact_dot
    else:
        approx, exact = 1.0 - approx_dot, 1.0 - exact_dot
    eps = contraction_eps(q, *store_bounds(x), l2=(space == "l2")).astype(np.float64)
    sizes = []
    for b in range(len(q)):
        a_k = np.sort(approx[b])[k - 1]
        candidates = set(np.nonzero(approx[b] <= a_k + 2.0 * eps[b])[0].tolist())
        order = np.lexsort((np.arange(n), exact[b]))                           # exact order, ties by lower row
        d_k = exact[b][order[k - 1]]
        needed = set(np.nonzero(exact[b] <= d_k)[0].tolist())                  # the top k and everything tied with its last
        assert needed <= candidates, (b, sorted(needed - candidates))
        sizes.append(len(candidates))
    assert min(sizes) >= k and np.median(sizes) < n / 4                         # a filter, not the whole corpus


@pytest.mark.parametrize("cpm,k", [(148, 10), (18, 10), (18, 16), (9, 16), (4, 10), (1, 7), (37, 1), (74, 16)])
def test_group_bound_never_undercuts_the_global_kth_best(cpm, k):
    """tensor_regime.cu, Args::tau_grp: min(cpm, k) groups of CTAs (cj % groups); every CTA atomicMin-s its own r-th
    best, r = ceil(k / groups), into its group's slot; T = the largest slot.  At ANY moment of the scan -- every
    CTA having seen an arbitrary prefix of its rows -- T must be at or above the k-th best of all rows (else rows of
    the true top-k would be rejected), and once everything is scanned it should be close to it."""
    rng = np.random.default_rng(cpm * 100 + k)
    groups = min(cpm, k, 16)
    r = -(-k // groups)
    assert groups * r >= k
    for trial in range(20):
        rows_per_cta = rng.integers(r, 400, size=cpm)
        d = [rng.standard_normal(n) for n in rows_per_cta]                  # distances of each CTA's rows, in scan order
        kth_global = np.sort(np.concatenate(d))[k - 1]
        for frac in (0.05, 0.3, 0.7, 1.0):
            slots = np.full(groups, np.inf)
            for cj in range(cpm):
                seen = d[cj][: max(int(len(d[cj]) * rng.uniform(frac * 0.5, frac) + 0.999), 0)]
                if len(seen) >= r:                                          # a CTA publishes once its list holds r rows
                    slots[cj % groups] = min(slots[cj % groups], np.sort(seen)[r - 1])
            T = slots.max()                                                 # +inf until every group has published
            assert T >= kth_global
    # ... and it is a useful bound: with every CTA done (300 rows each) it sits within a small multiple of k ranks
    d = [rng.standard_normal(300) for _ in range(cpm)]
    slots = np.full(groups, np.inf)
    for cj in range(cpm):
        slots[cj % groups] = min(slots[cj % groups], np.sort(d[cj])[r - 1])
    rank_of_T = int(np.sum(np.concatenate(d) <= slots.max()))
    assert k <= rank_of_T <= 8 * k * (1 + np.log(k)) + 8 * r + 8, rank_of_T

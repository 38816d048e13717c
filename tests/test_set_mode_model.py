"""Property test of the SET-MODE membership rule of the coarse refine (quake_b200/csrc/refine.cuh:
set_select_and_emit / exact_value_bounds), restated in numpy -- host arithmetic only, no GPU.

The rule: given the kc best candidates by filter score s (everything else has s >= a_score), [lo(s), hi(s)] bounds the
value the reference orders by (sqrt of its float32 squared distance) of any row with filter score s, PROVIDED the filter
score is within the modelled error of the true score. A candidate with hi(s) < lo(s_{k+1}) is a sure member, one with
lo(s) > hi(s_k) is surely out, the rows outside the candidate set must be surely out, the rest is decided by the exact
value (ties by id). Whatever perturbation of the filter scores inside the error bound, the returned set must be the
exact top-k by (reference distance, id) -- or the rule must decline (None)."""
import numpy as np
import pytest

from oracle import oracle as orc

EPS = 5.960464477539063e-08  # 2^-24


def bounds(s, qn, qnorm, U, fgam, d):
    """exact_value_bounds<l2>, not squared: float32 [lo, hi] of sqrt_rn(reference squared distance)."""
    gam = (d + 8) * EPS
    e2 = (d // 8 + 12) * EPS
    s = np.asarray(s, dtype=np.float64)
    e1 = gam * U * U + 2.0 * (gam + fgam) * qnorm * U + 4.0 * EPS * np.abs(s)
    lb = (qn * (1.0 - gam) + s - e1) * (1.0 - e2)
    ub = (qn * (1.0 + gam) + s + e1) * (1.0 + e2)
    lbf = np.nextafter(lb.astype(np.float32), np.float32(-np.inf))  # at or below the round-down
    ubf = np.nextafter(ub.astype(np.float32), np.float32(np.inf))   # at or above the round-up
    lo = np.where(lbf > 0, np.sqrt(np.maximum(lbf, 0).astype(np.float32)), np.float32(0))
    hi = np.where(ubf > 0, np.sqrt(np.maximum(ubf, 0).astype(np.float32)), np.float32(0))
    return lo.astype(np.float32), hi.astype(np.float32)


def set_select(cand_rows, cand_s, a_score, have_rejects, k, exact, ids, qn, qnorm, U, fgam, d):
    """The kernel's decision procedure. cand_*: the m > k candidates (any order). Returns a set of rows or None."""
    order = np.lexsort((np.arange(len(cand_s)), cand_s))
    rows, s = cand_rows[order], cand_s[order]
    lo, hi = bounds(s, qn, qnorm, U, fgam, d)
    lo_klo, hi_klo = lo[k - 1], hi[k - 1]
    lo_khi = lo[k]
    if have_rejects:
        lo_r, _ = bounds(np.array([a_score]), qn, qnorm, U, fgam, d)
        if not lo_r[0] > hi_klo:
            return None
    member = hi < lo_khi
    out = lo > hi_klo
    amb = ~member & ~out
    need = k - int(member.sum())
    na = int(amb.sum())
    if need < 0 or need > na:
        return None
    chosen = set(rows[member].tolist())
    if 0 < need < na:
        ar = rows[amb]
        o = np.lexsort((ids[ar], exact[ar]))
        chosen |= set(ar[o[:need]].tolist())
    elif need == na:
        chosen |= set(rows[amb].tolist())
    return chosen


@pytest.mark.parametrize("fgam", [7.62939453125e-06, 9.765625e-04 + 7.62939453125e-06])
@pytest.mark.parametrize("dup", [False, True])
def test_set_mode_rule_returns_the_exact_set_or_declines(fgam, dup):
    rng = np.random.default_rng(11 if dup else 5)
    d, n, k = 32, 600, 24
    kc = k + 6
    decided = declined = refined = 0
    for trial in range(60):
        V = rng.standard_normal((n, d)).astype(np.float32)
        q = rng.standard_normal(d).astype(np.float32)
        if dup:  # exact duplicates and near-duplicates of the row that sits at the k-th place: ties AT the boundary
            r = int(np.argsort(orc.pairwise(q[None, :], V, "l2")[0], kind="stable")[k - 1])
            far = np.argsort(orc.pairwise(q[None, :], V, "l2")[0])[-5:]
            V[far[:3]] = V[r]
            V[far[3:]] = V[r] + (rng.standard_normal((2, d)) * 1e-5).astype(np.float32)
        ids = rng.permutation(n).astype(np.int64)
        d2 = orc.pairwise(q[None, :], V, "l2")[0]              # the reference's float32 squared distances
        exact = np.sqrt(d2).astype(np.float32)                 # what it orders by (list_scanning.h:260)
        want = set(np.lexsort((ids, exact))[:k].tolist())
        qn = float(np.dot(q.astype(np.float64), q.astype(np.float64)))
        qnorm = np.sqrt(qn)
        U = float(np.sqrt((V.astype(np.float64) ** 2).sum(1).max()))
        true_s = (V.astype(np.float64) ** 2).sum(1) - 2.0 * V.astype(np.float64) @ q.astype(np.float64)
        # a filter inside its modelled error: the dot-product part of e1, adversarially signed near the boundary
        err = 2.0 * fgam * qnorm * U
        noise = rng.uniform(-1, 1, n) * err * 0.95
        kth = np.partition(true_s, k)[k]
        noise = np.where(true_s <= kth, np.abs(noise), -np.abs(noise)) if trial % 2 else noise  # push the two sides together
        s = (true_s + noise).astype(np.float32)
        cand = np.argsort(s, kind="stable")[:kc]
        a_score = float(np.sort(s)[kc - 1])
        got = set_select(cand, s[cand], a_score, True, k, exact, ids, qn, qnorm, U, fgam, d)
        if got is None:
            declined += 1
            continue
        decided += 1
        assert got == want, (trial, sorted(got ^ want))
        lo, hi = bounds(np.sort(s[cand]), qn, qnorm, U, fgam, d)
        refined += int(((~(hi < lo[k])) & (~(lo > hi[k - 1]))).sum() > 0)
    assert decided > 0
    if dup:
        assert refined > 0  # ties at the boundary are settled by exact values and ids, not by the filter
    elif fgam < 1e-4:
        assert declined == 0 and refined < decided  # the tight filter settles most queries without exact arithmetic

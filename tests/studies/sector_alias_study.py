"""Design study for DESIGN.md 8(3) (CPU only, numpy): a SECTOR-level alias table for the hubness-weighted negative
sampler.  The 4 lanes of an aligned group share the sector draw and its accept/alias decision (coalesced lookups) and
each lane picks its row inside the final sector from its own uniform.  Checks that every lane's negatives follow the
hubness law clamp(in_degree, 1, n) / sum exactly (chi-square), which is all the optimizer depends on.
Usage: python tests/studies/sector_alias_study.py"""
import numpy as np
from scipy import stats


def vose(q_in):
    n = len(q_in)
    q = np.asarray(q_in, np.float64) * n / np.sum(q_in)
    prob = np.ones(n); alias = np.arange(n)
    small = [i for i in range(n) if q[i] < 1.0]; large = [i for i in range(n) if q[i] >= 1.0]
    while small and large:
        s, l = small.pop(), large.pop()
        prob[s] = q[s]; alias[s] = l
        q[l] = (q[l] + q[s]) - 1.0
        (small if q[l] < 1.0 else large).append(l)
    return prob, alias


def build(w):
    """sector weights, Vose table over sectors, 3 cumulative thresholds per sector (missing rows of the last sector
    weigh 0, so they are never selected)."""
    n = len(w)
    nsec = (n + 3) // 4
    wp = np.zeros(nsec * 4); wp[:n] = w
    ws = wp.reshape(nsec, 4)
    W = ws.sum(axis=1)
    prob, alias = vose(W)
    cdf = np.cumsum(ws, axis=1) / W[:, None]
    return prob, alias, cdf[:, :3]


def draw(prob, alias, thr, n_groups_draws, rng):
    """one negative slot for `n_groups_draws` groups of 4 lanes: shared sector + shared accept, per-lane row choice"""
    nsec = len(prob)
    s0 = rng.integers(0, nsec, n_groups_draws)                 # group-shared word -> random sector
    ua = rng.random(n_groups_draws)                            # group-shared accept word
    s = np.where(ua < prob[s0], s0, alias[s0])
    # per-lane uniforms: lane r uses frac(u + r/4) of one more shared uniform (a rotation, like the uniform sampler):
    # marginally uniform for every lane, correlated across the lanes of the group only
    u = rng.random(n_groups_draws)
    out = np.empty((n_groups_draws, 4), np.int64)
    for r in range(4):
        ul = (u + r / 4.0) % 1.0
        out[:, r] = 4 * s + (ul >= thr[s, 0]).astype(int) + (ul >= thr[s, 1]) + (ul >= thr[s, 2])
    return out


def main():
    rng = np.random.default_rng(0)
    n = 4093                                                   # not a multiple of 4: partial last sector
    deg = np.clip(rng.zipf(1.6, n), 1, n).astype(np.float64)   # heavy-tailed in-degrees, clamp(., 1, n)
    prob, alias, thr = build(deg)
    law = deg / deg.sum()
    draws = draw(prob, alias, thr, 4_000_000, rng)
    assert draws.max() < n
    for r in range(4):
        cnt = np.bincount(draws[:, r], minlength=n)
        # merge rare nodes so that expected counts are >= 20
        order = np.argsort(law); e = law[order] * len(draws); c = cnt[order]
        edges = np.nonzero(np.diff(np.floor(np.cumsum(e) / 20.0)))[0] + 1
        ce = np.add.reduceat(c, np.r_[0, edges]); ee = np.add.reduceat(e, np.r_[0, edges])
        chi = stats.chisquare(ce, ee * ce.sum() / ee.sum())
        print(f"lane {r}: chi2 p-value {chi.pvalue:.3f} over {len(ce)} bins; hub node share {cnt[np.argmax(deg)] / len(draws):.5f} vs law {law.max():.5f}")
        assert chi.pvalue > 1e-3
    # lanes of one group land in the same sector (that is the point: one coalesced row sector)
    assert np.all(draws // 4 == draws[:, :1] // 4)
    print("sector-level alias sampler: per-lane law exact, 1 table sector + 1 row sector per group and negative")


if __name__ == "__main__":
    main()

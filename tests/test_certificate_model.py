"""Soundness of the nearest-neighbour certificates of the ICP loop (cupoch_b200/csrc/icp_kernels.cuh, "Certificates";
cphb_internal.cuh, WarpSearchC), checked on the CPU against exhaustive search.

This is a numpy restatement of the *decision rule* the kernel applies, with every approximation pushed in the
direction that makes a wrong certificate MORE likely (approximate sqrt erring by its full 2^-23 in the unsafe
direction), driven through sequences of shrinking rigid motions like a converging ICP, plus random jitter so
that lanes keep dropping out and re-searching.  Whenever the rule says "skip the search", the oracle's exhaustive
search (same float32 (d2, index) key arithmetic as the kernels) must return exactly the carried match.

The rule (per source point, between two searches):
  slack  : lower bound on the distance to every target point other than the match, rounded down
  step   : delta = |q' - q| rounded up;  slack <- slack - delta (rounded down)
  matched: certified  iff  key(d2(q', match)) < init  and  sqrt(d2(q', match)) * (1+1e-5) < slack
  no match: certified iff  slack > r * (1+2e-5)
A search refreshes slack = sqrt(min(second-smallest d2 [nearest if unmatched], (sqrt(best d2 or r2) + margin)^2))
* (1 - 1e-5): everything inside the relaxed bound was evaluated, everything else is outside it.
"""
import numpy as np
import pytest

from oracle import oracle_py as orc

F = np.float32
APPROX = F(1.0 + 2.0 ** -23)  # worst-case relative error of sqrt.approx.f32


def up(x):
    return np.nextafter(x.astype(F), F(np.inf))


def down(x):
    return np.nextafter(x.astype(F), F(-np.inf))


def d2_of(a, b):
    """the kernels' distance arithmetic: fma(dz,dz, fma(dx,dx, dy*dy)) in float32 (emulated in float64: products of
    float32 differences are exact in float64, one rounding per fma)"""
    d = (a.astype(F) - b.astype(F)).astype(np.float64)
    t = (d[:, 1] * d[:, 1]).astype(F).astype(np.float64)
    t = (d[:, 0] * d[:, 0] + t).astype(F).astype(np.float64)
    return (d[:, 2] * d[:, 2] + t).astype(F)


def rigid(rng, angle, shift):
    ax = rng.standard_normal(3)
    ax /= np.linalg.norm(ax)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = R
    T[:3, 3] = rng.standard_normal(3) * shift
    return T


def search_refresh(tgt, q, r2, margin):
    """what one search leaves behind: (match, slack) per query"""
    idx, d2, _ = orc.search(tgt, q, 2)
    matched = d2[:, 0] < r2
    match = np.where(matched, idx[:, 0], -1).astype(np.int32)
    base = np.where(matched, d2[:, 0], r2).astype(F)
    # relaxed bound: approximate sqrt padded upwards, sums and squares rounded up (WarpSearchC::refresh)
    e = up(up(np.sqrt(base) * APPROX * F(1.000001)) + margin)
    rb = np.maximum(up(e * e), base)
    rb = np.where(margin > 0, rb, base)
    other = np.where(matched, d2[:, 1], d2[:, 0]).astype(F)
    l2 = np.minimum(other, rb)
    # approximate sqrt erring upwards (the unsafe direction), then the kernel's guard
    slack = down(np.sqrt(l2) * APPROX * F(0.99999))
    return match, slack.astype(F)


def certify(tgt, q_old, q_new, match, slack, r2, r_up):
    disp = up(np.sqrt(d2_of(q_new, q_old)) / APPROX * F(1.00001))  # sqrt erring downwards = unsafe
    slk = down(slack - disp)
    safe = np.maximum(match, 0)
    d2p = d2_of(q_new, tgt[safe])
    dn = up(np.sqrt(d2p) / APPROX * F(1.00001))
    with np.errstate(invalid="ignore"):
        cert_m = (match >= 0) & (d2p < r2) & (dn < slk)
        cert_u = (match < 0) & (slk > r_up)
    return cert_m | cert_u, slk.astype(F), disp


@pytest.mark.parametrize("cloud", ["cube", "surface", "lattice"])
def test_certified_lanes_keep_their_exact_match(cloud):
    rng = np.random.default_rng(7)
    n_t, n_q = 4000, 1500
    if cloud == "cube":
        tgt = rng.random((n_t, 3), dtype=np.float32)
        r = F(0.06)
    elif cloud == "surface":
        xy = rng.random((n_t, 2))
        tgt = np.c_[xy, 0.1 * np.sin(4 * np.pi * xy[:, 0]) * np.cos(4 * np.pi * xy[:, 1])].astype(np.float32)
        r = F(0.03)
    else:  # exact ties: a regular lattice, queries start on cell centres
        g = np.arange(16, dtype=np.float32) / 16
        tgt = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
        r = F(0.08)
    pick = rng.integers(0, len(tgt), n_q)
    q = (tgt[pick] + (rng.standard_normal((n_q, 3)) * 0.004).astype(np.float32)).astype(np.float32)
    if cloud == "lattice":
        q[: n_q // 3] = tgt[pick[: n_q // 3]] + F(1.0 / 32)   # equidistant from 8 lattice points
    q[-50:] += F(0.5)  # some queries far outside: unmatched lanes
    r2 = F(r * r)
    r_up = F(np.sqrt(np.float64(r2)) * 1.00002)

    match, slack = search_refresh(tgt, q, r2, np.zeros(n_q, F))
    n_cert = n_cert_unmatched = 0
    step = 0.01
    for it in range(40):
        T = rigid(rng, step * 0.3, step * 0.01)
        q_new = orc.transform_points(q, T)
        if it % 7 == 3:  # jitter a tenth of the points so certificates keep failing somewhere
            j = rng.random(n_q) < 0.1
            q_new[j] += (rng.standard_normal((int(j.sum()), 3)) * 0.003).astype(np.float32)
        cert, slk, disp = certify(tgt, q, q_new, match, slack, r2, r_up)
        # ground truth for every lane (exhaustive search, same key arithmetic)
        idx, d2, _ = orc.search(tgt, q_new, 1, float(r))
        truth = idx[:, 0]
        bad = cert & (truth != match)
        assert not bad.any(), "iteration %d: %d certified lanes whose nearest neighbour changed" % (it, int(bad.sum()))
        n_cert += int(cert.sum())
        n_cert_unmatched += int((cert & (match < 0)).sum())
        # uncertified lanes search again, with the margin the kernel would use (4 x displacement, uncapped here:
        # a larger margin only makes the bound weaker to prove, never unsound)
        m2, s2 = search_refresh(tgt, q_new, r2, up(F(4.0) * disp))
        assert (m2[~cert] == truth[~cert]).all()
        match = np.where(cert, match, m2).astype(np.int32)
        slack = np.where(cert, slk, s2).astype(F)
        q = q_new
        step *= 0.75
    assert n_cert > 10 * n_q, "the test must exercise certificates (got %d)" % n_cert
    assert n_cert_unmatched > 0

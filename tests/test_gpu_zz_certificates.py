"""GPU invariants of the ICP certificates and tile schedules (icp_kernels.cuh): they may only change how fast a result is
found, never a bit of it.  Kept in its own file, last in collection order."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import cupoch_b200 as cph
from cupoch_b200.testing import datagen

R = cph.registration


def cloud(p, n=None, c=None):
    pc = cph.geometry.PointCloud(p)
    if n is not None:
        pc.normals = n
    if c is not None:
        pc.colors = c
    return pc


@pytest.mark.parametrize("kind", ["p2plane", "p2p", "colored"])
def test_certificates_do_not_change_results(kind, monkeypatch):
    """The certificates only ever skip a search whose answer is already known: with them switched off
    (CPHB_CERT_GAIN=0) every iteration does the full search and must produce the same bits -- pose, fitness, rmse
    and correspondence set -- and the same again under the static and the dynamic tile schedule."""
    n = 60000
    tgt, tn = datagen.surface(n, 11)
    tc = datagen.texture(tgt)
    src, sn, sc = datagen.make_source(tgt, datagen.gt_transform((-1.0, 1.5, 2.0), (0.01, -0.005, 0.008)), 13, 14, 3e-4,
                                      attrs=[(tn, True), (tc, False)])
    crit = R.ICPConvergenceCriteria(0, 0, 25)

    def run():
        if kind == "p2plane":
            return R.registration_icp(cloud(src), cloud(tgt, tn), 0.02, np.eye(4), R.TransformationEstimationPointToPlane(), crit)
        if kind == "p2p":
            return R.registration_icp(cloud(src), cloud(tgt), 0.02, np.eye(4), R.TransformationEstimationPointToPoint(), crit)
        return R.registration_colored_icp(cloud(src, sn, sc), cloud(tgt, tn, tc), 0.02, np.eye(4), crit)

    ref = run()
    for env in ({"CPHB_CERT_GAIN": "0"}, {"CPHB_STATIC_SCHED": "0"}, {"CPHB_CERT_GAIN": "16", "CPHB_CERT_CAP": "4"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        other = run()
        for k in env:
            monkeypatch.delenv(k)
        np.testing.assert_array_equal(ref.transformation, other.transformation)
        np.testing.assert_array_equal(ref.correspondence_set, other.correspondence_set)
        assert ref.fitness == other.fitness and ref.inlier_rmse == other.inlier_rmse
    assert ref.fitness > 0.9

"""The C oracle against golden vectors generated from the compiled, unmodified reference
(tests/golden/make_golden.py).  Runs anywhere (no /root/reference needed): this is what pins the oracle on the GPU box."""
import pytest

import golden_check
from oracle_py import Oracle


@pytest.mark.parametrize("name", golden_check.fixtures())
def test_oracle_matches_golden(name):
    g, planes, stages = golden_check.load(name)
    o = Oracle(planes)
    golden_check.check(
        g, stages,
        alpha=o.alpha,
        gradient_pass=o.gradient_pass,
        range1d=o.range1d,
        range_dyn=lambda n, m3: o.range_dyn(n, mode3=m3, want_dst=True),
        chroma=lambda cfg, modes: o.chroma(cfg, modes),
        state=lambda: dict(smoothMap=o.state(0), mipmapMask=o.state(1), mapSmoothTile=[o.state(2 + i) for i in range(3)],
                           mappedRGB=[o.state(5 + i) for i in range(3)], recon=[o.state(8 + i) for i in range(3)]))
    o.close()


def test_fixtures_exist():
    assert len(golden_check.fixtures()) >= 10

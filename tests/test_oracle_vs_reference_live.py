"""Live differential test against oracle/_ref/bsmap (the unmodified reference compiled by
oracle/Makefile).  Skipped where the binary is absent.  Uses fresh seeds, so it fuzzes beyond the
committed fixtures."""
import os

import pytest

import cases as CS
import oracle_lib as O
import runners as R

pytestmark = pytest.mark.skipif(not os.path.exists(O.REF_BIN), reason="oracle/_ref/bsmap not built")

FUZZ = [
    CS.Case("fz_se100", opts=dict(s=16, v=5, I=4, S=21, u=1), maker=CS.mk_se(101, [200_000] * 3, 1500, 100, "cfg2", repeats=25)),
    CS.Case("fz_se_mixed", opts=dict(s=16, v=3, I=4, S=22, u=1, A=[CS.ADAPTER]), maker=CS.mk_se_mixed(102, [200_000] * 2, 1500)),
    CS.Case("fz_pe", opts=dict(s=16, v=3, I=4, S=23, u=1), paired=True, maker=CS.mk_pe(103, [200_000] * 2, 800, 100, 80, 470, repeats=20)),
    CS.Case("fz_rrbs", opts=dict(D="C-CGG", v=2, S=24, A=[CS.ADAPTER], u=1), maker=CS.mk_rrbs(104, [300_000] * 2, 1000, 75)),
]


@pytest.mark.parametrize("case", FUZZ, ids=lambda c: c.name)
def test_live(case):
    got = R.oracle_run(case)
    exp_main, exp_un, _ = R.reference_run(case)
    assert got["main"] == exp_main, R.first_diff(got["main"], exp_main)
    assert got["unpair"] == exp_un

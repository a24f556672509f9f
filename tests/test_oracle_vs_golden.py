"""The oracle restatement must reproduce, byte for byte, what the UNMODIFIED reference binary wrote
for every parity case (fixtures: tests/golden/*.gz, generator: tests/golden/make_golden.py)."""
import json
import os

import pytest

import cases as CS
import runners as R

MANIFEST = json.load(open(os.path.join(R.GOLDEN, "MANIFEST.json")))


@pytest.mark.parametrize("case", CS.CASES, ids=lambda c: c.name)
def test_inputs_are_stable(case):
    assert CS.input_digest(case) == MANIFEST[case.name]["inputs_sha256"], \
        "synthetic input generator drifted; regenerate tests/golden with make_golden.py"


@pytest.mark.parametrize("case", CS.CASES, ids=lambda c: c.name)
def test_oracle_matches_reference_output(case):
    got = R.oracle_run(case)
    exp_main, exp_un = R.golden_load(case)
    assert got["main"] == exp_main, R.first_diff(got["main"], exp_main)
    assert got["unpair"] == exp_un, R.first_diff(got["unpair"], exp_un)


def test_selection_order_pin():
    """POS 90002, 150003, 150003, 20001, 150003, 20001, all FLAG 256"""
    got = R.oracle_run(CS.BY_NAME["se_pin_order"])["main"].decode().splitlines()
    rows = [l.split("\t") for l in got if not l.startswith("@")]
    assert [r[1] for r in rows] == ["256"] * 6
    assert [int(r[3]) for r in rows] == [90002, 150003, 150003, 20001, 150003, 20001]


def test_palindrome_pin():
    """App. B Q8: dedupe ignores strand and chain -> unique hit"""
    got = R.oracle_run(CS.BY_NAME["se_pin_palindrome"])["main"].decode().splitlines()
    assert all(l.split("\t")[3] == "UM" for l in got) and len(got) == 2

"""Randomised differential rounds under pytest (-m gpu): tests/fuzz_parity.py's generator with fixed seeds, so the
driver's GPU test run exercises random -s/-I/-v/-w/-r/-n/-f/-L/-A/-D combinations, ragged and N-rich reads, batching and
slot strides against the CPU oracle every round.  The packed read input is run on the same rounds where it is exact."""
import numpy as np
import pytest

import fuzz_parity as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_fuzz_rounds_identical_to_the_oracle(seed):
    rng = np.random.default_rng(seed)
    for k in range(4):
        desc, bad = F.one_round(rng, k)
        assert bad is None, f"{desc} -> {bad}"

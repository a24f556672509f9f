"""Known-answer vectors produced by the reference's own object code pin the
oracle's primitives: 2-bit packing, Param::XT, the asymmetric mismatch table, the seed profile and
myrand."""
import pytest


def test_pack16(oracle):
    L = oracle.lib()
    assert L.bso_pack16(b"ACGTACGTACGTACGT") == 0x1B1B1B1B      # dbseq.cpp:73-76
    assert L.bso_pack16(b"acgtacgtacgtacgt") == 0x1B1B1B1B
    assert L.bso_pack16(b"NACGTNNNNNNNNNNN") == 0x06C00000      # N -> 0


@pytest.mark.parametrize("word,key", [
    (0x00000000, 0), (0x1B1B1B1B, 8609344), (0xFFFFFFFF, 21523360), (0x55555555, 21523360),
    (0xAAAAAAAA, 43046720), (0x06C00000, 2834352), (0x1B1B1B1B & 0x00FFFFFF, 106288)])
def test_xt(oracle, word, key):
    assert oracle.lib().bso_xt(word) == key                    # param.h:123


def test_mismatch_truth_table(oracle):
    """0 on the diagonal and for read T vs ref C; 1 for the other 11 cells (param.h:126,139)"""
    L = oracle.lib()
    for q in range(4):
        for s in range(4):
            exp = 0 if (q == s or (q == 3 and s == 1)) else 1
            assert L.bso_mismatch_cell(q, s) == exp, (q, s)


def test_profile(oracle):
    L = oracle.lib()
    for n, exp in enumerate([(0, 4, 4, 4), (16, 20, 20, 20), (32, 36, 36, 36)]):
        assert tuple(L.bso_profile_a(16, 4, n, i) for i in range(4)) == exp     # param.cpp:85-93
    for n, exp in enumerate([(0, 4, 4, 4), (12, 16, 16, 16), (24, 28, 28, 28)]):
        assert tuple(L.bso_profile_a(12, 4, n, i) for i in range(4)) == exp


@pytest.mark.parametrize("seed,exp", [
    (1, [3753797568, 1753423252, 4169258992, 827216640, 3535082705]),
    (7, [3328671682, 397857341, 3778289687, 1762583418, 1020533728]),
    (2147, [3008336154, 3138697394, 53770607, 1000866281, 756224921])])
def test_myrand(oracle, seed, exp):
    L = oracle.lib()
    assert [L.bso_myrand(i, seed) for i in (0, 1, 2, 3, 19999999)] == exp       # utilities.cpp:40-50

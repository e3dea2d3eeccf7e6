"""The oracle's std::sample / mt19937_64 restatement against this toolchain's libstdc++
(the "reference itself run here" for that piece) and the known answers of SURVEY App. C."""
import numpy as np
import pytest

import oracle_lib as O

KNOWN = {
    (1234567, 10): [[3, 6, 7, 8], [0, 4, 6, 7], [0, 1, 2, 4]],
    (1234567, 500): [[256, 263, 364, 454], [27, 133, 451, 486], [61, 75, 185, 488], [9, 381, 427, 492]],
    (123, 500): [[53, 56, 216, 318], [196, 201, 370, 481], [125, 163, 413, 467], [52, 211, 223, 332]],
    (123, 130): [[35, 49, 56, 82], [22, 47, 85, 128], [2, 37, 104, 105], [81, 89, 112, 125]],
}


@pytest.mark.parametrize("key", sorted(KNOWN))
def test_known_answers(key):
    seed, n = key
    exp = np.array(KNOWN[key])
    assert np.array_equal(O.sample_stream(seed, n, len(exp)), exp)


@pytest.mark.parametrize("seed,n", [(1234567, 500), (42, 54), (7, 4), (99, 5), (5, 77), (2 ** 63 + 11, 1000), (0, 88)])
def test_matches_real_std_sample(seed, n):
    a = O.sample_stream(seed, n, 300)
    b = O.sample_stream(seed, n, 300, real=True)
    assert np.array_equal(a, b)
    assert (np.diff(a, axis=1) > 0).all()  # selection sampling returns ascending indices

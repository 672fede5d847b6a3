"""The device algorithms, restated sequentially in tests/algo_model.py, against the oracle."""
import random

import pytest

import algo_model as am
import fqgen


def _oracle_chain(oracle, data, sentinel, goff):
    blob = (b'\n' if sentinel else b'') + data
    table, st, tail, resume = oracle.parse_chain(blob, 0, goff)
    return table.tolist(), st, tail.tolist(), resume


@pytest.mark.parametrize('seed', range(8))
def test_general_model_matches_oracle(oracle, seed):
    for data in fqgen.corpus(1000 + seed, 400):
        for sentinel in (1, 0):
            want = _oracle_chain(oracle, data, sentinel, -1)
            for chunk in (1, 3, 8, 1024):
                got = am.model_general(data, sentinel, -1, chunk=chunk)
                assert (got[0], got[1], list(got[2]), got[3]) == want, (data, sentinel, chunk)


@pytest.mark.parametrize('seed', range(8))
def test_fast4_model_exact_or_declines(oracle, seed):
    accepted = 0
    for data in fqgen.corpus(2000 + seed, 400):
        for sentinel in (1, 0):
            want = _oracle_chain(oracle, data, sentinel, 7)
            for tile in (5, 16, 64, 4096):
                got = am.model_fast4(data, sentinel, 7, tile)
                if got is None:
                    continue
                accepted += 1
                assert (got[0], got[1], list(got[2]), got[3]) == want, (data, sentinel, tile)
    assert accepted > 100


def test_fast4_model_accepts_clean_four_line(oracle):
    rng = random.Random(5)
    for _ in range(200):
        data = fqgen.fastq_bytes(rng, rng.randint(1, 10), long_plus=0.3, trailing_newlines=rng.randint(0, 3),
                                 at_plus_bias=0.5)
        for tile in (7, 64, 4096):
            got = am.model_fast4(data, 1, -1, tile)
            assert got is not None, data
            assert (got[0], got[1], list(got[2]), got[3]) == _oracle_chain(oracle, data, 1, -1)

"""The CPU oracle (oracle/pmnet_oracle.c) against outputs of the real reference (tests/golden, made by
oracle/make_golden.py with GraphMatcher.run of /root/reference). This is what pins the oracle."""

import numpy as np
import pytest
from golden_util import CASES, FALLBACK_CASES, load_case, load_fallback, rel_err

import oracle as orc


def test_cases_present():
    assert len(CASES) >= 9


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference(name):
    c = load_case(name)
    out = orc.score(c["model"], c["batch"], c["weights"], threads=0)
    # observed: bit-identical fp64 results; the bound leaves room for libm differences between hosts
    assert rel_err(out["scores"], c["ref"]).max() <= 1e-9
    # reference returns exactly 0 for ligands without any candidate (graph_match.py:95-99)
    assert np.array_equal(out["scores"] == 0.0, c["ref"] == 0.0)
    empty = out["status"] == 1
    assert np.all(c["ref"][empty] == 0.0)


def test_oracle_thread_count_does_not_change_results():
    c = load_case("syn0_c8")
    a = orc.score(c["model"], c["batch"], c["weights"], threads=1)
    b = orc.score(c["model"], c["batch"], c["weights"], threads=4)
    assert np.array_equal(a["scores"], b["scores"])
    assert np.array_equal(a["stats"], b["stats"])


def test_oracle_subrange():
    c = load_case("syn0_c8")
    full = orc.score(c["model"], c["batch"], c["weights"], threads=2)
    part = orc.score(c["model"], c["batch"], c["weights"], threads=2, begin=10, end=50)
    assert np.array_equal(full["scores"][10:50], part["scores"])


@pytest.mark.parametrize("name", FALLBACK_CASES)
def test_oracle_matches_reference_numpy_fallback(name):
    """The reference's second scorer (match_utils.py:9-122, all fp32; selected at graph_match.py:12-15 when numba is
    missing) on the same ligands: same discrete decisions, scores equal up to its fp32 accumulation (observed 3e-8)."""
    c = load_case(name)
    fb = load_fallback(name)
    out = orc.score(c["model"], c["batch"], c["weights"], threads=0)
    assert rel_err(out["scores"], fb).max() <= 1e-6
    assert np.array_equal(out["scores"] == 0.0, fb == 0.0)
    assert rel_err(c["ref"], fb).max() <= 1e-6  # the two reference variants agree with each other

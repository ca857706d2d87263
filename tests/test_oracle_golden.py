"""The CPU oracle (oracle/pyramid_oracle.{c,py}) against the golden vectors produced by the live
reference (tests/golden/make_golden.py).  CPU only."""
import pytest

import golden_util


@pytest.mark.parametrize("name", golden_util.case_names())
def test_oracle_matches_golden(oracle, name):
    frames, kwargs, expect = golden_util.load_case(name)
    out, info = oracle.focus_stack(frames, **kwargs)
    golden_util.assert_matches(out, info, expect, name)

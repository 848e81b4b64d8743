"""Host logic of the engine pool that needs no GPU: how a large call is cut into pieces for the lanes (the same rule the
library's run_lanes uses: csrc/jg_abi.cu)."""
from hypothesis import given, settings, strategies as st

from jogramop_framework_b200.engine import EnginePool


@settings(max_examples=300, deadline=None)
@given(n=st.integers(0, 5_000_000), unit=st.integers(1, 1 << 20))
def test_pieces_cover_the_call_in_aligned_equal_parts(n, unit):
    pieces = EnginePool._pieces(None, n, unit)
    cap = max(32, unit // 32 * 32)
    assert sum(m for _, m in pieces) == n
    off = 0
    for o, m in pieces:
        assert o == off and o % 32 == 0 and 0 < m <= cap         # contiguous, packed words aligned, fits one pass
        off += m
    if n:
        sizes = [m for _, m in pieces[:-1]]
        assert len(set(sizes)) <= 1                              # equal pieces, the last one takes the remainder
        assert len(pieces) <= -(-n // cap) + 1


def test_a_call_that_fits_one_pass_is_not_cut():
    assert EnginePool._pieces(None, 1_000_000, 1 << 20) == [(0, 1_000_000)]
    assert EnginePool._pieces(None, 1 << 20, 1 << 20) == [(0, 1 << 20)]
    assert len(EnginePool._pieces(None, (1 << 20) + 1, 1 << 20)) == 2
    assert EnginePool._pieces(None, 0, 1 << 20) == []

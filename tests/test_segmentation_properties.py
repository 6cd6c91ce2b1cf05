"""Property tests (hypothesis) for the integer work of the path: syg_segment_table / syg_frame_count must reproduce the reference's
arithmetic bit for bit (segmentation.py:62-114: int() truncations, hop = max(1, int(seg * (1 - overlap))), pad / min-length rules;
manager.py:149-157 frame counts) for arbitrary sizes, incl. the floating-point corner cases of `seg * (1 - overlap)`."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import sygnals_oracle as orc
from sygnals_b200 import _ffi

lib = _ffi.library()


@settings(max_examples=300, deadline=None)
@given(total=st.integers(0, 200000), sr=st.sampled_from([1000, 8000, 16000, 22050, 44100, 48000]),
       sec=st.floats(0.001, 3.0, allow_nan=False), ovl=st.one_of(st.sampled_from([0.0, 0.25, 0.5, 0.75, 0.9, 0.99]), st.floats(0.0, 0.999)),
       pad=st.booleans(), mn=st.one_of(st.none(), st.floats(0.0, 1.0)))
def test_segment_table_matches_reference_arithmetic(total, sr, sec, ovl, pad, mn):
    seg, hop, table = orc.segment_table(total, sr, sec, ovl, pad, mn)
    seg_len, seg_hop, starts, valid = lib.segment_table(total, sr, sec, ovl, pad, mn)
    assert (seg_len, seg_hop) == (seg, hop) or (seg == 0 and seg_len == 0)
    assert [(int(s), int(v)) for s, v in zip(starts, valid)] == [(int(s), int(v)) for s, v in table]


@settings(max_examples=300, deadline=None)
@given(n=st.integers(0, 10 ** 7), log2=st.integers(5, 13), hop=st.integers(1, 5000), center=st.booleans())
def test_frame_count_matches_librosa_rule(n, log2, hop, center):
    fl = 1 << log2
    padded = n + 2 * (fl // 2) if center else n
    want = 1 + (padded - fl) // hop if padded >= fl else 0
    assert lib.frame_count(n, fl, hop, center) == want
    if center:
        assert want == 1 + n // hop                                   # manager.py:149-157 for even frame lengths

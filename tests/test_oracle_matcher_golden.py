"""The matcher restatement (oracle/orb_oracle.c) against golden vectors of the reference's own ORBMatcher.cpp compiled verbatim
(tests/golden/matcher_ref.npz, tools/gen_golden_matcher.py).  Unlike tests/test_oracle_matcher_ref.py this needs no reference
sources at run time, so the pin holds on any machine."""
from oracle import orb_oracle as orc
from golden_matcher import replay


def test_restatement_reproduces_the_reference_matcher_golden():
    orc.build()
    impl = {k: getattr(orc, k) for k in ("search_for_initialization", "search_by_projection", "search_local_points", "search_for_triangulation",
                                         "search_by_bow", "search_fuse")}
    assert replay(impl) == 13

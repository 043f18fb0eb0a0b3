"""Property tests (hypothesis) of the host-side logic that surrounds the device path: the block partition every
sharded call relies on, label enumeration, and the trial-list parsers (scoring/scorePLDA.py:40-73 formats)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from plda_b200 import pipeline
from plda_b200.dist import block_bounds


@settings(max_examples=200, deadline=None)
@given(n=st.integers(0, 10**7), world=st.integers(1, 16))
def test_block_bounds_is_a_balanced_contiguous_partition(n, world):
    spans = [block_bounds(n, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    assert all(s >= 0 for s in sizes) and max(sizes) - min(sizes) <= 1
    assert sizes == sorted(sizes, reverse=True)            # the larger blocks come first (PeerShardedScorer bounds)


_name = st.text(alphabet="abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789_", min_size=1, max_size=8)


@settings(max_examples=100, deadline=None)
@given(labels=st.lists(_name, min_size=1, max_size=40))
def test_enumerate_labels_is_np_unique_order(labels):
    table, ids = pipeline.enumerate_labels(labels)
    uniq = sorted(set(labels))
    assert table == {s: i for i, s in enumerate(np.unique(np.asarray(labels)))}
    assert len(table) == len(uniq) and ids.dtype.kind == "u"
    inv = {v: k for k, v in table.items()}
    assert [inv[int(i)] for i in ids] == list(labels)


@settings(max_examples=100, deadline=None)
@given(trials=st.lists(st.tuples(_name, _name, _name), min_size=1, max_size=30))
def test_test_ref_round_trip(tmp_path_factory, trials):
    """Lines "<target> <enrol>-<utt>" -> {target: [[utt, enrol], ...]} in file order (scorePLDA.py:45-49); the
    utterance id may itself contain '-' (everything after the first one)."""
    p = tmp_path_factory.mktemp("ref") / "trials.txt"
    p.write_text("".join("%s %s-%s-x\n" % (t, e, u) for t, e, u in trials))
    parsed = pipeline.parse_test_ref(str(p))
    want = {}
    for t, e, u in trials:
        want.setdefault(t, []).append([u + "-x", e])
    assert dict(parsed) == want

"""seqalign_multi_*: one batch over several engines from one process (host/sa_multi.c).  In the build
container both engines sit on the one emulated device; on the GPU box the same tests spread over every
visible B200 (and still pass with one)."""
import numpy as np
import pytest

import seqalign
from seqalign import NW, SW, MODE_SCORE, MODE_SCORE_ONLY, MODE_ALIGN, MODE_HITS
from helpers import SPECS, orc_batch_nw, orc_batch_sw, orc_from_scoring, orc_nw, orc_sw_hits, ragged_batch, scoring_from_spec

pytestmark = pytest.mark.parity


@pytest.fixture(scope="module")
def multi(backend):
    if backend == "gpu":
        n = max(2, seqalign.device_count())
        devs = [i % seqalign.device_count() for i in range(n)]
    else:
        devs = [0, 0, 0]
    m = seqalign.MultiAligner(devs)
    yield m
    m.close()


def test_multi_scores_ragged(multi, big):
    n, maxlen = (3000, 200) if big else (23, 40)
    sa, sb = ragged_batch(77, n, maxlen, maxlen)
    a, oa = seqalign.pack(sa)
    b, ob = seqalign.pack(sb)
    for algo, spec in ((SW, "sw_cli"), (NW, "nw_default")):
        sc = scoring_from_spec(SPECS[spec])
        multi.set_scoring(sc)
        multi.submit_packed(algo, MODE_SCORE, a, oa, b, ob)
        s, x, y = multi.ends()
        o = orc_from_scoring(sc)
        if algo == SW:
            es, ex, ey = orc_batch_sw(o, a, oa, b, ob)
            assert np.array_equal(s, es) and np.array_equal(x, ex) and np.array_equal(y, ey)
        else:
            assert np.array_equal(s, orc_batch_nw(o, a, oa, b, ob))
            assert np.array_equal(x, np.diff(oa)) and np.array_equal(y, np.diff(ob))
        multi.submit_packed(algo, MODE_SCORE_ONLY, a, oa, b, ob)
        s2, x2, y2 = multi.ends()
        assert np.array_equal(s2, s)
        assert np.array_equal(x2, np.diff(oa) if algo == NW else 0 * x2)
    # every engine got a share, and the shares are contiguous and in order
    where = [multi.where(i) for i in range(n)]
    firsts = [i for i in range(n) if where[i][1] == 0]
    assert len(firsts) == multi.devices and firsts[0] == 0


def test_multi_uniform_and_alignments(multi, big):
    from seqalign.synth import synth_batch
    n = 2000 if big else 11
    la, lb = (150, 150) if big else (31, 28)
    a, oa, b, ob = synth_batch(9, 5, n, la, lb)
    sc = scoring_from_spec(SPECS["sw_cli"])
    multi.set_scoring(sc)
    multi.submit_uniform(SW, MODE_SCORE_ONLY, a, la, b, lb, n)
    es, _, _ = orc_batch_sw(orc_from_scoring(sc), a, oa, b, ob)
    assert np.array_equal(multi.scores(), es)
    nwsc = scoring_from_spec(SPECS["nw_default"])
    multi.set_scoring(nwsc)
    m = min(n, 40)
    multi.submit_uniform(NW, MODE_ALIGN, a, la, b, lb, m)
    o = orc_from_scoring(nwsc)
    for i in list(range(0, m, 3)) + [m - 1]:
        al = multi.alignment(i)
        rc, score, ra, rb = orc_nw(o, a[i * la:(i + 1) * la].tobytes(), b[i * lb:(i + 1) * lb].tobytes())
        assert (al.result_a, al.result_b, al.score) == (ra, rb, score)
    assert np.array_equal(multi.scores()[:m], orc_batch_nw(o, a[:m * la], oa[:m + 1], b[:m * lb], ob[:m + 1]))


def test_multi_hits_and_small_batches(multi):
    sc = scoring_from_spec(SPECS["sw_cli"])
    multi.set_scoring(sc)
    multi.set_hit_limits(4, 1)
    sa, sb = ragged_batch(5, 7, 30, 30, min_len=4)
    a, oa = seqalign.pack(sa)
    b, ob = seqalign.pack(sb)
    multi.submit_packed(SW, MODE_HITS, a, oa, b, ob)
    o = orc_from_scoring(sc)
    for i in range(7):
        n, want = orc_sw_hits(o, sa[i], sb[i], 4)
        got = multi.hits(i)
        assert [(h.result_a, h.result_b, h.score, h.pos_a, h.pos_b) for h in got] == \
               [(h["result_a"], h["result_b"], h["score"], h["pos_a"], h["pos_b"]) for h in want]
    # fewer pairs than engines, and none at all
    multi.submit_packed(SW, MODE_SCORE, a[:oa[1]], oa[:2], b[:ob[1]], ob[:2])
    assert multi.scores().tolist() == orc_batch_sw(o, a, oa[:2], b, ob[:2])[0].tolist()
    multi.submit_packed(SW, MODE_SCORE, a[:0], oa[:1], b[:0], ob[:1])
    assert len(multi.scores()) == 0

"""Trial-list pipeline (plda_b200/pipeline.py, port of scoring/scorePLDA.py): parsers on CPU; the scoring flow on GPU
against the per-trial loop the reference runs (scorePLDA.py:302-318) driven through the oracle."""
import numpy as np
import pytest

from plda_b200 import pipeline


def test_parse_test_ref(tmp_path):
    p = tmp_path / "trials.txt"
    p.write_text("F001 F001-F001_s1_utt1 extra\nF001 F002-F002_s2-utt-3\n\nM007 M007-M007_x\n")
    t = pipeline.parse_test_ref(str(p))
    assert t["F001"] == [["F001_s1_utt1", "F001"], ["F002_s2-utt-3", "F002"]]       # scorePLDA.py:45-49
    assert t["M007"] == [["M007_x", "M007"]]


def test_parse_mlf(tmp_path):
    p = tmp_path / "test.mlf"
    p.write_text('#!MLF!#\n"*/F001-F001_s1_utt1.lab"\nF001\n.\n"*/F002-F002_s2-utt-3.lab"\nF001\n.\n')
    t = pipeline.parse_mlf(str(p))
    assert t["F001"] == [["F001_s1_utt1", "F001"], ["F002_s2-utt-3", "F002"]]       # scorePLDA.py:61-72


def test_enumerate_labels_matches_reference_order():
    tab, ids = pipeline.enumerate_labels(["spk_b", "spk_a", "spk_b", "spk_c"])
    assert tab == {"spk_a": 0, "spk_b": 1, "spk_c": 2} and ids.dtype.kind == "u"   # np.unique order, :243-246
    assert list(ids) == [1, 0, 1, 2]


@pytest.mark.gpu
def test_pipeline_equals_per_trial_loop():
    from oracle import kaldi_plda as kp
    from plda_b200 import PLDA
    d = 24
    a_b = kp.two_cov_generator(d, seed=1234)
    xb, lb, _ = kp.synth_speakers(a_b, [6] * 30, seed=1234)
    xe, le, z = kp.synth_speakers(a_b, [3] * 8, seed=1235)
    rng = np.random.RandomState(1236)
    xt = 0.5 + z @ a_b.T + rng.randn(8, d)
    zn, _, _ = kp.synth_speakers(a_b, [1] * 20, seed=1237)
    bkg_labels = ["b%03d" % i for i in lb]
    enrol_labels = ["m%02d" % i for i in le]
    test_labels = ["m%02d_utt" % i for i in range(8)]
    trials = {"m%02d" % e: [["m%02d_utt" % t, "m%02d" % t] for t in range(8)] for e in range(8)}
    trials["ghost"] = [["m00_utt", "m00"]]                                  # unknown model -> counted as error
    trials["m01"].append(["nope", "m01"])                                   # unknown utterance -> error
    g = PLDA()
    lines, errors = pipeline.score_trials(g, xb, bkg_labels, xe, enrol_labels, xt, test_labels, trials, iters=4,
                                          znorm_vectors=zn)
    assert errors == 2 and len(lines) == 64
    # the reference flow, per trial, through the oracle
    ref = kp.MPlda()
    _, bi = pipeline.enumerate_labels(bkg_labels)
    etab, ei = pipeline.enumerate_labels(enrol_labels)
    ttab, ti = pipeline.enumerate_labels(test_labels)
    ref.fit(xb, bi, 4)
    te, tt = ref.transform(xe, ei), ref.transform(xt, ti)
    ref.norm(zn, te)
    for line in lines:
        model, rest, score = line.split()
        target, utt = rest.split("-", 1)
        want = ref.score(etab[model], te[etab[model]], tt[ttab[utt]])
        assert abs(float(score) - want) <= 2e-3 * max(1.0, abs(want)) + 5e-4     # "{:.3f}" rounding


@pytest.mark.gpu
def test_lda_pipeline_equals_per_trial_loop():
    """scoreLDA.main (scoring/scoreLDA.py:212-246) per trial through the LDA oracle == one batched device call."""
    from oracle.lda_port import LDAOracle
    from plda_b200 import LDA
    rng = np.random.RandomState(4)
    k, d, per = 7, 16, 30
    centers = rng.randn(k, d) * 1.5
    names = ["spk%c" % c for c in "gcafbed"]                                # not in sorted order on purpose
    labels = [names[i % k] for i in range(k * per)]
    x = np.stack([centers[i % k] for i in range(k * per)]) + rng.randn(k * per, d)
    tests = {"utt%02d" % i: centers[i % k] + rng.randn(d) for i in range(12)}
    trials = {n: [["utt%02d" % i, names[i % k]] for i in range(12)] for n in names}
    trials["ghost"] = [["utt00", names[0]]]
    trials[names[0]].append(["missing", names[0]])
    lines, errors = pipeline.score_trials_lda(LDA(solver="svd", precision="fp64"), x, labels, tests, trials)
    assert errors == 2 and len(lines) == k * 12
    ref = LDAOracle(solver="svd")
    uniq = list(np.unique(labels))
    ref.fit(x, np.array([uniq.index(l) for l in labels]))
    for line in lines:
        model, rest, score = line.split()
        target, utt = rest.split("-", 1)
        want = ref.predict_log_proba(tests[utt][np.newaxis, :])[0][uniq.index(model)]      # scoreLDA.py:239-243
        assert abs(float(score) - want) <= 1e-3 * max(1.0, abs(want)) + 5e-4                # "{:.3f}" rounding


# ---- host logic on CPU: the same flows driven with the ORACLE objects standing in for the device classes ---------
class _OraclePldaAdapter:
    """oracle/kaldi_plda.MPlda behind the method set pipeline.score_trials uses (fit / transform / norm / score_grid)."""

    def __init__(self):
        from oracle import kaldi_plda as kp
        self.kp, self.m = kp, kp.MPlda()

    def fit(self, x, y, iters=10):
        return self.m.fit(x, y, iters)

    def transform(self, x, y):
        return self.m.transform(x, y)

    def norm(self, vectors, transformedvecs, numutts=0):
        return self.m.norm(vectors, transformedvecs, numutts)

    def score_grid(self, enrol, counts, test, enrol_ids=None):
        g = self.kp.score_grid(self.m.plda, enrol, np.asarray(counts), test)
        if enrol_ids is not None:
            mean = np.array([self.m.meanz[int(k)] for k in enrol_ids])
            std = np.array([self.m.stdvz[int(k)] for k in enrol_ids])
            g = (g - mean[:, None]) / std[:, None]
        return g

    def score_trials(self, enrol, counts, test, trial_enrol, trial_test, enrol_ids=None):
        """Listed trials (plda_score_trials): here simply a gather of the oracle's grid."""
        g = self.score_grid(enrol, counts, test, enrol_ids)
        return g[np.asarray(trial_enrol), np.asarray(trial_test)]


def test_plda_pipeline_host_logic_with_oracle():
    """Label enumeration, transform / norm plumbing, grid -> trial gather and the error count of score_trials,
    against the per-trial loop of scorePLDA.py:302-318 (oracle on both sides, so this runs without a GPU)."""
    from oracle import kaldi_plda as kp
    d = 12
    a_b = kp.two_cov_generator(d, seed=1234)
    xb, lb, _ = kp.synth_speakers(a_b, [5] * 12, seed=1234)
    xe, le, z = kp.synth_speakers(a_b, [2] * 5, seed=1235)
    rng = np.random.RandomState(1236)
    xt = 0.5 + z @ a_b.T + rng.randn(5, d)
    zn, _, _ = kp.synth_speakers(a_b, [1] * 9, seed=1237)
    enrol_labels = ["m%d" % i for i in le]
    test_labels = ["m%d_utt" % i for i in range(5)]
    trials = {"m%d" % e: [["m%d_utt" % t, "m%d" % t] for t in range(5)] for e in range(5)}
    trials["ghost"] = [["m0_utt", "m0"]]
    adapter = _OraclePldaAdapter()
    lines, errors = pipeline.score_trials(adapter, xb, ["b%02d" % i for i in lb], xe, enrol_labels, xt, test_labels,
                                          trials, iters=3, znorm_vectors=zn)
    assert errors == 1 and len(lines) == 25
    etab, _ = pipeline.enumerate_labels(enrol_labels)
    ttab, _ = pipeline.enumerate_labels(test_labels)
    te, tt = adapter.m.transform(xe, pipeline.enumerate_labels(enrol_labels)[1]), None
    tt = adapter.m.transform(xt, pipeline.enumerate_labels(test_labels)[1])
    for line in lines:
        model, rest, score = line.split()
        target, utt = rest.split("-", 1)
        want = adapter.m.score(etab[model], te[etab[model]], tt[ttab[utt]])
        assert abs(float(score) - want) <= 5.1e-4 + 1e-6 * abs(want)            # "{:.3f}" rounding (+ float32 score)


def test_lda_pipeline_host_logic_with_oracle():
    """score_trials_lda (scoreLDA.py:212-246) with the LDA oracle as the model object."""
    from oracle.lda_port import LDAOracle
    rng = np.random.RandomState(4)
    k, d, per = 4, 6, 25
    centers = rng.randn(k, d) * 2
    names = ["zed", "amy", "kim", "bob"]
    labels = [names[i % k] for i in range(k * per)]
    x = np.stack([centers[i % k] for i in range(k * per)]) + rng.randn(k * per, d)
    tests = {"u%d" % i: centers[i % k] + rng.randn(d) for i in range(6)}
    trials = {n: [["u%d" % i, names[i % k]] for i in range(6)] for n in names}
    trials["amy"].append(["missing", "amy"])
    lines, errors = pipeline.score_trials_lda(LDAOracle(solver="svd"), x, labels, tests, trials)
    assert errors == 1 and len(lines) == 24
    uniq = list(np.unique(labels))
    ref = LDAOracle(solver="svd")
    ref.fit(x, np.array([uniq.index(l) for l in labels]))
    for line in lines:
        model, rest, score = line.split()
        _, utt = rest.split("-", 1)
        want = ref.predict_log_proba(tests[utt][np.newaxis, :])[0][uniq.index(model)]
        assert abs(float(score) - want) <= 5.1e-4

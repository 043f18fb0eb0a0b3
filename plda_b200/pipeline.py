"""Trial-list scoring pipelines -- py3 ports of the CALLERS of the hot path, ``scoring/scorePLDA.py`` and
``scoring/scoreLDA.py``.

The reference parses a trial list, enumerates string labels to ``uint`` (``scorePLDA.py:243-255``), runs
``fit -> transform x2 -> [norm]`` (``:258-298``) and then scores every listed trial with one ``plda.score`` call in a
Python double loop (``:302-318``).  Here the same flow ends in ONE ``score_grid`` on the device followed by a gather of
the listed trials; the output lines keep the reference's format ``"{model} {target}-{utt} {score:.3f}\\n"`` (``:317-318``).
File formats beyond the two trial-list formats (HTK features, marshalled d-vector dumps) stay out of scope.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np


def parse_test_ref(path: str) -> Dict[str, List[List[str]]]:
    """``test_ref`` (``scoring/scorePLDA.py:40-50``): lines ``<targetmodel> <enrolmodel>-<testutt> ...``."""
    tests = defaultdict(list)
    with open(path, "r") as fh:
        for line in fh:
            line = line.rstrip("\n")
            if not line.strip():
                continue
            targetmdl, enrol_testutt = line.split()[:2]
            parts = enrol_testutt.split("-")
            tests[targetmdl].append(["-".join(parts[1:]), parts[0]])
    return tests


def parse_mlf(path: str) -> Dict[str, List[List[str]]]:
    """``mlffile`` (``scoring/scorePLDA.py:55-73``): HTK master label file, ``"*/<enrol>-<utt>.lab"`` then the target."""
    tests = defaultdict(list)
    with open(path, "r") as fh:
        next(fh)                                   # "#!MLF!#"
        for line in fh:
            line = line.rstrip("\n")
            if line.startswith('"'):
                without_lab = line.split(".")[0]
                without_slashes = without_lab.split("/")[1]
                parts = without_slashes.split("-")
                targetmdl = next(fh).rstrip("\n")
                tests[targetmdl].append(["-".join(parts[1:]), parts[0]])
    return tests


def enumerate_labels(labels: Sequence[str]) -> Tuple[Dict[str, int], np.ndarray]:
    """String labels -> dense uint ids in ``np.unique`` order (``scorePLDA.py:243-255``)."""
    uniq = np.unique(np.asarray(labels))
    table = {str(s): i for i, s in enumerate(uniq)}
    return table, np.array([table[str(s)] for s in labels], dtype="uint")


def score_trials(plda, bkg_vectors, bkg_labels: Sequence[str], enrol_vectors, enrol_labels: Sequence[str],
                 test_vectors, test_labels: Sequence[str], trials: Dict[str, Iterable[Sequence[str]]],
                 iters: int = 10, znorm_vectors=None, zutt: int = 0, fit: bool = True) -> Tuple[List[str], int]:
    """The body of ``scorePLDA.main`` (``:258-318``).  Returns (output lines, number of skipped trials).

    ``trials``: ``{enrolmodel: [[testutt, targetmodel], ...]}`` as produced by ``parse_test_ref`` / ``parse_mlf``.
    """
    enrol_tab, enrol_ids = enumerate_labels(enrol_labels)
    test_tab, test_ids = enumerate_labels(test_labels)
    if fit:
        _, bkg_ids = enumerate_labels(bkg_labels)
        plda.fit(np.asarray(bkg_vectors), bkg_ids, iters)
    enrol_t = plda.transform(np.asarray(enrol_vectors), enrol_ids)
    test_t = plda.transform(np.asarray(test_vectors), test_ids)
    if znorm_vectors is not None:
        plda.norm(np.asarray(znorm_vectors), enrol_t, zutt)
    ek, tk = sorted(enrol_t), sorted(test_t)
    e = np.stack([enrol_t[k][1] for k in ek])
    n = np.array([enrol_t[k][0] for k in ek], dtype=np.int32)
    t = np.stack([test_t[k][1] for k in tk])
    # the reference scores test vectors as transformed with their own utterance count; LLR uses n of the enrol side.
    # Only the LISTED trials are scored and brought back (plda_score_trials): the index arrays go to the device, the
    # scores of the list come back -- not the Ne x Nt grid the reference's per-trial loop would walk.
    e_pos = {k: i for i, k in enumerate(ek)}
    t_pos = {k: i for i, k in enumerate(tk)}
    wanted, te, tt, errors = [], [], [], 0
    for enrolmodel, vals in trials.items():
        if enrolmodel not in enrol_tab:
            errors += 1
            continue
        ei = e_pos[enrol_tab[enrolmodel]]
        for testutt, targetmdl in vals:
            if testutt not in test_tab:
                errors += 1
                continue
            wanted.append((enrolmodel, targetmdl, testutt))
            te.append(ei)
            tt.append(t_pos[test_tab[testutt]])
    if not wanted:
        return [], errors
    scores = plda.score_trials(e, n, t, np.asarray(te, dtype=np.int32), np.asarray(tt, dtype=np.int32),
                               enrol_ids=np.array(ek, dtype=np.uint64) if znorm_vectors is not None else None)
    lines = ["{} {}-{} {:.3f}\n".format(m, tg, u, float(sc)) for (m, tg, u), sc in zip(wanted, scores)]
    return lines, errors


def score_trials_lda(lda, dvectors, labels: Sequence[str], test_vectors: Dict[str, np.ndarray],
                     trials: Dict[str, Iterable[Sequence[str]]], fit: bool = True) -> Tuple[List[str], int]:
    """The body of ``scoreLDA.main`` (``scoring/scoreLDA.py:212-246``): enumerate the speaker labels in ``np.unique``
    order (``:212-214``), ``lda.fit`` (``:223``), then for every listed trial the log-probability of the ENROL model's
    class for the test utterance (``:237-243``).  The reference calls ``predict_log_proba`` once per trial on a
    1 x d matrix; here every distinct test utterance goes through ONE device call and the trials are gathered.

    ``test_vectors``: ``{testutt: d-vector}`` (``testtofeature``, ``:205``); ``trials`` as produced by
    ``parse_test_ref`` / ``parse_mlf``.  Returns (output lines ``"{model} {target}-{utt} {score:.3f}\n"``, errors).
    """
    uniq = np.unique(np.asarray(labels))
    spktonum = {str(s): i for i, s in enumerate(uniq)}
    if fit:
        lda.fit(np.asarray(dvectors), np.array([spktonum[str(s)] for s in labels]))
    wanted, errors = [], 0
    for enrolmodel, vals in trials.items():
        if enrolmodel not in spktonum:
            errors += 1
            continue
        for testutt, targetmdl in vals:
            if testutt not in test_vectors:
                errors += 1
                continue
            wanted.append((enrolmodel, targetmdl, testutt))
    utts = sorted({w[2] for w in wanted})
    lines: List[str] = []
    if not utts:
        return lines, errors
    pos = {u: i for i, u in enumerate(utts)}
    x = np.stack([np.asarray(test_vectors[u], dtype=np.float64).reshape(-1) for u in utts])
    logp = np.asarray(lda.predict_log_proba(x))
    for enrolmodel, targetmdl, testutt in wanted:
        lines.append("{} {}-{} {:.3f}\n".format(enrolmodel, targetmdl, testutt,
                                                float(logp[pos[testutt], spktonum[enrolmodel]])))
    return lines, errors

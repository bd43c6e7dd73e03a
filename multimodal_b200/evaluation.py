"""Drop-in for the nearest-example evaluation that follows the hot path in the reference
(multimodal/evaluation.py:58-130): same function names and results; `all_distances` / `classify_NN` /
`evaluate_NN_label` run on the GPU (csrc/evaluation.cu -> klnmf_pairwise_host) for the measures of
`multimodal_b200.lib.metrics` instead of broadcasting an n_test x n_ex x d temporary on the host.

There is no CPU path: an unknown `measure` callable raises TypeError (the reference would call it on the
broadcast arrays)."""
import numpy as np
import scipy.sparse as sp

from . import _native
from .lib import metrics as _metrics


def todense(X):
    """evaluation.py:97-101."""
    return np.asarray(X.todense()) if sp.issparse(X) else X


def _measure_key(measure):
    try:
        return _metrics.DEVICE_MEASURES[measure]
    except (KeyError, TypeError):
        name = getattr(measure, "__name__", None)
        if name in _native.MEASURES and getattr(measure, "__module__", "").endswith("metrics"):
            return name          # the reference's own multimodal.lib.metrics functions are accepted by name
        raise TypeError("measure %r has no device implementation (supported: %s)"
                        % (measure, ", ".join(sorted(_native.MEASURES))))


def all_distances(reco_data, ex_data, measure, device=0):
    """evaluation.py:103-106: dists[i, j] = measure(reco_data[i], ex_data[j])."""
    D, _ = _native.pairwise(todense(reco_data), todense(ex_data), _measure_key(measure), True, False, device)
    return D


def dists_to_found_labels(dists, ex_labels):
    """evaluation.py:73-76."""
    return [ex_labels[m] for m in np.argmin(dists, axis=1)]


def classify_NN(reco_data, ex_data, ex_labels, measure, device=0):
    """evaluation.py:109-116; the argmin is taken on the device, the distance matrix is never built."""
    _, idx = _native.pairwise(todense(reco_data), todense(ex_data), _measure_key(measure), False, True, device)
    return [ex_labels[m] for m in idx]


def found_labels_to_score(true, found):
    """evaluation.py:79-82."""
    return np.average([f == l for f, l in zip(found, true)])


def found_labels_to_confusion(true, found, n_labels):
    """evaluation.py:85-91 (fancy-index += : a repeated (true, found) pair counts once, as in the reference)."""
    conf = np.zeros((n_labels, n_labels))
    conf[true, found] += 1
    return conf


def scores_from_dists(dists, true_labels_0, true_labels_1=None, verbose=False):
    """evaluation.py:58-70 (deprecated there, still used by evaluate_NN_label)."""
    if true_labels_1 is None:
        assert dists.shape[0] == dists.shape[1]
        true_labels_1 = true_labels_0
    found = dists_to_found_labels(dists, true_labels_1)
    result = found_labels_to_score(true_labels_0, found)
    if verbose:
        print(result)
    return result


def evaluate_NN_label(reco_data, test_data, true_labels, test_labels, measure, device=0):
    """evaluation.py:119-130."""
    found = classify_NN(reco_data, test_data, test_labels, measure, device)
    return found_labels_to_score(true_labels, found)


def evaluate_label_reco(reco_acti, true_labels):
    """evaluation.py:47-55."""
    labels = np.asarray(true_labels)
    best = reco_acti.argmax(axis=1)
    assert best.shape == labels.shape
    return np.average(best == labels)


# ---- host-side label bookkeeping of the reference (evaluation.py:8-44): list / small-array logic, no arithmetic
def chose_examples(labels, label_set=None, number=1):
    """Indices of the first `number` occurrences of every label (evaluation.py:33-44); ValueError when a label has
    fewer occurrences, like list.index."""
    wanted = set(labels) if label_set is None else label_set
    picked = []
    for lab in wanted:
        pos = -1
        for _ in range(number):
            pos = labels.index(lab, pos + 1)
            picked.append(pos)
    return picked


def compare_labels_given_nb(reco_label_vect, true_label_vect):
    """Per example: are the nb_true highest activations exactly the true labels (evaluation.py:8-19)."""
    reco = np.atleast_2d(reco_label_vect)
    true = np.atleast_2d(true_label_vect)
    n_true = true.sum(axis=-1)
    order = np.argsort(-reco, axis=-1)            # decreasing activation
    out = np.empty(true.shape[0], dtype=bool)
    for i in range(true.shape[0]):
        top = np.zeros(true.shape[-1])
        np.add.at(top, order[i, :int(n_true[i])], 1.)
        out[i] = (top == true[i]).all()
    return out


def score_labels_given_nb(reco_label_vect, true_label_vect):
    return np.average(compare_labels_given_nb(reco_label_vect, true_label_vect))


def compare_labels_threshold(reco_label_vect, true_label_vect, threshold):
    return ((reco_label_vect >= threshold) == true_label_vect).all(axis=-1)


def score_labels_threshold(reco_label_vect, true_label_vect, threshold):
    return np.average(compare_labels_threshold(reco_label_vect, true_label_vect, threshold))

# encoding: utf-8
"""Disk formats either side of the KL-NMF path: the trained dictionary as the reference's
experiment logger stores it, and the feature files its data-base modules read.

* dictionary store -- `multimodal/lib/logger.py:79-136` writes `<name>.json` (everything that is not an array) next to
  `<name>.npz` (arrays under `glob_<key>` / `exp_<run>_<key>`); `multimodal/experiment.py:170` stores the trained
  dictionary of every run under the key 'dictionary', and `samples/plot_info_matrix.py:123` /
  `samples/image_sound_eval_sliding.py:119` put it back with `learner.dico = logger.get_last_value('dictionary')`.
  `save_run_dictionaries` / `load_run_dictionary` / `attach_dictionary` read and write exactly that layout, so files
  move between the reference and this package in both directions.
* features -- sparse bag-of-features histograms from MATLAB files (`db/acorns.py:113`: `loadmat(f)['hac']`) and dense
  arrays from `.npz` files (`db/choreo2.py:74`: `np.load(f)['Xmotion']`), returned in the dtype / layout the C ABI
  uploads without a conversion pass over the values (CSR with sorted int32 indices and float32 values; C-contiguous
  float32), in page-locked memory when a CUDA device is present.
"""
import json

import numpy as np
import scipy.sparse as sp

from . import _native


def _split(d, plain, arrays, prefix):
    # logger.py:147-152
    for k, v in d.items():
        if isinstance(v, np.ndarray):
            arrays["%s_%s" % (prefix, k)] = np.array(v)
        else:
            plain[k] = v


def save_run_dictionaries(filename, dictionaries, glob=None, extra=None, compress=True):
    """Write one experiment run per dictionary in the layout of `Logger.save` (logger.py:79-101).

    dictionaries: list of (k x f) arrays -> `exps[i]['dictionary']`; glob: dict of global values (`store_global`);
    extra: optional list of dicts with further per-run values (arrays or json-able)."""
    glob = dict(glob or {})
    exps = []
    for i, d in enumerate(dictionaries):
        e = dict(extra[i]) if extra else {}
        e['dictionary'] = np.asarray(d)
        exps.append(e)
    exp_keys = []
    for e in exps:
        for k in e:
            if k not in exp_keys:
                exp_keys.append(k)
    to_save = {'glob': {}, 'exps': [{} for _ in exps], 'exp_keys': exp_keys, 'result_keys': [], 'has_np': False}
    arrays = {}
    _split(glob, to_save['glob'], arrays, 'glob')
    for i, e in enumerate(exps):
        _split(e, to_save['exps'][i], arrays, "exp_%d" % i)
    if arrays:
        to_save['has_np'] = True
        (np.savez_compressed if compress else np.savez)(filename, **arrays)
    with open(filename + '.json', 'w') as f:
        json.dump(to_save, f, indent=2)


def load_run_dictionary(filename, run=-1, key='dictionary'):
    """`Logger.load(filename).exps[run][key]` (logger.py:116-136) without the rest of the logger."""
    with open(filename + '.json', 'r') as f:
        data = json.load(f)
    n_runs = len(data['exps'])
    if n_runs == 0:
        raise KeyError("no experiment run in %s.json" % filename)
    idx = run if run >= 0 else n_runs + run
    if not 0 <= idx < n_runs:
        raise IndexError("run %d out of range (%d runs)" % (run, n_runs))
    if key in data['exps'][idx]:                      # stored as a plain (json) value
        return np.asarray(data['exps'][idx][key])
    if not data.get('has_np'):
        raise KeyError(key)
    with np.load(filename + '.npz') as z:
        full = "exp_%d_%s" % (idx, key)
        if full not in z.files:
            raise KeyError(key)
        return z[full]


def attach_dictionary(learner, filename, run=-1):
    """`learner.dico = logger.get_last_value('dictionary')` (samples/plot_info_matrix.py:123) with the shape the
    learner was built for checked first."""
    dico = np.ascontiguousarray(load_run_dictionary(filename, run), dtype=np.float64)
    assert dico.shape == (learner.k, sum(learner.dim)), (dico.shape, (learner.k, sum(learner.dim)))
    learner.dico = dico
    return learner


def _pinned_like(a):
    """A copy of `a` in page-locked host memory when torch sees a CUDA device; `a` itself otherwise."""
    try:
        import torch
        if not torch.cuda.is_available():
            return a
        t = torch.empty(a.shape, dtype=getattr(torch, str(a.dtype)), pin_memory=True)
        out = t.numpy()
        out[...] = a
        return out
    except Exception:
        return a


def load_mat_features(path, var='hac', pinned=True):
    """Sparse histograms from a MATLAB file as `db/acorns.py:113` reads them, as CSR in the upload layout."""
    from scipy.io import loadmat
    X = loadmat(path)[var]
    X = sp.csr_matrix(X)
    X.sum_duplicates()
    X.sort_indices()
    data = np.ascontiguousarray(X.data, dtype=np.float32)
    indices = np.ascontiguousarray(X.indices, dtype=np.int32)
    indptr = np.ascontiguousarray(X.indptr, dtype=np.int32)      # scipy keeps one index dtype per matrix
    if pinned:
        data, indices = _pinned_like(data), _pinned_like(indices)
    return sp.csr_matrix((data, indices, indptr), shape=X.shape, copy=False)


def load_npz_features(path, var='Xmotion', pinned=True):
    """A dense feature array from an `.npz` file as `db/choreo2.py:74` reads it, C-contiguous float32."""
    with np.load(path) as z:
        X = np.ascontiguousarray(z[var], dtype=np.float32)
    return _pinned_like(X) if pinned else X


class DeviceFeatures(object):
    """A dense feature matrix that lives on the GPU (one upload, any number of fits / transforms): what
    `load_npz_features_device` and `to_device` return and what `KLdivNMF.fit / fit_transform / transform` accept in place
    of an array.  Every estimator call works on a device-side copy of the requested columns
    (klnmf_create_column_view), so the resident data are never modified."""

    def __init__(self, engine):
        self.engine = engine
        self.shape = (engine.n, engine.f)
        self.ndim = 2
        self.mode = engine.mode

    def view(self, k, columns=None):
        """A fresh engine over `columns` (list of (start, stop); default: all of them) with k components."""
        return self.engine.column_view(columns or [(0, self.shape[1])], k)

    def check_input(self):
        return self.engine.check_input()

    def close(self):
        self.engine.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def to_device(X, mode=None, device=0):
    """Upload a dense host array (C-contiguous float32 / float64; pinned memory uploads at PCIe speed) once."""
    X = np.asarray(X)
    if X.dtype not in (np.float32, np.float64):
        X = X.astype(np.float64)
    # a holder, not a problem: one component and the smallest ratio scratch (the views bring their own state)
    eng = _native.Engine(X.shape[0], X.shape[1], 1, mode=mode, device=device, scratch_limit=1 << 20)
    try:
        eng.set_dense(np.ascontiguousarray(X))
    except Exception:
        eng.close()
        raise
    return DeviceFeatures(eng)


def load_npz_features_device(path, var='Xmotion', mode=None, device=0):
    """`db/choreo2.py:74` straight to the device: file -> page-locked staging -> HBM, returned as DeviceFeatures."""
    return to_device(load_npz_features(path, var, pinned=True), mode=mode, device=device)


class DictionaryCheckpoint(object):
    """Dictionary checkpoints of a running fit in the reference logger's layout (one experiment run per checkpoint
    file, key 'dictionary' plus the iteration count and the objective): `KLdivNMF(..., checkpoint=DictionaryCheckpoint(
    name, every=25))` writes `name.json` / `name.npz` every 25 iterations, and `resume()` hands back what a restarted
    fit needs -- `est._init_dictionary = ck.resume()['dictionary']`.  The dictionary is read from the device between two
    pieces of the loop (klnmf_run_resume keeps the stop test continuous across them); at k x f = 512 x 8192 that is a
    17 MB copy against seconds of iterations."""

    def __init__(self, filename, every=50):
        assert every >= 1
        self.filename, self.every = filename, int(every)

    def write(self, dictionary, iterations_done, objective):
        save_run_dictionaries(self.filename, [dictionary], glob={'iterations_done': int(iterations_done),
                                                                 'objective': float(objective)})

    def resume(self):
        with open(self.filename + '.json', 'r') as f:
            glob = json.load(f)['glob']
        return {'dictionary': np.ascontiguousarray(load_run_dictionary(self.filename), dtype=np.float64),
                'iterations_done': glob['iterations_done'], 'objective': glob['objective']}

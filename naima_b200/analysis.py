# -*- coding: utf-8 -*-
"""Callers on the output side of the hot path ("next" rows of SURVEY.md section 8f):

* ``save_run`` / ``read_run`` -- the reference's run archive (analysis.py:366-588) with the
  same group/dataset layout (``mcmc/chain``, ``mcmc/log_prob``, ``mcmc/blobN`` with unit
  attributes, ``mcmc/data/<column>``, run-info and label attributes).  h5py is not in this
  image, so the container is a ``.npz`` whose keys are the HDF5 paths and whose attributes
  are one JSON document; with h5py importable and a ``.h5``/``.hdf5`` filename the real
  HDF5 file is written instead.
* ``find_ML`` (plot.py:667-702).
* ``model_samples`` -- the posterior-sample recompute behind ``plot_fit(e_range=...)``
  (plot.py:346-393): the reference maps ``modelfn`` over ~100 samples with a process pool;
  here it is ONE batched call (every radiative class evaluates all samples in one launch).
"""
import json
import logging
import os

import numpy as np

from . import units as u
from .units import Quantity

log = logging.getLogger(__name__)

__all__ = ["save_run", "read_run", "find_ML", "model_samples"]


def _blob_columns(sampler):
    """[(array [nsteps, nwalkers, ...], [unit strings])] per blob index, without walking
    the per-walker tuples when the sampler keeps its blob records as arrays."""
    blobs = sampler.get_blobs()
    if blobs is None:
        return []
    nsteps, nwalkers = blobs.shape[:2]
    first = blobs[-1][0]
    out = []
    for idx in range(len(first)):
        item = first[idx]
        if isinstance(item, Quantity):
            units = [item.unit.to_string()]
            get = lambda b: np.asarray(b[idx].value)  # noqa: E731
        elif isinstance(item, float):
            units = [""]
            get = lambda b: np.asarray(b[idx])  # noqa: E731
        elif isinstance(item, (tuple, list)) and all(
                isinstance(x, (np.ndarray, Quantity)) for x in item):
            units = [x.unit.to_string() if isinstance(x, Quantity) else "" for x in item]
            get = lambda b: np.array([np.asarray(getattr(x, "value", x)) for x in b[idx]])  # noqa: E731
        else:
            log.warning("blob number %d has unknown format and cannot be saved", idx)
            continue
        arr = np.array([[get(blobs[s, w]) for w in range(nwalkers)] for s in range(nsteps)])
        out.append((idx, arr, units))
    return out


def save_run(filename, sampler, compression=True, clobber=False):
    """Save chain, log-probabilities, blobs, data table, labels and run information
    (analysis.py:366-471).

    ``.h5`` / ``.hdf5`` (needs h5py, absent from this image) or ``.npz``: the same names
    either way -- group ``mcmc`` with datasets ``chain``, ``log_prob``, ``blob<i>`` (flattened
    over steps x walkers, attributes ``unit`` / ``unit<j>``) and the attributes
    ``acceptance_fraction``, ``label<i>`` and the run_info keys, as the reference writes them
    (checked against tests/golden/save_run_layout.json, read off the reference source).
    ONE DIFFERENCE: the reference stores the data table with astropy's ``write_table_hdf5``
    as one compound dataset ``mcmc/data`` plus serialised meta; here it is one dataset per
    column ``mcmc/data/<column>`` with a ``unit`` attribute.  Files are therefore read back by
    this module's read_run only; a file whose ``mcmc/data`` is a compound dataset (written
    by the reference) is refused with a clear error rather than misread."""
    filename = str(filename)
    ext = os.path.splitext(filename)[1]
    if ext not in (".hdf5", ".h5", ".npz"):
        raise ValueError("Filename must end in .hdf5, .h5 or .npz suffix")
    if os.path.exists(filename) and not clobber:
        log.warning("Not writing file because file exists and clobber is False")
        return
    arrays = {"mcmc/chain": sampler.get_chain(), "mcmc/log_prob": sampler.get_log_prob()}
    attrs = {}
    for idx, arr, units in _blob_columns(sampler):
        arrays["mcmc/blob%d" % idx] = arr.reshape((-1,) + arr.shape[2:])
        if len(units) > 1:
            for j, unit in enumerate(units):
                attrs["mcmc/blob%d@unit%d" % (idx, j)] = unit
        else:
            attrs["mcmc/blob%d@unit" % idx] = units[0]
    data = sampler.data
    for col in ("energy", "energy_error_lo", "energy_error_hi", "flux", "flux_error_lo",
                "flux_error_hi", "ul", "cl"):
        if col in data:
            v = data[col]
            arrays["mcmc/data/" + col] = np.asarray(getattr(v, "value", v))
            if isinstance(v, Quantity):
                attrs["mcmc/data/%s@unit" % col] = v.unit.to_string()
    for key, val in getattr(sampler, "run_info", {}).items():
        try:
            json.dumps(val)
            attrs["mcmc@" + key] = val
        except TypeError:
            attrs["mcmc@" + key] = str(val)
    attrs["mcmc@acceptance_fraction"] = float(np.mean(sampler.acceptance_fraction))
    for i, label in enumerate(sampler.labels):
        attrs["mcmc@label%d" % i] = label
    if ext == ".npz":
        arrays["__attrs__"] = np.array(json.dumps(attrs))
        (np.savez_compressed if compression else np.savez)(filename, **arrays)
        return
    import h5py  # not in this image; kept for installations that have it

    with h5py.File(filename, "w") as f:
        for key, arr in arrays.items():
            f.create_dataset(key, data=arr, compression="gzip" if compression else None)
        for key, val in attrs.items():
            path, name = key.split("@")
            f[path].attrs[name] = val


class _result:
    """Minimal EnsembleSampler-like container for chain results (analysis.py:474-494)."""

    def get_value(self, name, flat=False):
        v = getattr(self, name)
        if flat:
            s = list(v.shape[1:])
            s[0] = int(np.prod(v.shape[:2]))
            return v.reshape(s)
        return v

    def get_chain(self, **kwargs):
        return self.get_value("chain", **kwargs)

    def get_log_prob(self, **kwargs):
        return self.get_value("log_prob", **kwargs)

    def get_blobs(self, **kwargs):
        return self.get_value("_blobs", **kwargs)


def read_run(filename, modelfn=None):
    """analysis.py:497-588 for files written by :func:`save_run`."""
    filename = str(filename)
    if filename.endswith(".npz"):
        z = np.load(filename, allow_pickle=False)
        arrays = {k: z[k] for k in z.files if k != "__attrs__"}
        attrs = json.loads(str(z["__attrs__"]))
    else:
        import h5py

        arrays, attrs = {}, {}
        with h5py.File(filename, "r") as f:
            def visit(name, obj):
                if isinstance(obj, h5py.Dataset):
                    arrays[name] = np.array(obj)
                for k, v in obj.attrs.items():
                    attrs["%s@%s" % (name, k)] = v
            f.visititems(visit)
            for k, v in f["mcmc"].attrs.items():
                attrs["mcmc@" + k] = v
    if "mcmc/data" in arrays:
        raise ValueError(
            "%s stores its data table as one compound dataset 'mcmc/data' (a file written by "
            "the reference's save_run with astropy's write_table_hdf5); naima_b200 reads the "
            "files of its own save_run, which keep one dataset per column under 'mcmc/data/'"
            % filename)
    result = _result()
    result.modelfn = modelfn
    result.chain = arrays["mcmc/chain"]
    result.log_prob = arrays["mcmc/log_prob"]
    nsteps, nwalkers, npars = result.chain.shape
    blobs, rank = [], []
    i = 0
    while "mcmc/blob%d" % i in arrays:
        ds = arrays["mcmc/blob%d" % i]
        r = np.ndim(ds[0])
        rank.append(r)
        if r <= 1:
            blobs.append(Quantity(ds, attrs.get("mcmc/blob%d@unit" % i, "")))
        else:
            blobs.append([Quantity(ds[:, j, :], attrs.get("mcmc/blob%d@unit%d" % (i, j), ""))
                          for j in range(ds.shape[1])])
        i += 1
    out = np.empty((nsteps, nwalkers), dtype=object)
    for step in range(nsteps):
        for walker in range(nwalkers):
            n = step * nwalkers + walker
            wb = []
            for j, b in enumerate(blobs):
                wb.append(b[n] if rank[j] <= 1 else [x[n] for x in b])
            out[step, walker] = wb
    result._blobs = out
    result.run_info = {k.split("@")[1]: v for k, v in attrs.items() if k.startswith("mcmc@")}
    result.acceptance_fraction = result.run_info.get("acceptance_fraction")
    result.labels = [result.run_info["label%d" % i] for i in range(npars)]
    from .utils import DataTable

    data = DataTable()
    for key, arr in arrays.items():
        if key.startswith("mcmc/data/"):
            col = key[len("mcmc/data/"):]
            unit = attrs.get(key + "@unit")
            data[col] = Quantity(arr, unit) if unit is not None else arr
    result.data = data
    return result


def find_ML(sampler, modelidx):
    """Maximum-likelihood parameters = the chain entry with the highest log-probability
    (plot.py:667-702).  Returns (ML, MLp, MLerr, (modelx, model_ML))."""
    lnprobability = sampler.get_log_prob()
    index = np.unravel_index(np.argmax(lnprobability), lnprobability.shape)
    MLp = sampler.get_chain()[index]
    blobs = sampler.get_blobs()
    modelx, model_ML = None, None
    if modelidx is not None and blobs is not None:
        blob = blobs[index][modelidx]
        if isinstance(blob, Quantity):
            modelx, model_ML = Quantity(sampler.data["energy"]), blob
        elif len(blob) == 2:
            modelx, model_ML = blob[0], blob[1]
        else:
            raise TypeError("Model {0} has wrong blob format".format(modelidx))
    elif modelidx is not None and getattr(sampler, "modelfn", None) is not None:
        out = sampler.modelfn(MLp, sampler.data)
        out = out[modelidx] if isinstance(out, (tuple, list)) else out
        if isinstance(out, Quantity):
            modelx, model_ML = Quantity(sampler.data["energy"]), out
        else:
            modelx, model_ML = out[0], out[1]
    MLerr = []
    for dist in sampler.get_chain(flat=True).T:
        hilo = np.percentile(dist, [16.0, 84.0])
        MLerr.append((hilo[1] - hilo[0]) / 2.0)
    return lnprobability[index], MLp, MLerr, (modelx, model_ML)


def model_samples(sampler, e_range, e_npoints=100, n_samples=100, last_step=False, modelidx=0,
                  seed=None):
    """Model spectra of ``n_samples`` posterior samples on a new log-spaced energy grid
    (plot.py:346-393).  The model function is called ONCE with ``pars[P, n_samples]``: every
    radiative class evaluates the whole batch in one launch.  Returns (energy, model) with
    model a Quantity ``[n_samples, e_npoints]``."""
    if getattr(sampler, "modelfn", None) is None:
        raise ValueError("sampler.modelfn is needed to recompute model samples")
    e_range = Quantity(e_range)
    if e_range.unit.physical_type != "energy":
        raise TypeError("e_range should be given in units of energy")
    energy = Quantity(np.logspace(np.log10(e_range.value[0]), np.log10(e_range.value[1]),
                                  int(e_npoints)), e_range.unit)
    data = {"energy": energy,
            "flux": Quantity(np.zeros(energy.shape), Quantity(sampler.data["flux"]).unit)}
    chain = sampler.get_chain()[-1] if last_step else sampler.get_chain(flat=True)
    rng = np.random.RandomState(seed)
    pars = chain[rng.randint(len(chain), size=int(n_samples))]
    out = sampler.modelfn(pars.T, data)
    if isinstance(out, (tuple, list)):
        out = out[modelidx]
    if isinstance(out, (tuple, list)):  # (energy, flux) pair blob
        return out[0], Quantity(out[1])
    return energy, Quantity(out)

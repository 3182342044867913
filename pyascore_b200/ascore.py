"""`PyAscore`: the reference's scoring class, same constructor, `score()` call and result
attributes (reference: pyascore/ptm_scoring/Ascore.pyx:12-288), computed on a B200 through
libpyascore_b200.  A single `score()` is a batch of one through the same kernels.
"""
import numpy as np

from .batch import Scorer


def _as_buffer(name, arr, dtype, cname):
    # mirror Cython's typed-buffer argument checks (Ascore.pyx:103-108)
    if arr is None:
        raise TypeError("Argument '%s' must not be None" % name)
    if not isinstance(arr, np.ndarray):
        raise TypeError("Argument '%s' has incorrect type (expected numpy.ndarray, got %s)" % (name, type(arr).__name__))
    if arr.dtype != dtype:
        raise ValueError("Buffer dtype mismatch, expected '%s' but got '%s'" % (cname, _cname(arr.dtype)))
    if arr.ndim != 1:
        raise ValueError("Buffer has wrong number of dimensions (expected 1, got %d)" % arr.ndim)
    if not arr.flags.c_contiguous:
        raise ValueError("ndarray is not C-contiguous")
    return arr


def _cname(dt):
    return {"float32": "float", "float64": "double", "int64": "long", "int32": "int", "uint32": "unsigned int",
            "uint64": "unsigned long"}.get(np.dtype(dt).name, np.dtype(dt).name)


class PyAscore:
    """Drop-in for pyascore.PyAscore (see the reference docstring, Ascore.pyx:13-59)."""

    _MOD_CAP = 64          # mods per PSM the persistent result arrays hold (grown on demand)

    def __init__(self, bin_size, n_top, mod_group, mod_mass, mz_error=.5, fragment_types="by", device=0):
        import ctypes as C
        from . import _lib
        self._scorer = Scorer(bin_size, n_top, mod_group, mod_mass, mz_error, fragment_types, device=device)
        self._batch = None
        self._res = None
        self._k = 0
        # A single score() is a batch of one through pa_score_batch.  Everything about that batch that does not
        # change from call to call -- the one-entry CSR index arrays, the result arrays, the two ctypes structs --
        # is built once here, so that a call costs a few attribute stores before it enters the library.
        self._C, self._lib = C, _lib
        self._a = dict(spec_off=np.zeros(2, np.int64), psm_spec=np.zeros(1, np.int32), pep_off=np.zeros(2, np.int32),
                       n_mod=np.zeros(1, np.int32), max_charge=np.zeros(1, np.int32), aux_off=np.zeros(2, np.int32),
                       mod_off=np.zeros(2, np.int64))
        self._no_aux = (np.zeros(1, np.uint32), np.zeros(1, np.float32))
        self._pb = _lib.PaBatch()
        self._pb.n_spec, self._pb.n_psm = 1, 1
        for k, v in self._a.items():
            setattr(self._pb, k, v.ctypes.data)
        self._pr = _lib.PaResults()
        self._alloc_results(self._MOD_CAP)

    def _alloc_results(self, cap):
        self._cap = cap
        self._out = dict(best_sig=np.zeros(1, np.uint64), best_score=np.zeros(1, np.float32), n_iso=np.zeros(1, np.int64),
                         n_sites=np.zeros(1, np.int32), ascores=np.zeros(cap, np.float32), alt_sites=np.zeros(cap, np.uint64),
                         psm_status=np.zeros(1, np.int32))
        for k, v in self._out.items():
            setattr(self._pr, k, v.ctypes.data)

    def add_neutral_loss(self, group, mass):
        self._scorer.add_neutral_loss(group, mass)

    def score(self, mz_arr, int_arr, peptide, n_of_mod, max_fragment_charge=1, aux_mod_pos=None, aux_mod_mass=None):
        mz_arr = _as_buffer("mz_arr", mz_arr, np.float64, "double")
        int_arr = _as_buffer("int_arr", int_arr, np.float64, "double")
        if not isinstance(peptide, str):
            raise TypeError("Argument 'peptide' has incorrect type (expected str, got %s)" % type(peptide).__name__)
        if aux_mod_pos is not None:
            aux_mod_pos = _as_buffer("aux_mod_pos", aux_mod_pos, np.uint32, "unsigned int")
        if aux_mod_mass is not None:
            aux_mod_mass = _as_buffer("aux_mod_mass", aux_mod_mass, np.float32, "float")
        if n_of_mod < 0 or max_fragment_charge < 0:
            raise OverflowError("can't convert negative value to size_t")
        if mz_arr.size != int_arr.size:
            raise ValueError("mz_arr and int_arr differ in length")
        use_aux = aux_mod_pos is not None and aux_mod_mass is not None
        pep = np.frombuffer(peptide.encode("utf8"), np.uint8)
        n_aux = int(aux_mod_pos.size) if use_aux else 0
        if use_aux and aux_mod_mass.size < n_aux:
            raise ValueError("aux_mod_mass is shorter than aux_mod_pos")
        k = int(n_of_mod)
        if k > self._cap:
            self._alloc_results(max(k, 2 * self._cap))
        a, pb = self._a, self._pb
        a["spec_off"][1] = mz_arr.size
        a["pep_off"][1] = pep.size
        a["n_mod"][0] = k
        a["max_charge"][0] = max_fragment_charge
        a["aux_off"][1] = n_aux
        a["mod_off"][1] = k
        ap, am = (aux_mod_pos, aux_mod_mass) if n_aux else self._no_aux
        pb.mz, pb.inten, pb.pep = mz_arr.ctypes.data, int_arr.ctypes.data, (pep.ctypes.data if pep.size else self._no_aux[0].ctypes.data)
        pb.aux_pos, pb.aux_mass = ap.ctypes.data, am.ctypes.data
        self._res = None
        # the batch as the result properties see it (and what keeps the caller's arrays alive until the next call)
        self._batch = dict(mz=mz_arr, inten=int_arr, pep=pep, aux_pos=aux_mod_pos if n_aux else np.zeros(0, np.uint32),
                           aux_mass=aux_mod_mass[:n_aux] if n_aux else np.zeros(0, np.float32))
        self._k = k
        self._pep = peptide
        sc = self._scorer
        rc = sc.L.pa_score_batch(sc.h, self._C.byref(pb), self._C.byref(self._pr), self._lib.PA_KEEP_ISOFORMS)
        if rc != 0:
            sc._raise(rc)
        status = int(self._out["psm_status"][0])
        if status != 0:
            raise ValueError("pyascore_b200: cannot score %r: %s" % (peptide, self._lib.PSM_STATUS.get(status, status)))
        self._res = self._out

    def _need(self):
        if self._res is None:
            raise RuntimeError("no peptide has been scored yet")

    @property
    def best_sequence(self):
        self._need()
        if int(self._res["n_iso"][0]) == 0:
            return ""
        b = self._batch
        return self._scorer.format_sequence(bytes(b["pep"]), self._k, b["aux_pos"], b["aux_mass"],
                                            int(self._res["best_sig"][0]))

    @property
    def best_score(self):
        self._need()
        return float(self._res["best_score"][0])

    @property
    def pep_scores(self):
        self._need()
        sig, cnt, sc, w, tot = self._scorer.fetch_pep_scores(0)
        S = int(self._res["n_sites"][0])
        b = self._batch
        out = []
        for i in range(sig.size):
            bits = int(sig[i])
            out.append(dict(
                signature=np.array([(bits >> j) & 1 for j in range(S)], np.int32), counts=cnt[i].copy(),
                scores=sc[i].copy(), weighted_score=float(w[i]), total_fragments=int(tot[i]),
                sequence=self._scorer.format_sequence(bytes(b["pep"]), self._k, b["aux_pos"], b["aux_mass"], bits)))
        return out

    @property
    def ascores(self):
        self._need()
        return np.array(self._res["ascores"][:self._k], np.float32)

    @property
    def alt_sites(self):
        self._need()
        sites = self._scorer.site_positions(bytes(self._batch["pep"]))
        out = []
        for j in range(self._k):
            m = int(self._res["alt_sites"][j])
            out.append(np.array([sites[u] for u in range(len(sites)) if (m >> u) & 1], np.uint32))
        return out

    def calculate_ambiguity(self, ref_score, other_score):
        self._need()

        def bits(d):
            return sum(int(v != 0) << j for j, v in enumerate(np.asarray(d["signature"])))
        return self._scorer.calculate_ambiguity(0, bits(ref_score), ref_score["scores"], ref_score["weighted_score"],
                                                bits(other_score), other_score["scores"], other_score["weighted_score"])

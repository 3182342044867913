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

    def __init__(self, bin_size, n_top, mod_group, mod_mass, mz_error=.5, fragment_types="by", device=0):
        self._scorer = Scorer(bin_size, n_top, mod_group, mod_mass, mz_error, fragment_types, device=device)
        self._batch = None
        self._res = None
        self._k = 0

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
        batch = dict(
            spec_off=np.array([0, mz_arr.size], np.int64), mz=mz_arr, inten=int_arr,
            psm_spec=np.zeros(1, np.int32), pep_off=np.array([0, pep.size], np.int32), pep=pep,
            n_mod=np.array([n_of_mod], np.int32), max_charge=np.array([max_fragment_charge], np.int32),
            aux_off=np.array([0, aux_mod_pos.size if use_aux else 0], np.int32),
            aux_pos=aux_mod_pos if use_aux else np.zeros(0, np.uint32),
            aux_mass=aux_mod_mass if use_aux else np.zeros(0, np.float32))
        res = self._scorer.score_batch(batch, keep_isoforms=True)
        status = int(res["psm_status"][0])
        self._batch, self._res, self._k = batch, res, int(n_of_mod)
        self._pep = peptide
        if status != 0:
            from ._lib import PSM_STATUS
            self._res = None
            raise ValueError("pyascore_b200: cannot score %r: %s" % (peptide, PSM_STATUS.get(status, status)))

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

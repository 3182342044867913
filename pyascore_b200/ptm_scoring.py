"""The reference's secondary public classes (exported at pyascore/__init__.py:17) as thin cursors
over libpyascore_b200:

    PyBinnedSpectra                     pyascore/ptm_scoring/Spectra.pyx:8-125
    PyModifiedPeptide, PyFragmentGraph  pyascore/ptm_scoring/ModifiedPeptide.pyx:10-329
    PyLogMath, PyBinomialDist, PyPowerSetSum   pyascore/ptm_scoring/Util.pyx:6-134

Every number they hand out is computed by the library -- the binning kernel, the fragment /
site-determining-ion / log-math probe kernels that reuse the hot path's device functions, and
the host routine that builds the kernels' neutral-loss tables.  Python only keeps the cursor state
(which bin / rank / signature / fragment the caller is looking at), as the C++ classes do.
"""
import ctypes as C

import numpy as np

from . import _lib
from .batch import Scorer


def _u32(a, name):
    if a is None:
        return np.zeros(0, np.uint32)
    if not isinstance(a, np.ndarray) or a.dtype != np.uint32:
        raise ValueError("Buffer dtype mismatch, expected 'unsigned int' for %s" % name)
    return np.ascontiguousarray(a)


_service = None


def _service_scorer():
    """scorer used by the classes that carry no scoring configuration of their own"""
    global _service
    if _service is None:
        _service = Scorer(100., 10, "STY", 79.966331)
    return _service


# ---------------------------------------------------------------------------------------------
class PyBinnedSpectra:
    """Top-`n_top` peaks per `bin_size`-Th bin with a (bin, rank) cursor (Spectra.pyx:8-125)."""

    def __init__(self, bin_size, n_top, device=0):
        self.L = _lib.load()
        h = C.c_void_p()
        rc = self.L.pa_create_binner(bin_size, int(n_top), int(device), C.byref(h))
        if rc != 0:
            raise (RuntimeError if rc == -1 else ValueError)("pyascore_b200: " + _lib.last_error(None))
        self.h = h
        self._bin_size, self._n_top = float(np.float32(bin_size)), int(n_top)
        self._bins, self._n_bins, self._min, self._max = [], 0, 0., 0.
        self._bin = self._rank = 0

    def __del__(self):
        if getattr(self, "h", None):
            self.L.pa_destroy(self.h)
            self.h = None

    def consume_spectra(self, mz_arr, int_arr):
        from .ascore import _as_buffer
        mz_arr = _as_buffer("mz_arr", mz_arr, np.float64, "double")
        int_arr = _as_buffer("int_arr", int_arr, np.float64, "double")
        n = mz_arr.size
        if n == 0 or int_arr.size != n:
            raise ValueError("empty spectrum or arrays of different length")
        off = np.array([0, n], np.int64)
        omz, ork = np.zeros(n, np.float32), np.zeros(n, np.uint8)
        oidx, obin = np.zeros(n, np.int32), np.zeros(n, np.int32)
        ocnt, obnd = np.zeros(1, np.int32), np.zeros(3, np.float32)
        rc = self.L.pa_bin_spectra_ex(self.h, 1, off.ctypes.data, mz_arr.ctypes.data, int_arr.ctypes.data,
                                      omz.ctypes.data, ork.ctypes.data, ocnt.ctypes.data, oidx.ctypes.data,
                                      obin.ctypes.data, obnd.ctypes.data)
        if rc != 0:
            raise RuntimeError("pyascore_b200: " + _lib.last_error(self.h))
        c = int(ocnt[0])
        self._min, self._max, self._n_bins = float(obnd[0]), float(obnd[1]), int(obnd[2])
        bins = [[] for _ in range(self._n_bins)]
        for j in np.lexsort((ork[:c], obin[:c])):
            bins[obin[j]].append((float(mz_arr[oidx[j]]), float(int_arr[oidx[j]])))
        self._bins = bins
        self._bin = self._rank = 0

    def _peak(self):
        try:
            return self._bins[self._bin][self._rank]
        except IndexError:      # the reference's .at() throws std::out_of_range here (Spectra.cpp:70-80)
            raise IndexError("no peak at bin %d, rank %d" % (self._bin, self._rank)) from None

    mz = property(lambda self: self._peak()[0])
    intensity = property(lambda self: self._peak()[1])

    @property
    def n_peaks(self):
        if self._bin >= len(self._bins):
            raise IndexError("bin %d out of range" % self._bin)
        return len(self._bins[self._bin])

    @property
    def bin(self):
        return self._bin

    @bin.setter
    def bin(self, new_bin):
        self._bin = min(int(new_bin), self._n_bins)

    def reset_bin(self):
        self._bin = 0

    def next_bin(self):
        self._bin = min(self._bin + 1, self._n_bins)

    @property
    def rank(self):
        return self._rank

    @rank.setter
    def rank(self, new_rank):
        self._rank = min(int(new_rank), self._n_top)

    def reset_rank(self):
        self._rank = 0

    def next_rank(self):
        self._rank = min(self._rank + 1, self._n_top)

    min_mz = property(lambda self: self._min)
    max_mz = property(lambda self: self._max)
    bin_size = property(lambda self: self._bin_size)
    n_bins = property(lambda self: self._n_bins)


# ---------------------------------------------------------------------------------------------
class PyModifiedPeptide:
    """Fragments / site-determining ions / bracketed sequences of one peptide
    (ModifiedPeptide.pyx:10-158).  Scorer settings as in PyAscore."""

    def __init__(self, mod_group, mod_mass, mz_error=.5, fragment_types="by", device=0):
        self._scorer = Scorer(100., 10, mod_group, mod_mass, mz_error, fragment_types, device=device)
        self.L = self._scorer.L
        self._pep = None

    def add_neutral_loss(self, group, mass):
        self._scorer.add_neutral_loss(group, mass)

    def consume_peptide(self, peptide, n_of_mod, max_fragment_charge=1, aux_mod_pos=None, aux_mod_mass=None):
        use_aux = aux_mod_pos is not None and aux_mod_mass is not None
        self._aux_pos = _u32(aux_mod_pos, "aux_mod_pos") if use_aux else np.zeros(0, np.uint32)
        if use_aux and (not isinstance(aux_mod_mass, np.ndarray) or aux_mod_mass.dtype != np.float32):
            raise ValueError("Buffer dtype mismatch, expected 'float' for aux_mod_mass")
        self._aux_mass = np.ascontiguousarray(aux_mod_mass) if use_aux else np.zeros(0, np.float32)
        self._pep = peptide.encode("ascii")
        self._pep_arr = np.frombuffer(self._pep, np.uint8)
        self._k, self._Z = int(n_of_mod), int(max_fragment_charge)
        self._sites = [int(p) - 1 for p in self._scorer.site_positions(self._pep)]    # residue index of site j
        if self._pep and len(self._sites) > 63:
            raise ValueError("more than 63 modifiable residues")

    def _need(self):
        if self._pep is None:
            raise RuntimeError("no peptide has been consumed yet")

    def _bits(self, signature, what="signature"):
        sig = np.asarray(signature)
        if sig.size != len(self._sites):
            raise ValueError("%s has %d entries for %d modifiable residues" % (what, sig.size, len(self._sites)))
        return sum(1 << j for j in range(sig.size) if int(sig[j]) != 0)

    def get_peptide(self, signature=None):
        self._need()
        sig = np.zeros(0, np.uint32) if signature is None else _u32(signature, "signature")
        if sig.size == 0:                       # "just use the first signature" (cpp/ModifiedPeptide.cpp:201-204)
            bits = (1 << min(self._k, len(self._sites))) - 1
        else:
            bits = sum(1 << j for j in range(min(sig.size, 64)) if int(sig[j]) == 1)
        return self._scorer.format_sequence(self._pep, self._k, self._aux_pos, self._aux_mass, bits)

    def get_fragment_graph(self, fragment_type, charge_state, mode="all"):
        self._need()
        return PyFragmentGraph(self, fragment_type, charge_state, mode)

    def get_site_determining_ions(self, sig_1, sig_2, fragment_type, max_charge):
        self._need()
        sig_1, sig_2 = _u32(sig_1, "sig_1"), _u32(sig_2, "sig_2")
        n = min(sig_1.size, sig_2.size)
        ba, bb = self._bits(sig_1[:n], "sig_1"), self._bits(sig_2[:n], "sig_2")
        cap = 2 * max(len(self._pep), 1) * 16 * max(int(max_charge), 1)
        oa, ob = np.zeros(cap, np.float32), np.zeros(cap, np.float32)
        na, nb = C.c_int32(), C.c_int32()
        rc = self.L.pa_site_determining_ions(self._scorer.h, self._pep_arr.ctypes.data, len(self._pep),
                                             self._aux_pos.ctypes.data, self._aux_mass.ctypes.data, self._aux_pos.size,
                                             ba, bb, fragment_type.encode("ascii")[:1], int(max_charge),
                                             oa.ctypes.data, C.byref(na), ob.ctypes.data, C.byref(nb), cap)
        if rc != 0:
            self._scorer._raise(rc)
        return oa[:na.value].copy(), ob[:nb.value].copy()

    # used by PyFragmentGraph
    def _fragment_table(self, bits, fragment_type, charge):
        n = len(self._pep)
        mz, nv = np.zeros((n, 16), np.float32), np.zeros(n, np.int32)
        rc = self.L.pa_fragment_table(self._scorer.h, self._pep_arr.ctypes.data, n, self._aux_pos.ctypes.data,
                                      self._aux_mass.ctypes.data, self._aux_pos.size, bits,
                                      fragment_type.encode("ascii")[:1], int(charge), mz.ctypes.data, nv.ctypes.data)
        if rc != 0:
            self._scorer._raise(rc)
        return mz, nv


class PyFragmentGraph:
    """Iterator over (positional isoform x residue x neutral-loss variant) of one ion series
    (ModifiedPeptide.pyx:160-329).  The cursor logic restates cpp/ModifiedPeptide.cpp:410-568 on
    traversal steps (step t = t-th residue from the series' own terminus); the m/z values of the
    current isoform come from the library (`pa_fragment_table`)."""

    def __init__(self, peptide, fragment_type, charge_state, mode="all"):
        assert mode in ("all", "reduced")
        if isinstance(fragment_type, (bytes, bytearray)):
            fragment_type = fragment_type.decode()
        elif isinstance(fragment_type, int):
            fragment_type = chr(fragment_type)
        if fragment_type not in "bcyzZ" or len(fragment_type) != 1:
            raise ValueError("fragment type %r not in 'bcyzZ'" % (fragment_type,))
        self.mode, self._p, self._type, self._charge = mode, peptide, fragment_type, int(charge_state)
        self._fwd = fragment_type in "bc"
        self._L = len(peptide._pep)
        steps = [i if self._fwd else self._L - 1 - i for i in peptide._sites]
        self._site_steps = sorted(steps)                        # traversal step of every modifiable residue
        self.reset_iterator()

    fragment_type = property(lambda self: self._type)
    charge_state = property(lambda self: self._charge)

    # ---- signature level -------------------------------------------------------------------
    def reset_iterator(self):
        k, S = self._p._k, len(self._site_steps)
        self._sig = [1 if j < k else 0 for j in range(S)]       # traversal order
        self._outstanding = max(k - S, 0)
        self.reset_fragment()

    def is_signature_end(self):
        return self._outstanding > 0

    def incr_signature(self):
        if self.is_signature_end():
            raise RuntimeError("signature iterator exhausted")   # the reference throws 40 (process abort)
        S = len(self._sig)
        self._outstanding = 1
        final = None
        it = S - 1
        while it >= 0 and self._outstanding:
            if self._sig[it] > 0:
                final = it
                self._sig[it] = 0
                if self._outstanding == S - it:
                    self._outstanding += 1
                    it -= 1
                    continue
                while self._outstanding:
                    it += 1
                    self._sig[it] += 1
                    self._outstanding -= 1
            it -= 1
        if not self.is_signature_end():
            # rewind to the left-most changed residue (prefix sums before it stay valid)
            self._t = min(self._t, self._site_steps[final])
            self._load()
            self._calc()

    def set_signature(self, new_signature):
        sig = _u32(new_signature, "new_signature")
        if sig.size != len(self._sig):
            raise ValueError("signature has %d entries for %d modifiable residues" % (sig.size, len(self._sig)))
        vals = [int(v) for v in sig]
        self._sig = vals if self._fwd else vals[::-1]
        self.reset_fragment()

    def get_signature(self):
        sig = self._sig if self._fwd else self._sig[::-1]
        return np.array(sig, dtype=np.uint64)

    # ---- fragment level ----------------------------------------------------------------------
    def _load(self):
        bits = 0
        S = len(self._sig)
        for j, v in enumerate(self._sig):
            if v:
                site = j if self._fwd else S - 1 - j             # site index N->C
                bits |= 1 << site
        self._mz, self._nvar = self._p._fragment_table(bits, self._type, self._charge)

    def _calc(self):
        """calculateFragment + updateLosses at the current step: variant cursor back to 0"""
        self._t_calc = self._t
        self._iter_n = int(self._nvar[self._t])
        self._v = 0

    def reset_fragment(self):
        self._t = 0
        self._load()
        self._calc()

    def is_fragment_end(self):
        return self._t == self._L - 1 and not (self._v + 1 < self._iter_n)

    def incr_fragment(self):
        if self.is_signature_end() or self.is_fragment_end():
            raise RuntimeError("fragment iterator exhausted")    # the reference throws 40
        if self._v + 1 < self._iter_n:
            self._v += 1
        else:
            self._t += 1
            if not self.is_fragment_end():                       # evaluated on the exhausted variant cursor
                self._calc()

    def get_fragment_mz(self):
        return float(self._mz[self._t_calc, self._v])

    def get_fragment_size(self):
        return self._t_calc + 1

    def get_fragment_seq(self):
        pep = self._p._pep.decode()
        return pep[:self._t_calc + 1] if self._fwd else pep[::-1][:self._t_calc + 1]

    def iter_permutations(self):
        while not self.is_signature_end():
            yield self
            self.incr_signature()
            if self.mode == "all":
                self.reset_fragment()

    def iter_fragments(self):
        while not self.is_fragment_end():
            label = self.fragment_type + str(self.get_fragment_size())
            result = (self.get_fragment_mz(), label)
            self.incr_fragment()
            yield result


# ---------------------------------------------------------------------------------------------
def _log_math(op, x=None, y=None, k=None, tr=None, prob=0.5):
    s = _service_scorer()
    out = np.zeros(1, np.float32)
    if op == 0:
        a, b = np.array([x], np.float32), np.array([y], np.float32)
        rc = s.L.pa_log_math(s.h, 0, 1, a.ctypes.data, b.ctypes.data, None, None, C.c_float(prob), out.ctypes.data)
    else:
        if k < 0 or tr < 0:
            raise OverflowError("can't convert negative value to size_t")
        a, b = np.array([k], np.int32), np.array([tr], np.int32)
        rc = s.L.pa_log_math(s.h, op, 1, None, None, a.ctypes.data, b.ctypes.data, C.c_float(prob), out.ctypes.data)
    if rc != 0:
        s._raise(rc)
    return float(out[0])


class PyLogMath:
    """float32 log-space helpers (Util.pyx:6-46 -> cpp/Util.cpp:16-41)."""

    def log_sum(self, a, b):
        return _log_math(0, x=a, y=b)

    def log_bin_coef(self, k, n):
        return _log_math(1, k=int(k), tr=int(n))


class PyBinomialDist:
    """log pmf / upper tail of Binomial(trials, prob) in the reference's float32 arithmetic
    (Util.pyx:48-98 -> cpp/Util.cpp:47-83)."""

    def __init__(self, prob):
        self._prob = float(np.float32(prob))

    def log_pmf(self, successes, trials):
        return _log_math(2, k=int(successes), tr=int(trials), prob=self._prob)

    def log_pvalue(self, successes, trials):
        return _log_math(3, k=int(successes), tr=int(trials), prob=self._prob)

    def log10_pvalue(self, successes, trials):
        return _log_math(4, k=int(successes), tr=int(trials), prob=self._prob)


class PyPowerSetSum:
    """Sorted, de-duplicated float32 subset sums with a cursor (Util.pyx:100-134)."""

    def __init__(self, target=None, max_depth=0):
        self._sums = np.zeros(1, np.float32)
        self._pos = 0
        if target is not None:
            self.reset(target, max_depth)

    def reset(self, target=None, max_depth=0):
        self._pos = 0
        if target is None:
            return
        if not isinstance(target, np.ndarray) or target.dtype != np.float32:
            raise ValueError("Buffer dtype mismatch, expected 'float' for target")
        t = np.ascontiguousarray(target)
        L = _lib.load()
        n = L.pa_power_set_sums(t.ctypes.data, t.size, int(max_depth), None, 0)
        if n < 0:
            raise ValueError("pa_power_set_sums failed (%d)" % n)
        out = np.zeros(n, np.float32)
        L.pa_power_set_sums(t.ctypes.data, t.size, int(max_depth), out.ctypes.data, n)
        self._sums = out

    def has_next(self):
        return self._pos < self._sums.size - 1

    def next(self):
        if not self.has_next():
            raise StopIteration("no further sum")                # the reference throws 40
        self._pos += 1

    def get_sum(self):
        return float(self._sums[self._pos])

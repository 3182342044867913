"""TEST INFRASTRUCTURE ONLY -- ctypes views of the two CPU checkers.

* oracle/_ref/librefshim.so : the UNMODIFIED reference C++ (cpp/*.cpp under
  /root/reference) compiled by oracle/Makefile behind oracle/ref_shim.cpp  -> `RefPyAscore`
* oracle/liboracle.so       : the plain-C restatement oracle/ascore_oracle.c -> `OraclePyAscore`

Both get the surface of the reference's Cython class (pyascore/ptm_scoring/Ascore.pyx:12-288)
plus stage probes, so tests and the golden generator can treat either as "the reference".

Nothing in pyascore_b200/ may import this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = {"refshim_": os.path.join(_HERE, "_ref", "librefshim.so"),
       "orc_": os.path.join(_HERE, "liboracle.so")}

_libs = {}

c_f = C.c_float
c_sz = C.c_size_t
c_vp = C.c_void_p
c_i64 = C.c_int64
c_long = C.c_long


def available(prefix="refshim_"):
    return os.path.exists(_SO[prefix])


class _Pref:
    """attribute access L.refshim_x -> real symbol <prefix>x"""
    def __init__(self, dll, prefix):
        self._dll, self._prefix = dll, prefix
    def __getattr__(self, name):
        return getattr(self._dll, self._prefix + name[len("refshim_"):])


def lib(prefix="refshim_"):
    if prefix not in _libs:
        L = _Pref(C.CDLL(_SO[prefix]), prefix)
        L.refshim_new.restype = c_vp
        L.refshim_new.argtypes = [c_f, c_sz, C.c_char_p, c_f, c_f, C.c_char_p]
        L.refshim_free.argtypes = [c_vp]
        L.refshim_add_neutral_loss.argtypes = [c_vp, C.c_char_p, c_f]
        L.refshim_score.argtypes = [c_vp, c_vp, c_vp, c_sz, C.c_char_p, c_sz, c_sz, c_vp, c_vp, c_sz]
        L.refshim_best_sequence.argtypes = [c_vp, C.c_char_p, C.c_int]
        L.refshim_best_score.restype = c_f
        L.refshim_best_score.argtypes = [c_vp]
        L.refshim_n_pep_scores.restype = c_sz
        L.refshim_n_pep_scores.argtypes = [c_vp]
        L.refshim_sig_len.restype = c_sz
        L.refshim_sig_len.argtypes = [c_vp]
        L.refshim_pep_scores.argtypes = [c_vp] * 6
        L.refshim_sequences.restype = c_long
        L.refshim_sequences.argtypes = [c_vp, C.c_char_p, c_long]
        L.refshim_n_ascores.restype = c_sz
        L.refshim_n_ascores.argtypes = [c_vp]
        L.refshim_ascores.argtypes = [c_vp, c_vp]
        L.refshim_alt_sites.restype = c_long
        L.refshim_alt_sites.argtypes = [c_vp, c_sz, c_vp, c_long]
        L.refshim_calculate_ambiguity.restype = c_f
        L.refshim_calculate_ambiguity.argtypes = [c_vp, c_sz, c_sz, c_vp, c_vp, c_vp, c_f, c_i64,
                                                  c_vp, c_vp, c_vp, c_f, c_i64]
        L.refshim_binned.restype = c_long
        L.refshim_binned.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_long, c_vp, c_vp, c_vp]
        L.refshim_consume_spectra.argtypes = [c_vp, c_vp, c_vp, c_sz]
        L.refshim_consume_peptide.argtypes = [c_vp, C.c_char_p, c_sz, c_sz, c_vp, c_vp, c_sz]
        L.refshim_fragment_graph.restype = c_long
        L.refshim_fragment_graph.argtypes = [c_vp, C.c_char, c_sz, c_vp, c_long, c_vp, c_vp, c_long, c_vp]
        L.refshim_site_determining.restype = c_long
        L.refshim_site_determining.argtypes = [c_vp, c_sz, c_vp, c_vp, C.c_char, c_sz, c_vp, c_vp,
                                               c_vp, c_vp, c_long]
        L.refshim_get_peptide.argtypes = [c_vp, c_sz, c_vp, C.c_char_p, C.c_int]
        L.refshim_has_match.argtypes = [c_vp, c_f, c_vp, c_vp]
        L.refshim_log_sum.restype = c_f
        L.refshim_log_sum.argtypes = [c_f, c_f]
        L.refshim_log_bin_coef.restype = c_f
        L.refshim_log_bin_coef.argtypes = [c_sz, c_sz]
        L.refshim_binom_new.restype = c_vp
        L.refshim_binom_new.argtypes = [c_f]
        L.refshim_binom_free.argtypes = [c_vp]
        for n in ("log_pmf", "log_pvalue", "log10_pvalue"):
            f = getattr(L, "refshim_binom_" + n)
            f.restype = c_f
            f.argtypes = [c_vp, c_sz, c_sz]
        L.refshim_power_set_sum.restype = c_long
        L.refshim_power_set_sum.argtypes = [c_vp, c_sz, c_sz, c_vp, c_long]
        L.refshim_score_batch.argtypes = [c_vp, c_i64] + [c_vp] * 13 + [C.c_int32]
        _libs[prefix] = L
    return _libs[prefix]


def _p(a):
    return None if a is None else a.ctypes.data_as(c_vp)


class _CScorer:
    PREFIX = "refshim_"

    """Reference scorer with the constructor / score() / properties of Ascore.pyx:64-288."""

    def __init__(self, bin_size, n_top, mod_group, mod_mass, mz_error=.5, fragment_types="by"):
        self.L = lib(self.PREFIX)
        self.n_top = int(n_top)
        self.h = self.L.refshim_new(bin_size, n_top, mod_group.encode(), mod_mass, mz_error,
                                    fragment_types.encode())
        self._k = 0

    def __del__(self):
        if getattr(self, "h", None):
            self.L.refshim_free(self.h)
            self.h = None

    def add_neutral_loss(self, group, mass):
        self.L.refshim_add_neutral_loss(self.h, group.encode(), mass)

    def score(self, mz_arr, int_arr, peptide, n_of_mod, max_fragment_charge=1,
              aux_mod_pos=None, aux_mod_mass=None):
        mz_arr = np.ascontiguousarray(mz_arr, dtype=np.float64)
        int_arr = np.ascontiguousarray(int_arr, dtype=np.float64)
        if aux_mod_pos is not None and aux_mod_mass is not None:
            aux_mod_pos = np.ascontiguousarray(aux_mod_pos, dtype=np.uint32)
            aux_mod_mass = np.ascontiguousarray(aux_mod_mass, dtype=np.float32)
            na = aux_mod_pos.size
            if na == 0:  # keep non-NULL pointers like &arr[0] would be in Cython
                aux_mod_pos = np.zeros(1, np.uint32)
                aux_mod_mass = np.zeros(1, np.float32)
        else:
            aux_mod_pos = aux_mod_mass = None
            na = 0
        self._k = int(n_of_mod)
        self.L.refshim_score(self.h, _p(mz_arr), _p(int_arr), mz_arr.size, peptide.encode(),
                             n_of_mod, max_fragment_charge, _p(aux_mod_pos), _p(aux_mod_mass), na)

    @property
    def best_sequence(self):
        buf = C.create_string_buffer(4096)
        self.L.refshim_best_sequence(self.h, buf, 4096)
        return buf.value.decode()

    @property
    def best_score(self):
        return self.L.refshim_best_score(self.h)

    def pep_score_tables(self):
        n = self.L.refshim_n_pep_scores(self.h)
        S = self.L.refshim_sig_len(self.h)
        D = self.n_top
        sig = np.zeros((n, S), np.int32)
        cnt = np.zeros((n, D), np.int32)
        sc = np.zeros((n, D), np.float32)
        w = np.zeros(n, np.float32)
        tot = np.zeros(n, np.int64)
        if n:
            self.L.refshim_pep_scores(self.h, _p(sig), _p(cnt), _p(sc), _p(w), _p(tot))
        return sig, cnt, sc, w, tot

    def sequences(self):
        need = self.L.refshim_sequences(self.h, None, 0)
        buf = C.create_string_buffer(need)
        self.L.refshim_sequences(self.h, buf, need)
        s = buf.value.decode()
        return s.split("\n") if self.L.refshim_n_pep_scores(self.h) else []

    @property
    def pep_scores(self):
        sig, cnt, sc, w, tot = self.pep_score_tables()
        seqs = self.sequences()
        return [dict(signature=sig[i], counts=cnt[i], scores=sc[i], weighted_score=float(w[i]),
                     total_fragments=int(tot[i]), sequence=seqs[i]) for i in range(len(w))]

    @property
    def ascores(self):
        n = self.L.refshim_n_ascores(self.h)
        out = np.zeros(n, np.float32)
        if n:
            self.L.refshim_ascores(self.h, _p(out))
        return out

    @property
    def alt_sites(self):
        res = []
        for j in range(self._k):
            buf = np.zeros(256, np.uint32)
            n = self.L.refshim_alt_sites(self.h, j, _p(buf), 256)
            res.append(buf[:n].copy())
        return res

    def calculate_ambiguity(self, a, b):
        def arrs(d):
            return (np.ascontiguousarray(d["signature"], np.int32), np.ascontiguousarray(d["counts"], np.int32),
                    np.ascontiguousarray(d["scores"], np.float32))
        sa, ca, fa = arrs(a)
        sb, cb, fb = arrs(b)
        return self.L.refshim_calculate_ambiguity(
            self.h, sa.size, ca.size, _p(sa), _p(ca), _p(fa), a["weighted_score"], a["total_fragments"],
            _p(sb), _p(cb), _p(fb), b["weighted_score"], b["total_fragments"])

    # ---- stage probes ----
    def binned(self, mz_arr, int_arr):
        mz_arr = np.ascontiguousarray(mz_arr, dtype=np.float64)
        int_arr = np.ascontiguousarray(int_arr, dtype=np.float64)
        self.L.refshim_consume_spectra(self.h, _p(mz_arr), _p(int_arr), mz_arr.size)
        cap = mz_arr.size
        b = np.zeros(cap, np.int32); r = np.zeros(cap, np.int32)
        m = np.zeros(cap, np.float64); it = np.zeros(cap, np.float64)
        lo = C.c_float(); hi = C.c_float(); nb = C.c_int64()
        n = self.L.refshim_binned(self.h, _p(b), _p(r), _p(m), _p(it), cap, C.byref(lo), C.byref(hi), C.byref(nb))
        return dict(bin=b[:n], rank=r[:n], mz=m[:n], intensity=it[:n], min_mz=lo.value, max_mz=hi.value,
                    n_bins=nb.value)

    def consume_peptide(self, peptide, n_of_mod, max_fragment_charge=1, aux_mod_pos=None, aux_mod_mass=None):
        if aux_mod_pos is not None:
            aux_mod_pos = np.ascontiguousarray(aux_mod_pos, dtype=np.uint32)
            aux_mod_mass = np.ascontiguousarray(aux_mod_mass, dtype=np.float32)
            na = aux_mod_pos.size
        else:
            na = 0
        self.L.refshim_consume_peptide(self.h, peptide.encode(), n_of_mod, max_fragment_charge,
                                       _p(aux_mod_pos) if na else None, _p(aux_mod_mass) if na else None, na)

    def fragment_graph(self, ftype, charge, S):
        nf = C.c_int64()
        n = self.L.refshim_fragment_graph(self.h, ftype.encode(), charge, None, 0, None, None, 0, C.byref(nf))
        sig = np.zeros((n, max(S, 1)), np.int32)
        off = np.zeros(n + 1, np.int64)
        fr = np.zeros(nf.value, np.float32)
        self.L.refshim_fragment_graph(self.h, ftype.encode(), charge, _p(sig), n, _p(off), _p(fr), nf.value,
                                      C.byref(nf))
        off[n] = nf.value
        return sig[:, :S], off, fr

    def site_determining(self, sig_a, sig_b, ftype, max_charge, cap=8192):
        sa = np.ascontiguousarray(sig_a, np.int32); sb = np.ascontiguousarray(sig_b, np.int32)
        oa = np.zeros(cap, np.float32); ob = np.zeros(cap, np.float32)
        na = c_long(); nb = c_long()
        self.L.refshim_site_determining(self.h, sa.size, _p(sa), _p(sb), ftype.encode(), max_charge,
                                        _p(oa), C.byref(na), _p(ob), C.byref(nb), cap)
        return oa[:na.value].copy(), ob[:nb.value].copy()

    def get_peptide(self, sig):
        s = np.ascontiguousarray(sig, np.int32)
        buf = C.create_string_buffer(4096)
        self.L.refshim_get_peptide(self.h, s.size, _p(s), buf, 4096)
        return buf.value.decode()

    def score_batch(self, batch, max_k=8, want_ascores=True):
        """batch: dict of CSR arrays as produced by pyascore_b200.batch (host numpy)."""
        n = int(batch["n_mod"].size)
        best = np.zeros(n, np.float32)
        asc = np.zeros((n, max_k), np.float32) if want_ascores else None
        aux_off = batch.get("aux_off")
        self.L.refshim_score_batch(self.h, n, _p(batch["spec_off"]), _p(batch["mz"]), _p(batch["inten"]),
                                   _p(batch["psm_spec"]), _p(batch["pep_off"]), _p(batch["pep"]),
                                   _p(batch["n_mod"]), _p(batch["max_charge"]), _p(aux_off),
                                   _p(batch.get("aux_pos")), _p(batch.get("aux_mass")), _p(best), _p(asc),
                                   max_k)
        return best, asc


class RefPyAscore(_CScorer):
    """The compiled, unmodified reference (oracle/_ref)."""
    PREFIX = "refshim_"


class OraclePyAscore(_CScorer):
    """The plain-C restatement (oracle/ascore_oracle.c)."""
    PREFIX = "orc_"

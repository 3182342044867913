"""Identification files -> PSM records `{scan, charge_state, score, peptide, mod_positions, mod_masses}`.

Mirror of the reference's `MassCorrector` (pyascore/parsing/id_parsers.py:22-180) and
`IdentificationParser` (:656-815): same constructor arguments, `to_list()` / `to_dict()`, record
schema and ordering (records sorted by the scan of their first hit, hits in file order).  The four
formats are read by the stdlib readers of `_xml.py` (pepXML, mzIdentML) and `csv` (percolator /
mokapot tab-separated output) instead of pyteomics / pandas.
"""
import csv
import re
import warnings

import numpy as np
from numpy import isclose

from . import _xml
from ._xml import STD_AA_MASS

# id_parsers.py:14-20
COMMON_MODS = {"n": 42.010565,
               "M": 15.9949,
               "K": 8.014199,
               "S": 79.966331,
               "T": 79.966331,
               "Y": 79.966331,
               "C": 57.021464}


class MassCorrector:
    """Un-round and de-combine modification masses reported by search engines
    (reference: id_parsers.py:22-130; the decision ladder of `correct` is :61-98)."""

    def __init__(self, mod_mass_dict=COMMON_MODS, aa_mass_dict=STD_AA_MASS, mz_tol=1.5, n_mod_ind=0):
        self.mod_mass_dict = mod_mass_dict
        self.aa_mass_dict = aa_mass_dict
        self.mz_tol = mz_tol
        self.n_mod_ind = n_mod_ind

    def correct(self, res, pos, mass):
        """-> (residues,), (positions,), (masses,): length 2 when an N-terminal mod was fused with a
        mod on the first residue, else length 1"""
        inf = np.inf
        std_mass = STD_AA_MASS.get(res, inf)
        mod_mass = self.mod_mass_dict.get(res, inf)
        n_mod_mass = self.mod_mass_dict.get('n', inf)
        tol = self.mz_tol
        if pos == 0 and isclose(mass, n_mod_mass, rtol=0., atol=tol):
            return ('n',), (self.n_mod_ind,), (n_mod_mass,)
        if pos == 1 and isclose(mass, std_mass + n_mod_mass, rtol=0., atol=tol):
            return ('n',), (self.n_mod_ind,), (n_mod_mass,)
        if pos == 1 and isclose(mass, std_mass + mod_mass + n_mod_mass, rtol=0., atol=tol):
            return ('n', res), (self.n_mod_ind, pos), (n_mod_mass, mod_mass)
        if isclose(mass, std_mass + mod_mass, rtol=0., atol=tol):
            return (res,), (pos,), (mod_mass,)
        pred_mod_mass = mass - STD_AA_MASS.get(res, 0.)
        warnings.warn("Unrecognized mod on {} at position {} with mass: {}"
                      " Using uncorrected mass.".format(res, pos, pred_mod_mass))
        return (res,), (pos,), (pred_mod_mass,)

    def correct_multiple(self, peptide, positions, masses):
        """correct every (position, mass) of one peptide -> (positions array, masses array)"""
        out_pos, out_mass = [], []
        for pos, mass in zip(positions, masses):
            res = 'n' if pos == 0 else peptide[pos - 1]
            _, p, m = self.correct(res, pos, mass)
            out_pos.extend(p)
            out_mass.extend(m)
        return np.array(out_pos), np.array(out_mass)


# ---------------------------------------------------------------------------------------------
# tab-separated percolator / mokapot PSM tables (reference: id_parsers.py:183-221, :475-653)
# ---------------------------------------------------------------------------------------------
_BRACKET = re.compile(r"\[([^A-Za-z\[\]]+)\]")
_RESIDUE = re.compile(r"([A-Z])(?:\[([^A-Za-z\[\]]+)\])?")
_NTERM = re.compile(r"n?\[([^A-Za-z\[\]]+)\]")


def _parse_bracket_sequence(seq, static_mods):
    """'n[42.01]PEPS[79.97]K' -> (peptide, positions i32[], masses f32[]) as the reference's
    percolator/mokapot extractors do: the N-terminal entry carries the bare mod mass, residue entries
    carry residue mass + delta, static mods are re-attached to unannotated residues"""
    pos, mass = [], []
    m = _NTERM.match(seq)
    if m is not None:
        mass.append(float(m.group(1)))
        pos.append(0)
        seq = seq[m.end():]
    else:
        if "n" in static_mods:
            mass.append(static_mods["n"])
            pos.append(0)
        if seq.startswith("n"):
            seq = seq[1:]
    pep = []
    for ind, r in enumerate(_RESIDUE.finditer(seq), 1):
        aa, delta = r.group(1), r.group(2)
        pep.append(aa)
        if delta is not None:
            mass.append(STD_AA_MASS[aa] + float(delta))
            pos.append(ind)
        elif aa in static_mods:
            mass.append(STD_AA_MASS[aa] + static_mods[aa])
            pos.append(ind)
    return "".join(pep), np.array(pos, dtype=np.int32), np.array(mass, dtype=np.float32)


def _iter_tsv(path, scan_col, charge_col, score_col, seq_col, strip_flanks, static_mods):
    groups = {}
    with open(path, newline="") as src:
        for row in csv.DictReader(src, delimiter="\t"):
            scan = int(float(row[scan_col]))
            groups.setdefault(scan, []).append(row)
    for scan in sorted(groups):                       # pandas groupby sorts the keys
        hits = []
        charge = None
        for row in groups[scan]:
            seq = row[seq_col]
            if strip_flanks:
                seq = re.sub(r"(^.\.)|(\..$)", "", seq)
            pep, pos, mass = _parse_bracket_sequence(seq, static_mods)
            ch = int(float(row[charge_col])) if charge_col else None
            hits.append((pep, float(row[score_col]), pos, mass, ch))
            charge = ch
        yield scan, charge, hits


class IdentificationParser:
    """Read PSMs from pepXML / mzIdentML / percolatorTXT / mokapotTXT (reference: id_parsers.py:656-815)."""

    def __init__(self, id_file_name, id_file_format, mass_corrector=None, score_string=None,
                 score_threshold=None, score_lower_better=True, score_func=None,
                 static_mods={"C": 57.021464}, spec_file_name=None):
        if id_file_format == "mzIdentML":
            self._source = lambda: _xml.iter_mzid(id_file_name, score_string)
        elif id_file_format == "pepXML":
            self._source = lambda: _xml.iter_pepxml(id_file_name, score_string)
        elif id_file_format == "percolatorTXT":
            self._source = lambda: _iter_tsv(id_file_name, "scan", "charge", "percolator score", "sequence",
                                             False, static_mods)
        elif id_file_format == "mokapotTXT":
            self._source = lambda: _iter_tsv(id_file_name, "ScanNr", None, "mokapot score", "Peptide",
                                             True, static_mods)
        else:
            raise ValueError("{} not supported at this time."
                             " Must be on of: mzIdentML, pepXML,"
                             " percolatorTXT, or mokapotTXT".format(id_file_format))
        self.mass_corrector = mass_corrector if mass_corrector is not None else MassCorrector()
        self.score_threshold = score_threshold
        self.score_lower_better = score_lower_better
        self.score_func = score_func
        self.spec_file_name = spec_file_name
        self._match_records = []

    def _get_match_records(self):
        if not self._match_records:
            recs = [r for r in self._source() if len(r[2]) > 0]
            self._match_records = sorted(recs, key=lambda r: r[0])      # stable, like the reference

    def _passes_scoring(self, score):
        if self.score_threshold is None:
            return True
        if score is None:
            return False
        # id_parsers.py:767: `(-1 ** flag)` parses as -(1 ** flag) == -1 whatever the flag, so the
        # reference keeps scores BELOW the threshold in both settings; kept for drop-in behaviour
        return (score - self.score_threshold) * (-1 ** self.score_lower_better) > 0

    def _generate_hits(self):
        self._get_match_records()
        for scan, charge, hits in self._match_records:
            for hit in hits:
                pep, score, pos, mass = hit[:4]
                hit_charge = hit[4] if len(hit) > 4 else charge
                mod_positions, mod_masses = self.mass_corrector.correct_multiple(pep, pos, mass)
                if self.score_func is not None and score is not None:
                    score = self.score_func(score)
                if not self._passes_scoring(score):
                    continue
                yield {"scan": scan, "charge_state": hit_charge, "score": score, "peptide": pep,
                       "mod_positions": mod_positions, "mod_masses": mod_masses}

    def to_list(self):
        """PSMs of the file, sorted by scan number"""
        return [hit for hit in self._generate_hits()]

    def to_dict(self):
        """{scan number : PSM} (the last hit of a scan wins, as in the reference)"""
        return {hit.pop("scan"): hit for hit in self._generate_hits()}

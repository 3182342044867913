"""Streaming stdlib readers for the four XML formats the reference reads through pyteomics
(pyascore/parsing/spec_parsers.py:3-4, id_parsers.py:6-7).  pyteomics/lxml are third-party
and absent here; these readers use `xml.etree.ElementTree.iterparse` + `base64`/`zlib`/numpy and
yield only the fields the scoring path consumes, already normalised:

    iter_mzml / iter_mzxml   -> (scan, ms_level, precursor_mz, precursor_charge, mz f64[], inten f64[])
    iter_pepxml / iter_mzid  -> (scan, charge, [hit, ...]),  hit = (peptide, score, positions i32[], masses f32[])

Field semantics follow the reference's extractors (cited per function); where the reference
would crash on a missing key caught nowhere, the value its `except KeyError` branch yields is used.
"""
import base64
import re
import zlib
import xml.etree.ElementTree as ET

import numpy as np

_SCAN_RE = re.compile(r"(?<=scan=)[0-9]+")
_NUM_RE = re.compile(r"[0-9]+")

# residue masses of pyteomics.mass.std_aa_mass (monoisotopic; the published table the reference
# imports at id_parsers.py:11) -- same values as cpp/Types.h:7-30 plus the two termini
STD_AA_MASS = {
    'G': 57.02146, 'A': 71.03711, 'S': 87.03203, 'P': 97.05276, 'V': 99.06841, 'T': 101.04768,
    'C': 103.00919, 'L': 113.08406, 'I': 113.08406, 'N': 114.04293, 'D': 115.02694, 'Q': 128.05858,
    'K': 128.09496, 'E': 129.04259, 'M': 131.04049, 'H': 137.05891, 'F': 147.06841, 'U': 150.95364,
    'R': 156.10111, 'Y': 163.06333, 'W': 186.07931, 'O': 237.14773, 'H-': 1.00783, '-OH': 17.00274}


def _local(tag):
    return tag.rsplit('}', 1)[-1]


def _children(el, name):
    return [c for c in el if _local(c.tag) == name]


def _child(el, name):
    for c in el:
        if _local(c.tag) == name:
            return c
    return None


def _number(text):
    """pyteomics-style value conversion: int when it parses as one, else float, else the string"""
    try:
        return int(text)
    except (TypeError, ValueError):
        pass
    try:
        return float(text)
    except (TypeError, ValueError):
        return text


# ---------------------------------------------------------------------------------------------
# mzML (reference: MzMLExtractor, spec_parsers.py:49-110)
# ---------------------------------------------------------------------------------------------
_MZML_DTYPES = {"MS:1000523": "<f8", "MS:1000521": "<f4", "MS:1000519": "<i4", "MS:1000522": "<i8"}
_MZML_NUMPRESS = {"MS:1002312", "MS:1002313", "MS:1002314", "MS:1002746", "MS:1002747", "MS:1002748"}


def _decode_mzml_array(bda, groups):
    dtype, zl, kind = None, False, None
    params = list(_children(bda, "cvParam"))
    for ref in _children(bda, "referenceableParamGroupRef"):
        params.extend(groups.get(ref.get("ref"), ()))
    for p in params:
        acc = p.get("accession")
        if acc in _MZML_DTYPES:
            dtype = _MZML_DTYPES[acc]
        elif acc == "MS:1000574":
            zl = True
        elif acc in _MZML_NUMPRESS:
            raise ValueError("mzML numpress compression (%s) is not supported by the stdlib reader" % acc)
        elif acc == "MS:1000514":
            kind = "mz"
        elif acc == "MS:1000515":
            kind = "inten"
    if kind is None or dtype is None:
        return None, None
    b = _child(bda, "binary")
    raw = base64.b64decode(b.text or "") if b is not None else b""
    if zl and raw:
        raw = zlib.decompress(raw)
    return kind, np.frombuffer(raw, dtype=dtype).astype(np.float64)


def iter_mzml(path):
    groups = {}
    for ev, el in ET.iterparse(path, events=("end",)):
        name = _local(el.tag)
        if name == "referenceableParamGroup":
            groups[el.get("id")] = _children(el, "cvParam")
            continue
        if name != "spectrum":
            continue
        m = _SCAN_RE.search(el.get("id", ""))
        scan = int(m.group()) if m else -1
        ms_level = 0
        params = list(_children(el, "cvParam"))
        for ref in _children(el, "referenceableParamGroupRef"):
            params.extend(groups.get(ref.get("ref"), ()))
        for p in params:
            if p.get("accession") == "MS:1000511":
                ms_level = int(p.get("value"))
        prec_mz = prec_z = None
        pl = _child(el, "precursorList")
        if pl is not None:
            precs = _children(pl, "precursor")
            if int(pl.get("count", len(precs))) > 1:
                raise ValueError("Multiple precursors not supported at this time")
            try:
                ion = _children(_child(precs[0], "selectedIonList"), "selectedIon")[0]
                vals = {p.get("accession"): p.get("value") for p in _children(ion, "cvParam")}
                # both keys must exist, as in the reference's single try block (spec_parsers.py:85-96)
                prec_mz, prec_z = float(vals["MS:1000744"]), int(vals["MS:1000041"])
            except (KeyError, IndexError, TypeError):
                prec_mz = prec_z = None
        mz = inten = None
        bl = _child(el, "binaryDataArrayList")
        if bl is not None:
            for bda in _children(bl, "binaryDataArray"):
                kind, arr = _decode_mzml_array(bda, groups)
                if kind == "mz":
                    mz = arr
                elif kind == "inten":
                    inten = arr
        if mz is None or inten is None:
            mz, inten = np.zeros(0, np.float64), np.zeros(0, np.float64)
        yield scan, ms_level, prec_mz, prec_z, mz, inten
        el.clear()


# ---------------------------------------------------------------------------------------------
# mzXML (reference: MzXMLExtractor, spec_parsers.py:113-172)
# ---------------------------------------------------------------------------------------------
def iter_mzxml(path):
    for ev, el in ET.iterparse(path, events=("end",)):
        if _local(el.tag) != "scan":
            continue
        try:
            scan = int(el.get("num"))
        except (TypeError, ValueError):
            scan = -1
        try:
            ms_level = int(el.get("msLevel"))
        except (TypeError, ValueError):
            ms_level = 0
        prec_mz = prec_z = None
        precs = _children(el, "precursorMz")
        if len(precs) > 1:
            raise ValueError("Multiple precursors not supported at this time")
        if precs and precs[0].get("precursorCharge") is not None:
            prec_mz, prec_z = float(precs[0].text), int(precs[0].get("precursorCharge"))
        mz, inten = np.zeros(0, np.float64), np.zeros(0, np.float64)
        pk = _child(el, "peaks")
        if pk is not None and pk.text and pk.text.strip():
            raw = base64.b64decode(pk.text)
            if pk.get("compressionType", "none") == "zlib":
                raw = zlib.decompress(raw)
            dt = (">f8" if pk.get("precision", "32") == "64" else ">f4")
            pairs = np.frombuffer(raw, dtype=dt)
            mz, inten = pairs[0::2].astype(np.float64), pairs[1::2].astype(np.float64)
        yield scan, ms_level, prec_mz, prec_z, mz, inten
        # nested scans (MS2 inside MS1) have been yielded already; drop the payload only
        if pk is not None:
            pk.text = None


# ---------------------------------------------------------------------------------------------
# pepXML (reference: PepXMLExtractor, id_parsers.py:388-472, on pyteomics' search_hit dicts)
# ---------------------------------------------------------------------------------------------
def iter_pepxml(path, score_string=None):
    for ev, el in ET.iterparse(path, events=("end",)):
        if _local(el.tag) != "spectrum_query":
            continue
        try:
            scan = int(el.get("start_scan"))
        except (TypeError, ValueError):
            scan = -1
        try:
            charge = int(el.get("assumed_charge"))
        except (TypeError, ValueError):
            charge = 0
        hits = []
        for sr in _children(el, "search_result"):
            for h in _children(sr, "search_hit"):
                pep = h.get("peptide", "")
                pos, mass = [], []
                # a hit may carry several <modification_info> blocks (Crux writes one per mod kind);
                # pyteomics keeps the last one, which test/test_id_parsers.py:192-193 pins
                mis = _children(h, "modification_info")
                mi = mis[-1] if mis else None
                if mi is not None:
                    # pyteomics folds the terminal attributes into the modification list
                    if mi.get("mod_nterm_mass") is not None:
                        pos.append(0)
                        mass.append(float(mi.get("mod_nterm_mass")))
                    for m in _children(mi, "mod_aminoacid_mass"):
                        pos.append(int(m.get("position")))
                        mass.append(float(m.get("mass")))
                    if mi.get("mod_cterm_mass") is not None:
                        pos.append(len(pep) + 1)
                        mass.append(float(mi.get("mod_cterm_mass")))
                score = None
                if score_string is not None:
                    for sc in _children(h, "search_score"):
                        if sc.get("name") == score_string:
                            score = float(sc.get("value"))
                hits.append((pep, score, np.array(pos, np.int32), np.array(mass, np.float32)))
        yield scan, charge, hits
        el.clear()


# ---------------------------------------------------------------------------------------------
# mzIdentML (reference: MzIdentMLExtractor, id_parsers.py:286-385, on pyteomics' items with
# the referenced <Peptide> merged in)
# ---------------------------------------------------------------------------------------------
def iter_mzid(path, score_string=None):
    peptides = {}
    for ev, el in ET.iterparse(path, events=("end",)):
        name = _local(el.tag)
        if name == "Peptide":
            seq = _child(el, "PeptideSequence")
            mods = [(int(m.get("location")), m.get("residues"), float(m.get("monoisotopicMassDelta")))
                    for m in _children(el, "Modification")]
            peptides[el.get("id")] = (seq.text if seq is not None else "", mods)
            el.clear()
            continue
        if name != "SpectrumIdentificationResult":
            continue
        sid = el.get("spectrumID")
        scan = -1
        if sid is not None:
            m = _SCAN_RE.search(sid) or _NUM_RE.search(sid)
            scan = int(m.group())
        hits = []
        charge = 0
        for it in _children(el, "SpectrumIdentificationItem"):
            try:
                charge_i = int(it.get("chargeState"))
            except (TypeError, ValueError):
                charge_i = 0
            pep, mods = peptides.get(it.get("peptide_ref"), ("", []))
            pos = np.zeros(len(mods), np.int32)
            mass = np.zeros(len(mods), np.float32)
            for i, (loc, residues, delta) in enumerate(mods):
                pos[i] = loc
                aa = residues[0] if residues else ("n" + pep + "c")[loc]
                mass[i] = STD_AA_MASS.get(aa, 0.) + delta
            score = None
            if score_string is not None:
                for p in list(_children(it, "cvParam")) + list(_children(it, "userParam")):
                    if p.get("name") == score_string:
                        score = float(p.get("value"))
            hits.append((pep, score, pos, mass, charge_i))
            charge = charge_i
        yield scan, charge, hits
        el.clear()

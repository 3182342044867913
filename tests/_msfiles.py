"""Test-side writers of minimal mzML / mzXML / pepXML / mzIdentML / percolator / mokapot files.

The reference's example inputs live under /root/reference/test/example_inputs and do not travel to
the GPU box; tests therefore materialise small files from arrays (seeded synthetic data or the
committed config-1 golden) and read them back through pyascore_b200.parsing.
"""
import base64
import zlib

import numpy as np

from pyascore_b200.parsing._xml import STD_AA_MASS


def _b64(arr, dtype, compress):
    raw = np.asarray(arr).astype(dtype).tobytes()
    if compress:
        raw = zlib.compress(raw)
    return base64.b64encode(raw).decode()


def write_mzml(path, spectra, compress=False, inten_bits=32, ms1_every=0, indexed=True):
    """spectra: list of dict(scan, precursor_mz, precursor_charge, mz, inten[, ms_level])"""
    out = ['<?xml version="1.0" encoding="utf-8"?>']
    if indexed:
        out.append('<indexedmzML xmlns="http://psi.hupo.org/ms/mzml">')
    out.append('<mzML xmlns="http://psi.hupo.org/ms/mzml" id="t" version="1.1.0">')
    out.append('<referenceableParamGroupList count="1"><referenceableParamGroup id="g64">'
               '<cvParam cvRef="MS" accession="MS:1000523" name="64-bit float" value=""/>'
               '</referenceableParamGroup></referenceableParamGroupList>')
    out.append('<run id="r"><spectrumList count="%d">' % len(spectra))
    comp = ('<cvParam cvRef="MS" accession="MS:1000574" name="zlib compression" value=""/>' if compress else
            '<cvParam cvRef="MS" accession="MS:1000576" name="no compression" value=""/>')
    idt = '<cvParam cvRef="MS" accession="%s" name="%d-bit float" value=""/>' % (
        "MS:1000521" if inten_bits == 32 else "MS:1000523", inten_bits)
    for i, s in enumerate(spectra):
        lvl = s.get("ms_level", 2)
        out.append('<spectrum index="%d" id="controllerType=0 controllerNumber=1 scan=%d" defaultArrayLength="%d">'
                   % (i, s["scan"], len(s["mz"])))
        out.append('<cvParam cvRef="MS" accession="MS:1000511" name="ms level" value="%d"/>' % lvl)
        if lvl > 1 and s.get("precursor_mz") is not None:
            out.append('<precursorList count="1"><precursor><selectedIonList count="1"><selectedIon>'
                       '<cvParam cvRef="MS" accession="MS:1000744" name="selected ion m/z" value="%r"/>' % float(s["precursor_mz"]))
            if s.get("precursor_charge"):
                out.append('<cvParam cvRef="MS" accession="MS:1000041" name="charge state" value="%d"/>' % s["precursor_charge"])
            out.append('</selectedIon></selectedIonList></precursor></precursorList>')
        out.append('<binaryDataArrayList count="2"><binaryDataArray><referenceableParamGroupRef ref="g64"/>' + comp +
                   '<cvParam cvRef="MS" accession="MS:1000514" name="m/z array" value=""/><binary>%s</binary></binaryDataArray>'
                   % _b64(s["mz"], "<f8", compress))
        out.append('<binaryDataArray>' + idt + comp +
                   '<cvParam cvRef="MS" accession="MS:1000515" name="intensity array" value=""/><binary>%s</binary>'
                   '</binaryDataArray></binaryDataArrayList></spectrum>'
                   % _b64(s["inten"], "<f4" if inten_bits == 32 else "<f8", compress))
    out.append('</spectrumList></run></mzML>')
    if indexed:
        out.append('<indexList count="0"/></indexedmzML>')
    with open(path, "w") as f:
        f.write("\n".join(out))


def write_mzxml(path, spectra, compress=False, bits=64):
    out = ['<?xml version="1.0" encoding="ISO-8859-1"?>',
           '<mzXML xmlns="http://sashimi.sourceforge.net/schema_revision/mzXML_3.2"><msRun scanCount="%d">' % len(spectra)]
    for s in spectra:
        pairs = np.empty(2 * len(s["mz"]), np.float64)
        pairs[0::2], pairs[1::2] = s["mz"], s["inten"]
        out.append('<scan num="%d" msLevel="%d" peaksCount="%d">' % (s["scan"], s.get("ms_level", 2), len(s["mz"])))
        if s.get("precursor_mz") is not None:
            z = ' precursorCharge="%d"' % s["precursor_charge"] if s.get("precursor_charge") else ""
            out.append('<precursorMz precursorIntensity="1"%s>%r</precursorMz>' % (z, float(s["precursor_mz"])))
        out.append('<peaks compressionType="%s" compressedLen="0" precision="%d" byteOrder="network" '
                   'contentType="m/z-int">%s</peaks></scan>'
                   % ("zlib" if compress else "none", bits, _b64(pairs, ">f8" if bits == 64 else ">f4", compress)))
    out.append('</msRun></mzXML>')
    with open(path, "w") as f:
        f.write("\n".join(out))


def write_pepxml(path, queries, score_name="xcorr_score"):
    """queries: list of dict(scan, charge, hits=[dict(peptide, score, mods=[(pos, total_mass)],
    nterm=None|mass)]) -- masses as pepXML reports them (residue + modification)"""
    out = ['<?xml version="1.0" encoding="UTF-8"?>',
           '<msms_pipeline_analysis xmlns="http://regis-web.systemsbiology.net/pepXML"><msms_run_summary base_name="NA">']
    for i, q in enumerate(queries, 1):
        out.append('<spectrum_query spectrum="t.%d.%d.%d" start_scan="%d" end_scan="%d" assumed_charge="%d" index="%d">'
                   '<search_result>' % (q["scan"], q["scan"], q["charge"], q["scan"], q["scan"], q["charge"], i))
        for r, h in enumerate(q["hits"], 1):
            out.append('<search_hit hit_rank="%d" peptide="%s" protein="P">' % (r, h["peptide"]))
            if h.get("mods") or h.get("nterm") is not None:
                nt = ' mod_nterm_mass="%s"' % h["nterm"] if h.get("nterm") is not None else ""
                out.append('<modification_info%s>' % nt)
                for pos, mass in h.get("mods", []):
                    out.append('<mod_aminoacid_mass position="%d" mass="%s"/>' % (pos, mass))
                out.append('</modification_info>')
            out.append('<search_score name="%s" value="%s"/></search_hit>' % (score_name, h["score"]))
        out.append('</search_result></spectrum_query>')
    out.append('</msms_run_summary></msms_pipeline_analysis>')
    with open(path, "w") as f:
        f.write("\n".join(out))


def write_mzid(path, queries, score_name="SEQUEST:xcorr"):
    """same `queries` as write_pepxml; mods are written as mass deltas (total - residue mass)"""
    peps, items = [], []
    for q in queries:
        its = []
        for r, h in enumerate(q["hits"], 1):
            pid = "PEP_%d" % len(peps)
            mods = []
            if h.get("nterm") is not None:
                mods.append('<Modification location="0" monoisotopicMassDelta="%s"/>' % h["nterm"])
            for j, (pos, mass) in enumerate(h.get("mods", [])):
                aa = h["peptide"][pos - 1]
                res = ' residues="%s"' % aa if j % 2 == 0 else ""      # the attribute is optional: cover both
                mods.append('<Modification location="%d"%s monoisotopicMassDelta="%.6f"/>' % (pos, res, float(mass) - STD_AA_MASS[aa]))
            peps.append('<Peptide id="%s"><PeptideSequence>%s</PeptideSequence>%s</Peptide>' % (pid, h["peptide"], "".join(mods)))
            its.append('<SpectrumIdentificationItem id="SII_%d" rank="%d" chargeState="%d" peptide_ref="%s">'
                       '<cvParam cvRef="MS" accession="MS:1001155" name="%s" value="%s"/></SpectrumIdentificationItem>'
                       % (len(peps), r, q["charge"], pid, score_name, h["score"]))
        items.append('<SpectrumIdentificationResult id="SIR_%d" spectrumID="controllerType=0 controllerNumber=1 scan=%d">%s'
                     '</SpectrumIdentificationResult>' % (q["scan"], q["scan"], "".join(its)))
    with open(path, "w") as f:
        f.write('<?xml version="1.0" encoding="UTF-8"?>\n<MzIdentML xmlns="http://psidev.info/psi/pi/mzIdentML/1.1">'
                '<SequenceCollection>%s</SequenceCollection><DataCollection><AnalysisData>'
                '<SpectrumIdentificationList id="SIL">%s</SpectrumIdentificationList></AnalysisData></DataCollection>'
                '</MzIdentML>' % ("\n".join(peps), "\n".join(items)))


def bracket_sequence(h, fmt="%.4f"):
    """'n[42.0106]PEPS[79.9663]K' annotation (mass deltas) of a hit"""
    mods = dict(h.get("mods", []))
    s = ("n[" + fmt % float(h["nterm"]) + "]") if h.get("nterm") is not None else ""
    for i, aa in enumerate(h["peptide"], 1):
        s += aa
        if i in mods:
            s += "[" + fmt % (float(mods[i]) - STD_AA_MASS[aa]) + "]"
    return s


def write_percolator_txt(path, queries):
    with open(path, "w") as f:
        f.write("file_idx\tscan\tcharge\tpercolator score\tsequence\n")
        for q in queries:
            for h in q["hits"]:
                f.write("0\t%d\t%d\t%s\t%s\n" % (q["scan"], q["charge"], h["score"], bracket_sequence(h)))


def write_mokapot_txt(path, queries):
    with open(path, "w") as f:
        f.write("SpecId\tLabel\tScanNr\tmokapot score\tPeptide\n")
        for q in queries:
            for h in q["hits"]:
                f.write("s\tTrue\t%d\t%s\tK.%s.A\n" % (q["scan"], h["score"], bracket_sequence(h)))

"""Command-line options of `python -m pyascore_b200`: the reference's flag set
(pyascore/config.py:19-93: same names, types and defaults, so parameter files and scripts carry
over) plus `--device` / `--chunk_psms` for the GPU pipeline."""
import argparse
import re

# (flag, type, default, help) -- one row per reference option (config.py:31-91)
_OPTIONS = [
    ("--residues", str, "STY", "Residues that can carry the modification."),
    ("--mod_mass", float, 79.966331, "Exact modification mass (search engines often report it rounded)."),
    ("--mz_error", float, 0.5, "m/z tolerance for matching a spectral peak to a theoretical fragment."),
    ("--mod_correction_tol", float, 1., "Mass tolerance for recognising a reported modification as the "
                                        "variable one (wide values absorb rounding)."),
    # type=bool as in the reference (config.py:49): any non-empty string enables it
    ("--zero_based", bool, False, "Reported mod positions are 0-based (default: 1-based)."),
    ("--neutral_loss_groups", str, "", "Comma separated residue groups with a neutral loss; lower case = "
                                       "the modified residue loses it ('st' vs 'ST')."),
    ("--neutral_loss_masses", str, "", "Comma separated loss masses, one per group (negative = gain)."),
    ("--static_mod_groups", str, "C", "Comma separated residue groups that always carry a fixed modification."),
    ("--static_mod_masses", str, "57.021464", "Comma separated masses, one per static group."),
    ("--fragment_types", str, "by", "Fragment ion series to score, any of b c y z Z (Z = z+H)."),
    ("--max_fragment_charge", int, 5, "Upper limit of the fragment charge (never above precursor charge - 1)."),
    ("--hit_depth", int, 1, "PSMs taken per scan; negative = all."),
    ("--parameter_file", str, "", "File of 'name = value' lines; command-line flags override it."),
    ("--spec_file_type", str, "mzML", "mzML or mzXML."),
    ("--ident_file_type", str, "pepXML", "pepXML, mzIdentML, percolatorTXT or mokapotTXT."),
]


def args_from_file(file_path):
    """'name = value' lines ('#' starts a comment) -> ['--name', 'value', ...] (config.py:5-17)"""
    out = []
    with open(file_path, "r") as src:
        for line in src:
            m = re.search(r"^([^\s]+)\s*=\s*([^\s]+)$", line.split("#")[0].strip())
            if m is not None:
                out += ["--" + m.group(1), m.group(2)]
    return out


def build_parser():
    p = argparse.ArgumentParser(
        prog="pyascore_b200",
        description="PTM localisation with the Ascore algorithm, scored in batches on a B200 GPU. "
                    "Reads spectra (mzML/mzXML) and identifications (pepXML/mzIdentML/percolator/mokapot), "
                    "writes one row per scored PSM.")
    p.add_argument("--match_save", action="store_true",
                   help="Dump the inputs of the scored PSMs (spectra CSR + PSM arrays) to dump_batch.npz.")
    for flag, typ, default, text in _OPTIONS:
        p.add_argument(flag, type=typ, default=default, help=text)
    p.add_argument("--device", type=int, default=0, help="CUDA device ordinal.")
    p.add_argument("--devices", type=str, default="", help="Comma separated CUDA device ordinals: every batch is "
                   "sharded over these GPUs (overrides --device).")
    p.add_argument("--chunk_psms", type=int, default=65536, help="PSMs packed per GPU batch.")
    p.add_argument("spec_file", type=str, help="MS spectra file.")
    p.add_argument("ident_file", type=str, help="Results of the database search.")
    p.add_argument("out_file", type=str, help="Destination of the tab-separated results.")
    return p


def validate_args(args):
    """reference: __main__.py:48-65"""
    allowed_residues = "ncACDEFGHIKLMNOPQRSTUVWY"
    for aa in args.residues:
        if aa not in allowed_residues:
            raise ValueError("The residue inputed, {}, is not allowed."
                             " Must be one of: {}".format(aa, allowed_residues))
    allowed_fragments = "cbyzZ"
    for frag in args.fragment_types:
        if frag not in allowed_fragments:
            raise ValueError("The fragment type inputed, {}, is not allowed."
                             " Must be one of: {}".format(frag, allowed_fragments))
    if args.max_fragment_charge < 1:
        raise ValueError("The max fragment charge must be greater than or equal to 1")

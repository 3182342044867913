"""`python -m pyascore_b200 spec_file ident_file out_file [options]` -- the reference's CLI
(pyascore/__main__.py:114-173) with the per-PSM Python loop replaced by batched GPU scoring:

    parse spectra -> one pinned CSR block        (SpectraParser.to_csr)
    parse identifications, sort by scan          (IdentificationParser.to_list)
    pack chunks of PSMs (hit_depth, mod split, charge rule)  ||  score previous chunk on the GPU
    format rows, write the TSV (same columns and number formatting)
"""
import sys
from datetime import datetime

import numpy as np

from .batch import MultiScorer, Scorer
from .config import args_from_file, build_parser, validate_args
from .parsing import (COMMON_MODS, IdentificationParser, MassCorrector, SpectraParser, iter_batches, result_rows,
                      score_stream, write_tsv)


def _stamp():
    return datetime.now().strftime("%m/%d/%y %H:%M:%S")


def static_mod_dict(args):
    """reference: __main__.py:26-29"""
    out = {}
    for group, mass in zip(args.static_mod_groups.split(","), args.static_mod_masses.split(",")):
        out.update({aa: float(mass) for aa in group})
    return out


def parse_identifications(args):
    """reference: __main__.py:22-45 -> scan-sorted PSM records"""
    static_mods = static_mod_dict(args)
    mods = COMMON_MODS.copy()
    mods.update({aa: args.mod_mass for aa in args.residues})
    mods.update(static_mods)
    parser = IdentificationParser(args.ident_file, args.ident_file_type, MassCorrector(mod_mass_dict=mods),
                                  static_mods=static_mods)
    return sorted(parser.to_list(), key=lambda m: m["scan"])


def build_scorer(args):
    """reference: __main__.py:68-80 (bin_size 100, n_top 10 are fixed there too)"""
    devices = [int(d) for d in getattr(args, "devices", "").split(",") if d.strip() != ""]
    if devices:
        scorer = MultiScorer(100., 10, args.residues, args.mod_mass, args.mz_error, args.fragment_types, devices=devices)
    else:
        scorer = Scorer(100., 10, args.residues, args.mod_mass, args.mz_error, args.fragment_types, device=args.device)
    if args.neutral_loss_groups and args.neutral_loss_masses:
        for g, m in zip(args.neutral_loss_groups.split(","), args.neutral_loss_masses.split(",")):
            scorer.add_neutral_loss(g, float(m))
    return scorer


def run(args, log=print):
    log("{} -- Ascore Started".format(_stamp()))
    log("{} -- Reading spectra from: {}".format(_stamp(), args.spec_file))
    spectra = SpectraParser(args.spec_file, args.spec_file_type).to_csr()
    log("{} -- Reading identifications from: {}".format(_stamp(), args.ident_file))
    psms = parse_identifications(args)
    log("{} -- Anlyzing PSMs".format(_stamp()))
    scorer = build_scorer(args)
    batches = iter_batches(spectra, psms, chunk_psms=args.chunk_psms, residues=args.residues,
                           mod_mass=args.mod_mass, mod_correction_tol=args.mod_correction_tol,
                           zero_based=args.zero_based, max_fragment_charge=args.max_fragment_charge,
                           hit_depth=args.hit_depth)
    rows = []
    saved = []
    n_psm = 0
    try:
        for scans, batch, res in score_stream(scorer, batches):
            rows.extend(result_rows(scorer, scans, batch, res))
            n_psm += int(scans.size)
            if args.match_save:
                saved.append((scans, batch))
    finally:
        scorer.close()
    if args.match_save and saved:
        scans, batch = saved[-1]
        np.savez("dump_batch.npz", scans=scans, **{k: v for k, v in batch.items() if v is not None})
    if len(rows) != n_psm:
        # the reference would have written a row for these (or crashed on them): make the difference visible
        log("{} -- {} of {} PSMs could not be scored (see the warnings) and have no row in the output".format(
            _stamp(), n_psm - len(rows), n_psm))
    write_tsv(args.out_file, rows)
    log("{} -- Ascore Completed".format(_stamp()))
    return rows


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    parser = build_parser()
    args = parser.parse_args(argv)
    if args.parameter_file:
        args = parser.parse_args(args_from_file(args.parameter_file) + argv)
    validate_args(args)
    run(args)


if __name__ == "__main__":
    main()

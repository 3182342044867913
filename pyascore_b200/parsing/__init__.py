"""File parsing and CSR packing around the scoring path.

Same public names as the reference's `pyascore.parsing` (`SpectraParser`, `IdentificationParser`,
`MassCorrector`), implemented on the standard library because pyteomics/lxml are third-party."""
from .id_parsers import COMMON_MODS, IdentificationParser, MassCorrector
from .packer import PsmPacker, iter_batches, process_mods, result_rows, score_stream, write_tsv
from .spec_parsers import SpectraCSR, SpectraParser

__all__ = ["COMMON_MODS", "IdentificationParser", "MassCorrector", "PsmPacker", "SpectraCSR", "SpectraParser",
           "iter_batches", "process_mods", "result_rows", "score_stream", "write_tsv"]

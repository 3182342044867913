"""Small batch utilities shared by the GPU tests."""
import numpy as np

from pyascore_b200 import synth


def repeat_batch(batch, reps):
    """`reps` copies of a CSR batch, concatenated (offsets rebased) -> (batch, reps)"""
    parts = [{k: v for k, v in batch.items() if k != "mod_off"} for _ in range(reps)]
    return synth.concat_batches(parts), reps

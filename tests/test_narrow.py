"""CPU: the host-side m/z narrowing helper (pa_narrow_mz).  A spectrum may go over the host link as float32 only if the
kernels' view of it is provably unchanged: same bounds (cpp/Spectra.cpp:46-48) and the same bin for every peak
(:58-60).  Checked here against a straightforward numpy evaluation of both formulas: no spectrum that needs its
float64 values may be missed, and (almost) none may be flagged without need."""
import numpy as np
import pytest

from pyascore_b200 import _lib, synth


def narrow(mz, off, bin_size):
    L = _lib.load()
    out = np.zeros(mz.size, np.float32)
    flag = np.zeros(off.size - 1, np.uint8)
    n = L.pa_narrow_mz(mz.ctypes.data, off.ctypes.data, off.size - 1, bin_size, out.ctypes.data, flag.ctypes.data)
    assert n == int(flag.sum())
    return out, flag.astype(bool)


def needs_exact(mz, off, bin_size):
    """per spectrum: bounds or some bin differ between the float64 values and their float32 roundings"""
    n = off.size - 1
    need = np.zeros(n, bool)
    bs = np.float64(np.float32(bin_size))
    for s in range(n):
        m = mz[off[s]:off[s + 1]]
        if m.size == 0:
            continue
        w = m.astype(np.float32).astype(np.float64)
        if not np.all(np.isfinite(m)):
            need[s] = True
            continue
        lo, hi = np.float32(np.floor(m.min() / 100.) * 100.), np.float32(np.ceil(m.max() / 100.) * 100.)
        lo2, hi2 = np.float32(np.floor(w.min() / 100.) * 100.), np.float32(np.ceil(w.max() / 100.) * 100.)
        if lo != lo2 or hi != hi2:
            need[s] = True
            continue
        d = np.float64(lo)
        need[s] = bool(np.any(np.floor((m - d) / bs) != np.floor((w - d) / bs)))
    return need


@pytest.mark.parametrize("bin_size", [100., 50., 7.3])
def test_narrowing_is_sound_on_synthetic_spectra(bin_size):
    b = synth.make_batch("lowres_phospho", 3000, seed=21)
    out, flag = narrow(b["mz"], b["spec_off"], bin_size)
    assert np.array_equal(out, b["mz"].astype(np.float32))
    need = needs_exact(b["mz"], b["spec_off"], bin_size)
    assert not np.any(need & ~flag)                         # nothing that needs its float64 values is missed
    assert int((flag & ~need).sum()) <= 3 and flag.mean() < 0.05


def test_narrowing_flags_boundary_and_broken_spectra():
    rng = np.random.default_rng(2)
    specs = []
    base = np.sort(rng.uniform(150., 1900., 200))
    specs.append(base)                                                        # ordinary
    specs.append(np.sort(np.append(base, np.nextafter(700., 0.))))            # just below a bin boundary: (float) rounds it up to 700
    specs.append(np.sort(np.append(base, 700.)))                              # exactly on it: float32 keeps it there
    specs.append(np.sort(np.append(base, np.nextafter(1999.99999999, 3000.))))
    specs.append(np.append(base[:50], np.nan))                                # not finite
    specs.append(np.array([np.nextafter(300., 0.), 450.]))                    # minimum just below a multiple of 100: bounds move
    specs.append(base[::-1].copy())                                           # unsorted: narrowing is per peak, order is irrelevant
    specs.append(np.zeros(0))
    off = np.zeros(len(specs) + 1, np.int64)
    np.cumsum([s.size for s in specs], out=off[1:])
    mz = np.concatenate(specs)
    out, flag = narrow(mz, off, 100.)
    need = needs_exact(mz, off, 100.)
    assert not np.any(need & ~flag)
    assert list(flag) == [False, True, False, bool(need[3]) or flag[3], True, True, False, False]

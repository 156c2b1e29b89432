"""TEST INFRASTRUCTURE ONLY (oracle).  Finite-difference response functions formed from the reference's own spectra.

Base: benchmark column 1 (fixture synth70_c1).  For each parameter row p (T, v_z, B, gamma, chi) and a few depth
points k, the unmodified rhf1d() is run on the column with atmosphere[p][k] +- delta; stored is
(S+ - S-) / (2 delta) computed in float64 exactly as a pyrh caller would.  Output: tests/golden/rf_fd.npz.
Usage: python -m oracle.gen_golden_rf_fd
"""
import numpy as np

from oracle import refdriver as rd
from oracle.gen_golden import GOLD

ROWS = np.array([1, 3, 5, 6, 7], dtype=np.int32)            # T [K], v_z [km/s], B [G], gamma [rad], chi [rad]
DELTA = np.array([1.0, 0.01, 1.0, 1.0e-3, 1.0e-3])
DEPTHS = np.array([12, 30, 41, 52, 69], dtype=np.int32)


def main():
    g = np.load(GOLD / "synth70_c1.npz")
    atm, wave = g["atmosphere"], g["wave"]
    cwd = rd.make_workdir("benchmark")
    rf = np.zeros((len(ROWS), len(DEPTHS), 4, len(wave)))
    rd.rhf1d(atm, wave, cwd)        # warm-up: see the note on the first call below
    for ip, (r, d) in enumerate(zip(ROWS, DELTA)):
        for ik, k in enumerate(DEPTHS):
            sp = []
            for sgn in (+1.0, -1.0):
                a = atm.copy()
                a[r, k] = a[r, k] + d if sgn > 0 else a[r, k] - d
                o = rd.rhf1d(a, wave, cwd)
                sp.append(np.array([o[s] for s in "IQUV"]))
            rf[ip, ik] = (sp[0] - sp[1]) / (2.0 * d)
    # self-check: every entry again, in reverse order.  (Observed once: the very first rhf1d() call of a process
    # started as `python -m oracle.gen_golden_rf_fd` returned a spectrum 1.6 % off the one every later call --
    # and every other process -- gives for the same input; hence the warm-up call and this check.)
    for ip in reversed(range(len(ROWS))):
        for ik in reversed(range(len(DEPTHS))):
            r, d, k = ROWS[ip], DELTA[ip], DEPTHS[ik]
            sp = []
            for sgn in (+1.0, -1.0):
                a = atm.copy()
                a[r, k] = a[r, k] + d if sgn > 0 else a[r, k] - d
                o = rd.rhf1d(a, wave, cwd)
                sp.append(np.array([o[s] for s in "IQUV"]))
            assert np.array_equal(rf[ip, ik], (sp[0] - sp[1]) / (2.0 * d)), (ip, ik)
    np.savez_compressed(GOLD / "rf_fd.npz", atmosphere=atm, wave=wave, rows=ROWS, delta=DELTA, depths=DEPTHS, rf=rf)
    print(f"[golden] rf_fd: {rf.shape}, max |dI/dT| = {np.abs(rf[0, :, 0]).max():.3e} -> "
          f"{(GOLD / 'rf_fd.npz').stat().st_size/1e3:.0f} kB")


if __name__ == "__main__":
    main()

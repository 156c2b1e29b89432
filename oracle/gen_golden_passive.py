"""TEST INFRASTRUCTURE ONLY (oracle).  Golden vectors for bound-bound lines of PASSIVE atoms in the background.

Runs the compiled, unmodified reference on FAL-C (v_z != 0) around Na I D (589.2 nm) and H-alpha (656.5 nm), where
the PASSIVE Na.atom and H_6.atom of benchmark/atoms.input have lines, and records every passive_bb() call that
found a line (rh/metal.c:174-344) with the lines' parameters and per-depth inputs (level populations, Doppler
width, Damping()).  Output: tests/golden/falc_passive_bb.npz.   Usage: python -m oracle.gen_golden_passive
"""
import numpy as np

from oracle import refdriver as rd
from oracle import portdriver as pd
from oracle.gen_golden import GOLD, recs_by_tag, one, falc_case_atm


def main():
    atm = falc_case_atm()
    wave = np.concatenate([np.linspace(588.9, 589.3, 9), np.linspace(656.1, 656.7, 13)])
    cwd = rd.make_workdir("benchmark")
    o = rd.rhf1d(atm, wave, cwd, probe=rd.PROBE_RLK | rd.PROBE_SNAP)
    R = recs_by_tag(o["records"])
    N = atm.shape[1]
    rows, cs, cf, pcol = [], [], [], []
    for m, d in sorted(R["pbb_line"], key=lambda x: (x[0][0], x[0][1])):       # atom order, then line order
        nc = m[2]
        rows.append(list(d[:7]) + [len(cs)])
        cs += list(d[8:8 + nc]); cf += list(d[8 + nc:8 + 2 * nc])
        pcol.append(d[8 + 2 * nc:].reshape(4, N))
    fl = one(R, "flags")
    out = dict(atmosphere=atm, wave=wave, lam_spect=one(R, "lambda"), muz=one(R, "muz"), flags=fl,
               col_vel=one(R, "vel"), plines=np.array(rows), c_shift=np.array(cs), c_fraction=np.array(cf),
               pcol=np.array(pcol), pbb_meta=np.array([m[:3] for m, _ in R["pbb"]], np.int32),
               pbb=np.array([d.reshape(2, N) for _, d in R["pbb"]]))
    np.savez_compressed(GOLD / "falc_passive_bb.npz", **out)
    ok = 0
    for m, d in zip(out["pbb_meta"], out["pbb"]):
        chi, eta, has = pd.passive_bb(out["plines"], out["c_shift"], out["c_fraction"], fl[6], out["lam_spect"][m[0]],
                                      float(out["muz"][m[1]]), bool(fl[0]), int(m[2]), out["col_vel"], out["pcol"])
        ok += has == 1 and np.array_equal(chi, d[0]) and np.array_equal(eta, d[1])
    print(f"[golden] falc_passive_bb: {len(rows)} lines, {len(out['pbb'])} calls with a line; port exact {ok}/{len(out['pbb'])} "
          f"-> {(GOLD / 'falc_passive_bb.npz').stat().st_size/1e3:.0f} kB")


if __name__ == "__main__":
    main()

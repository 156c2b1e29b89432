"""TEST INFRASTRUCTURE ONLY (oracle).  Windows in which Background() adds bound-bound lines of the PASSIVE model atoms
(passive_bb, metal.c:174-344): Na I D (Na.atom's 3s-3p line, UNSOLD damping), H-alpha (H_6.atom, linear Stark, wide
wings), the Mg I b region (Mg.atom) -- benchmark column 1 with the benchmark's own Kurucz list (no Kurucz line in
these windows), mu = 1 and 0.8.  Output: tests/golden/passive_fused.npz.  Usage: python -m oracle.gen_golden_passive_fused
"""
import numpy as np

from oracle import refdriver as rd
from oracle.gen_golden import GOLD, recs_by_tag, one

WINDOWS = {"NaD": np.linspace(589.0, 589.7, 71), "Halpha": np.linspace(655.2, 657.6, 81), "Mgb": np.linspace(517.7, 518.3, 61)}


def main():
    atm = np.load(GOLD / "synth70_c1.npz")["atmosphere"]
    cwd = rd.make_workdir("benchmark")
    rd.rhf1d(atm, WINDOWS["NaD"], cwd)                 # warm-up
    out = dict(atmosphere=atm)
    for name, wave in WINDOWS.items():
        for tag, mu in (("mu1", 1.0), ("mu08", 0.8)):
            o = rd.rhf1d(atm, wave, cwd, mu=mu, probe=rd.PROBE_SNAP)
            R = recs_by_tag(o["records"])
            out[f"{name}_wave"], out[f"{name}_lam"] = wave, o["lam"]
            out[f"{name}_{tag}_stokes"] = np.array([o[s] for s in "IQUV"])
            out[f"{name}_{tag}_flags"] = one(R, "backgrflags").reshape(-1, 2).astype(np.int32)
        I = out[f"{name}_mu1_stokes"][0]
        print(f"[golden] passive_fused/{name}: line depth {1 - I.min() / I.max():.3f}, hasline at "
              f"{int(out[f'{name}_mu1_flags'][:, 0].sum())} of {len(wave) + 1} wavelengths")
    np.savez_compressed(GOLD / "passive_fused.npz", **out)


if __name__ == "__main__":
    main()

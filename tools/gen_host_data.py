"""Data files of pyrh_b200.host (run in the build container, where /root/reference is mounted).

* elements.npz   -- the 99 element IDs and atomic weights RH keeps in a header (rh/atomweights.h), read at
                    generation time like oracle/scrape_tables.py does for the published opacity tables;
* background_falc11.npz -- the flat background model (level table, bound-free edges, Rayleigh lines, published
                    continuum tables, chemical network) of the reference's standard atoms.input / molecules.input
                    (11 PASSIVE atoms, 12 molecules), as recorded from the reference's parsed state in fixture
                    tests/golden/falc_full.npz.  pyrh_b200.host does not parse *.atom / *.molecule files yet.
Usage: python tools/gen_host_data.py
"""
import re
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")


def main():
    src = (REF / "rh" / "atomweights.h").read_text()
    pairs = re.findall(r'\{"(..)",\s*([0-9.]+)\}', src)
    assert len(pairs) == 99, len(pairs)
    out = ROOT / "pyrh_b200" / "data"
    out.mkdir(exist_ok=True)
    np.savez_compressed(out / "elements.npz", ID=np.array([p[0] for p in pairs]),
                        weight=np.array([float(p[1]) for p in pairs]))
    g = np.load(ROOT / "tests" / "golden" / "falc_full.npz")
    keep = {k: g[k] for k in g.files if k.startswith(("ct_hdr", "ct_lev", "ct_bf", "ct_tab_", "ct_ray", "tab_", "ce_nuclei", "ce_mol"))}
    atoms = [ln.split()[0] for ln in (REF / "benchmark" / "atoms.input").read_text().splitlines()
             if ln.strip() and not ln.strip().startswith("#") and ".atom" in ln]
    ids = [a.split(".")[0].split("_")[0].upper().ljust(2) for a in atoms]
    el = [p[0] for p in pairs]
    keep["atom_files"] = np.array(atoms)
    keep["atom_pt_index"] = np.array([el.index(i) + 1 for i in ids], np.int32)
    np.savez_compressed(out / "background_falc11.npz", **keep)
    print("elements:", len(pairs), "atoms:", list(zip(atoms, keep["atom_pt_index"])))


if __name__ == "__main__":
    sys.exit(main())

"""pyrh_b200/data/opacity_tables.txt: the published background cross-section tables (H- bf/ff, H2- ff, H2+ ff,
Rayleigh H2, OH and CH bf) as plain text for C hosts -- the same numbers as data/background_falc11.npz (tab_*), one
block per table: ``name count`` then the values (%.17g round-trips doubles exactly).  The RH host holds these as
function-static arrays in hydrogen.c / ohchbf.c; a host that cannot reach them hands the library this file's
content through rhb200_continuum_model (integration/pyrh_b200_bridge.c does)."""
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
g = np.load(ROOT / "pyrh_b200" / "data" / "background_falc11.npz")
with open(ROOT / "pyrh_b200" / "data" / "opacity_tables.txt", "w") as f:
    for k in g.files:
        if k.startswith("tab_"):
            v = np.asarray(g[k], np.float64).ravel()
            f.write(f"{k[4:]} {len(v)}\n")
            for i in range(0, len(v), 6):
                f.write(" ".join(f"{x:.17g}" for x in v[i:i + 6]) + "\n")
print("written", (ROOT / "pyrh_b200" / "data" / "opacity_tables.txt").stat().st_size, "bytes")

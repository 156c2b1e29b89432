import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLD = ROOT / "tests" / "golden"
CASES = ["falc_B1kG", "falc_B1kG_wide", "synth70_c0", "synth70_c1", "synth70_c2"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", params=CASES)
def golden(request):
    return request.param, dict(np.load(GOLD / f"{request.param}.npz"))


@pytest.fixture(scope="session")
def golden_falc():
    return dict(np.load(GOLD / "falc_B1kG.npz"))


def port_objects(g, matinv_simd=False):
    """(LineTable, PortTables, PortColumn) for a golden fixture."""
    from oracle import portdriver as pd
    from pyrh_b200.linelist import LineTable
    lt = LineTable.from_npz(g)
    tab = pd.PortTables(lt, matinv_simd=matinv_simd)
    col = pd.PortColumn(muz=float(g["muz"][0]), moving=bool(g["flags"][0]),
                        **{k: g["col_" + k] for k in pd.PortColumn.FIELDS})
    return lt, tab, col

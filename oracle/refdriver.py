"""oracle/refdriver.py -- TEST INFRASTRUCTURE ONLY.

ctypes driver for the *unmodified* reference library built by oracle/build_ref.sh
(oracle/_ref/liboracle_{scalar,simd}.so).  It calls the reference's own C entry
point ``rhf1d()`` (rh/rhf1d/pyrh_compute1dray.h:26-36) exactly the way
``pyrh.compute1d`` does (pyrh.pyx:621-632) and reads back the record log that
oracle/probe.c fills through ``ld --wrap``.

Only tests/, the golden-vector generator, __graft_entry__.smoke() and the
cpu_baseline / ``--impl reference`` legs of bench.py may import this module.
The product package (pyrh_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REFDIR = HERE / "_ref"

PROBE_RLK, PROBE_BG, PROBE_DELO, PROBE_SNAP = 1, 2, 4, 8
PROBE_BEZ, PROBE_FEAU, PROBE_NLTE, PROBE_FORMAL = 16, 32, 64, 128
PROBE_ALL = 255
PROBE_CONT = 256

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class AtomPops(C.Structure):
    # pyrh_compute1dray.h:4-10
    _fields_ = [("ID", C.c_char * 10), ("Nlevel", C.c_int), ("Nz", C.c_int),
                ("n", C.POINTER(c_double_p)), ("nstar", C.POINTER(c_double_p))]


class MySpectrum(C.Structure):
    # pyrh_compute1dray.h:12-19 (note: rh.pxd:120-132 omits J20; layout is the header's)
    _fields_ = [("nlw", C.c_int), ("Nrays", C.c_int), ("stokes", C.c_int),
                ("lam", c_double_p), ("sI", c_double_p), ("sQ", c_double_p),
                ("sU", c_double_p), ("sV", c_double_p),
                ("J", C.POINTER(c_double_p)), ("J20", C.POINTER(c_double_p)),
                ("rfs", C.POINTER(c_double_p)),
                ("Nactive_atoms", C.c_int), ("atom_pops", C.POINTER(AtomPops))]


class ProbeRec(C.Structure):
    _fields_ = [("tag", C.c_char * 24), ("meta", C.c_int * 8), ("n", C.c_long),
                ("data", c_double_p)]


def available(variant: str = "scalar") -> bool:
    return (REFDIR / f"liboracle_{variant}.so").exists() and (REFDIR / "pyrh_path").exists()


_libs: dict[str, C.CDLL] = {}


def load(variant: str = "scalar") -> C.CDLL:
    if variant in _libs:
        return _libs[variant]
    # "bridged": the reference's library with integration/pyrh_b200_bridge.c compiled in (integration/build_bridged.sh):
    # same rhf1d() prototype, per-column work on the GPU through librhb200
    path = HERE / "_build" / "libpyrh_bridged.so" if variant == "bridged" else REFDIR / f"liboracle_{variant}.so"
    if not path.exists():
        raise FileNotFoundError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
    os.environ["PYRH_PATH"] = str(REFDIR / "pyrh_path")
    os.environ.setdefault("RHB200_DATA", str(HERE.parent / "pyrh_b200" / "data"))
    lib = C.CDLL(str(path), mode=os.RTLD_LOCAL)
    lib.rhf1d.restype = MySpectrum
    lib.rhf1d.argtypes = [
        C.c_char_p, C.c_double, C.c_int,
        c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
        c_double_p, c_double_p, c_double_p, c_double_p, C.c_int,
        C.c_int, c_double_p,
        C.c_int, c_double_p, c_double_p,
        C.c_int, c_int_p, c_double_p,
        C.c_int, c_int_p, c_double_p,
        C.c_int, c_int_p, c_double_p,
        C.c_int, C.c_int, C.c_int, C.c_char_p]
    if variant != "bridged":
        lib.probe_enable.argtypes = [C.c_uint]
        lib.probe_reset.argtypes = []
        lib.probe_count.restype = C.c_long
        lib.probe_get.restype = C.POINTER(ProbeRec)
        lib.probe_get.argtypes = [C.c_long]
    else:
        lib.rhf1d_batch.restype = MySpectrum
        lib.rhf1d_batch.argtypes = [C.c_char_p, C.c_double, C.c_int, C.c_int, c_double_p, C.c_int, C.c_int, c_double_p,
                                    C.c_int, c_double_p, c_double_p, C.c_int, c_int_p, c_double_p, C.c_int, c_int_p,
                                    c_double_p, C.c_int, c_int_p, c_double_p, c_double_p, c_double_p, c_double_p, c_int_p]
    _libs[variant] = lib
    return lib


def make_workdir(kind: str = "benchmark", keywords: dict | None = None,
                 atoms_active: tuple = (), kurucz_lines: str | None = None,
                 no_kurucz: bool = False, atoms_extra: tuple = (), root: str | None = None) -> str:
    """Stage a cwd for rhf1d(): the reference's own input set (`benchmark/` or
    `tests/`) with optional keyword overrides (``KEY = value`` lines replaced or
    appended) and optional ACTIVE atoms."""
    src = REFDIR / "inputs" / kind
    d = tempfile.mkdtemp(prefix=f"rhref_{kind}_", dir=root)
    for f in src.iterdir():
        if f.is_file() and f.suffix not in (".py", ".fits", ".spec", ".png"):
            shutil.copy(f, d)
    if keywords:
        p = Path(d) / "keyword.input"
        lines = p.read_text().splitlines()
        for k, v in keywords.items():
            done = False
            for i, ln in enumerate(lines):
                s = ln.strip()
                if s.startswith("#"):
                    continue
                if s.split("=")[0].strip() == k:
                    lines[i] = f"  {k} = {v}"
                    done = True
            if not done:
                lines.append(f"  {k} = {v}")
        p.write_text("\n".join(lines) + "\n")
    if atoms_active:
        p = Path(d) / "atoms.input"
        lines = p.read_text().splitlines()
        for i, ln in enumerate(lines):
            w = ln.split()
            if w and w[0] in atoms_active:
                lines[i] = ln.replace("PASSIVE", "ACTIVE ")
        p.write_text("\n".join(lines) + "\n")
    if atoms_extra:
        # append model atoms that the shipped atoms.input lacks (e.g. CaII.atom) and bump Nmetal
        p = Path(d) / "atoms.input"
        lines = p.read_text().splitlines()
        for i, ln in enumerate(lines):
            w = ln.split()
            if w and not ln.strip().startswith("#") and w[0].isdigit():
                lines[i] = f"   {int(w[0]) + len(atoms_extra)}"
                break
        last = max(i for i, ln in enumerate(lines) if ".atom" in ln)
        for name, mode in atoms_extra:
            lines.insert(last + 1, f"  {name}        {mode}     LTE_POPULATIONS   pops.{name.split('.')[0]}.out")
        p.write_text("\n".join(lines) + "\n")
    if kurucz_lines is not None:
        (Path(d) / "kurucz_lines.dat").write_text(kurucz_lines)
        (Path(d) / "kurucz.input").write_text("kurucz_lines.dat\n")
    if no_kurucz:
        # an empty list of line files: readKuruczLines() (kurucz.c:159) reads nothing, Nrlk stays 0
        (Path(d) / "kurucz.input").write_text("# no Kurucz line files\n")
    return d


def polarizable_cn_tree(dst: str) -> str:
    """A $PYRH_PATH whose CN.molecule lists a POLARIZABLE copy of the CN B-X list the reference ships: none of the
    shipped molecular lists carries the Hund's-case columns readMolecularLines() looks for behind column 71
    (readmolecule.c:859-912), so MolZeeman() is unreachable with stock data.  The copy appends " B S 0.5 B S 0.5"
    (B 2Sigma+ - X 2Sigma+: Hund's case b, Lambda = 0, S = 1/2) to every line, labels the list KURUCZ_CD18 and raises log gf by 7.
    Everything else is symlinked."""
    src = REFDIR / "pyrh_path" / "rh"
    d = Path(dst) / "rh"
    (d / "Molecules" / "CN").mkdir(parents=True)
    for f in src.iterdir():
        if f.name != "Molecules":
            os.symlink(f, d / f.name)
    for f in (src / "Molecules").iterdir():
        if f.name not in ("CN", "CN.molecule"):
            os.symlink(f, d / "Molecules" / f.name)
    lines = (src / "Molecules" / "CN" / "CN_B-X_SCIP_CH1.asc").read_text().splitlines()
    # the shipped list has its subbranch digits where KURUCZ_CD18 expects them (columns 56 / 64); under its own
    # KURUCZ_NEW label sscanf("%1d") at column 57 fails and MolZeeman() would read an uninitialised mrt->subi
    # log gf + 7: the shipped values (around -10) leave the lines invisible against the continuum
    out = [ln.replace("KURUCZ_NEW", "KURUCZ_CD18") if i == 0 else ln if (i < 2 or not ln.strip() or ln[0] == "#")
           else ln[:10] + f"{float(ln[10:17]) + 7.0:7.3f}" + ln[17:].ljust(53) + " B S 0.5 B S 0.5" for i, ln in enumerate(lines)]
    (d / "Molecules" / "CN" / "CN_B-X_polarizable.asc").write_text("\n".join(out) + "\n")
    mol = (src / "Molecules" / "CN.molecule").read_text().replace("CN/CN_B-X_SCIP_CH1.asc", "CN/CN_B-X_polarizable.asc")
    (d / "Molecules" / "CN.molecule").write_text(mol)
    return str(Path(dst))


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def rhf1d(atmosphere: np.ndarray, wave: np.ndarray, cwd: str, mu: float = 1.0,
          atm_scale: int = 0, variant: str = "scalar", probe: int = 0,
          loggf_ids=None, loggf_values=None, get_atomic_rfs: bool = False,
          get_populations: bool = False, fudge_wave=None, fudge_value=None):
    """Call the reference rhf1d().  `atmosphere` is the pyrh layout [9+, Ndep]
    (pyrh.pyx:621-625): scale, T[K], ne[cm^-3], vz[km/s], vmic[km/s], B[G],
    gamma[rad], chi[rad], nH[cm^-3].  Returns dict(lam, I, Q, U, V[, rfs][, records])."""
    lib = load(variant)
    atm = np.ascontiguousarray(atmosphere, dtype=np.float64).copy()
    wave = np.ascontiguousarray(wave, dtype=np.float64).copy()
    ndep = atm.shape[1]
    rows = [np.ascontiguousarray(atm[i]) for i in range(9)]
    nl = 0 if loggf_ids is None else len(loggf_ids)
    lg_ids = np.ascontiguousarray(loggf_ids if nl else [0], dtype=np.int32)
    lg_val = np.ascontiguousarray(loggf_values if nl else [0.0], dtype=np.float64)
    nf = 0 if fudge_wave is None else len(fudge_wave)
    f_lam = np.ascontiguousarray(fudge_wave if nf else [0.0], dtype=np.float64)
    f_val = np.ascontiguousarray(fudge_value if nf else [0.0], dtype=np.float64)     # [3][nf] row-major
    if variant != "bridged":
        lib.probe_reset()
        lib.probe_enable(probe)
    old = os.getcwd()
    os.chdir(cwd)   # Kurucz list entries are opened relative to the process cwd (kurucz.c:160-165)
    try:
        spec = lib.rhf1d(str(cwd).encode(), float(mu), ndep,
                         *[_dp(r) for r in rows], int(atm_scale),
                         len(wave), _dp(wave),
                         nf, _dp(f_lam) if nf else None, _dp(f_val) if nf else None,
                         nl, lg_ids.ctypes.data_as(c_int_p), _dp(lg_val),
                         0, None, None,
                         0, None, None,
                         int(get_atomic_rfs), int(get_populations), 0, None)
    finally:
        os.chdir(old)
    n = spec.nlw
    out = {k: np.ctypeslib.as_array(getattr(spec, f), shape=(n,)).copy()
           for k, f in (("lam", "lam"), ("I", "sI"), ("Q", "sQ"), ("U", "sU"), ("V", "sV"))}
    if get_atomic_rfs and nl:
        out["rfs"] = np.array([[spec.rfs[i][j] for j in range(nl)] for i in range(n)])
    if get_populations and spec.Nactive_atoms > 0:
        pops = {}
        for a in range(spec.Nactive_atoms):
            ap = spec.atom_pops[a]
            nlv, nz = ap.Nlevel, ap.Nz
            pops[ap.ID.decode()] = dict(
                n=np.array([[ap.n[i][k] for k in range(nz)] for i in range(nlv)]),
                nstar=np.array([[ap.nstar[i][k] for k in range(nz)] for i in range(nlv)]))
        out["pops"] = pops
    if probe:
        out["records"] = records(lib)
    if variant != "bridged":
        lib.probe_enable(0)
    return out


def rhf1d_batch(atmospheres: np.ndarray, wave: np.ndarray, cwd: str, mu: float = 1.0, atm_scale: int = 0,
                get_populations: bool = False, nlev: int = 0):
    """rhf1d_batch() of the bridged library: [ncol, 9, Ndep] pyrh rows -> dict(lam, stokes [ncol, 4, nlw][, n, nstar, niter])."""
    lib = load("bridged")
    atm = np.ascontiguousarray(atmospheres, np.float64)[:, :9].copy()
    wave = np.ascontiguousarray(wave, np.float64).copy()
    ncol, _, ndep = atm.shape
    cap = len(wave) + 4096
    st = np.zeros((ncol, 4, cap))
    n = np.zeros((ncol, nlev, ndep)) if get_populations else None
    ns = np.zeros((ncol, nlev, ndep)) if get_populations else None
    nit = np.zeros(ncol, np.int32)
    old = os.getcwd()
    os.chdir(cwd)
    try:
        spec = lib.rhf1d_batch(str(cwd).encode(), float(mu), ndep, ncol, _dp(atm), int(atm_scale), len(wave), _dp(wave),
                               0, None, None, 0, None, None, 0, None, None, 0, None, None, _dp(st),
                               _dp(n) if get_populations else None, _dp(ns) if get_populations else None,
                               nit.ctypes.data_as(c_int_p))
    finally:
        os.chdir(old)
    nlw = spec.nlw
    out = dict(lam=np.ctypeslib.as_array(spec.lam, shape=(nlw,)).copy(), stokes=st.reshape(-1)[:ncol * 4 * nlw].reshape(ncol, 4, nlw).copy())
    if get_populations:
        out.update(n=n, nstar=ns, niter=nit)
    return out


def records(lib):
    recs = []
    for i in range(lib.probe_count()):
        r = lib.probe_get(i).contents
        data = np.ctypeslib.as_array(r.data, shape=(max(r.n, 1),))[: r.n].copy()
        recs.append((r.tag.decode(), tuple(r.meta), data))
    return recs


# ---------------------------------------------------------------- atmospheres

def spinor2multi(atm: np.ndarray) -> np.ndarray:
    """SPINOR 12-column table -> pyrh [9, Ndep] rows; same arithmetic as the
    reference's own helper (tests/test_compute1d.py:6-32, k_B = 1.380649e-23)."""
    k = 1.380649e-23
    new = np.empty((9, atm.shape[-1]), dtype=np.float64)
    new[0] = atm[0]
    new[1] = atm[2]
    new[2] = atm[4] / 10 / k / atm[2] / 1e6
    new[3] = atm[9] / 1e5
    new[4] = atm[8] / 1e5
    new[5] = atm[7]
    new[6] = atm[-2]
    new[7] = atm[-1]
    new[8] = (atm[3] - atm[4]) / 10 / k / atm[2] / 1e6 / 1.26
    return new


def falc(kind: str = "benchmark") -> np.ndarray:
    raw = np.loadtxt(REFDIR / "inputs" / kind / "falc.dat", skiprows=1).T
    return spinor2multi(np.array(raw, dtype=np.float64))


def air_to_vacuum(w):
    """benchmark/synth.py:48-58 (all our wavelengths are > 200 nm)."""
    w = np.asarray(w, dtype=np.float64)
    s2 = (1.0e7 / w) ** 2
    fact = 1.0000834213 + 2.406030e6 / (1.3e10 - s2) + 1.5997e4 / (3.89e9 - s2)
    return w * fact


def hinode_wave(n: int = 301) -> np.ndarray:
    """benchmark/synth.py:76-77."""
    return air_to_vacuum(np.linspace(630.2 - 0.15, 630.2 + 0.15, num=n))

"""Host side of the pyrh boundary: ``compute1d(cwd, mu, atm_scale, atmosphere, wave, ...)`` with the argument list
of ``pyrh.compute1d`` (pyrh.pyx:537-551), reading the working directory's text inputs itself and running every
per-column step on the B200 through ``rhb200_compute1d_batch``.

What is parsed here (once per working directory, cached), with the reference's own expressions:

* ``keyword.input``                      readInput, rh/readinput.c:43-420 (the keywords the LTE path looks at)
* ``abundance.input`` + ``pf_Kurucz.input`` (XDR)   readAbundance, rh/abundance.c:72-222
* the Kurucz line lists of ``KURUCZ_DATA``          readKuruczLines, rh/kurucz.c:121-431; getUnsoldcross :985-1024;
  getABOcross / getBarklemcross with the Barklem tables and cubeconvol, rh/barklem.c:61-212, rh/cubeconvol.c;
  Zeeman patterns by the library's RLKdeterminate / RLKZeeman
* the merged wavelength grid                         SortLambda, rh/sortlambda.c:180-210 (user grid + lambda_ref)

* the ``*.atom`` files of ``atoms.input``   readAtom, rh/readatom.c:100-425 (levels, lines with their damping
  constants, bound-free continua) and the ``*.molecule`` files of ``molecules.input`` (readMolecule,
  rh/readmolecule.c:60-215: constituents, dissociation energy, equilibrium-constant fit)

Only the published opacity tables RH keeps inside its C sources (H-, H2-, H2+, OH, CH) come from a data file
(``data/background_falc11.npz``).  A working directory that asks for ACTIVE atoms,
MAGNETO_OPTICAL or anything else this path does not implement is refused loudly
(``NotImplementedError``), never approximated.
"""
from __future__ import annotations

import math
import os
import re
import struct
from collections import OrderedDict
from dataclasses import dataclass
from pathlib import Path

import numpy as np

from . import linelist as ll
from . import zeeman

DATA = Path(__file__).resolve().parent / "data"

# rh/constant.h:27-60, rh/rh.h:36 (same literals)
CLIGHT, HPLANCK, KBOLTZMANN, AMU = 2.99792458E+08, 6.6260755E-34, 1.380658E-23, 1.6605402E-27
M_ELECTRON, Q_ELECTRON, EPSILON_0 = 9.1093897E-31, 1.60217733E-19, 8.854187817E-12
RBOHR, E_RYDBERG, ABARH = 5.29177349E-11, 2.1798741E-18, 7.42E-41
NM_TO_M, CM_TO_M, PI, LG10, MILLI = 1.0E-09, 1.0E-02, 3.14159265358979, 2.30258509299404568402, 1.0E-03
AIR_TO_VACUUM_LIMIT = 199.9352           # spectrum.h:20
RLK_RECORD_LENGTH, RLK_LABEL_LENGTH = 171, 10   # kurucz.c:88, atom.h:25


def POW10(x):                            # rh.h:50
    return math.exp(LG10 * x)


_FLOAT = re.compile(r"\s*([+-]?(?:\d+\.?\d*(?:[eEdD][+-]?\d+)?|\.\d+(?:[eEdD][+-]?\d+)?))")
_INT = re.compile(r"\s*([+-]?\d+)")


def _scan_float(s: str):
    """sscanf(s, "%lf"): leading float or None."""
    m = _FLOAT.match(s)
    return (float(m.group(1).replace("d", "e").replace("D", "e")), m.end()) if m else (None, 0)


def _scan_int_w(s: str, width: int):
    """sscanf(s, "%<width>d"): skips white space, then reads at most `width` characters."""
    t = s.lstrip()
    m = _INT.match(t[:width])
    return (int(m.group(1)), len(s) - len(t) + m.end()) if m else (None, 0)


# ------------------------------------------------------------------------------------------- keyword.input
KEYWORD_DEFAULTS = {"LS_LANDE": "TRUE", "KURUCZ_DATA": "none", "METALLICITY": "0.0", "LAMBDA_REF": "500.0",
                    "VMACRO_TRESH": "0.1", "STOKES_MODE": "NO_STOKES", "MAGNETO_OPTICAL": "TRUE",
                    "RLK_SCATTER": "FALSE", "N_MAX_SCATTER": "5", "S_INTERPOLATION": "S_BEZIER3",
                    "S_INTERPOLATION_STOKES": "DELO_BEZIER3", "ATOMS_FILE": "atoms.input",
                    "MOLECULES_FILE": "molecules.input", "HYDROGEN_LTE": "FALSE", "SOLVE_NE": "NONE",
                    "OPACITY_FUDGE": "none", "ALLOW_PASSIVE_BB": "TRUE"}


def read_keywords(cwd) -> dict:
    """``KEY = VALUE`` pairs of ``<cwd>/keyword.input`` ('#' starts a comment), defaults of readinput.c:43-215 for
    the keywords this path reads."""
    kw = dict(KEYWORD_DEFAULTS)
    for ln in (Path(cwd) / "keyword.input").read_text().splitlines():
        ln = ln.split("#", 1)[0].strip()
        if "=" in ln:
            k, v = ln.split("=", 1)
            kw[k.strip().upper()] = v.strip()
    if "VMICRO_CHAR" not in kw:
        raise ValueError("keyword.input: VMICRO_CHAR is required (readinput.c:169)")
    return kw


def _true(v: str) -> bool:
    return v.strip().upper() == "TRUE"


# ------------------------------------------------------------------------------------------- elements
@dataclass
class Elements:
    ID: list                 # 99 two-character IDs
    weight: np.ndarray
    abund: np.ndarray        # linear, relative to H (0 where not set)
    abundance_set: np.ndarray
    nstage: np.ndarray
    ionpot: list             # per element [nstage] J
    pf: list                 # per element [nstage][Npf] ln U
    Tpf: np.ndarray
    totalAbund: float
    wght_per_H: float
    avgMolWght: float


def refuse_unported_keywords(kw):
    """Keywords that change rhf1d()'s result and that this path does not implement: refused, never ignored."""
    if _true(kw.get("HYDROSTATIC", "FALSE")):
        raise NotImplementedError("HYDROSTATIC = TRUE (Hydrostatic() inside Iterate(), iterate.c:120-128) is not ported")
    if _true(kw.get("BACKGROUND_POLARIZATION", "FALSE")):
        raise NotImplementedError("BACKGROUND_POLARIZATION = TRUE (scattering polarisation through J20, formal.c:143-321) is not ported")
    if kw.get("ATMOS_ITOP", "none").lower() != "none":
        raise NotImplementedError("ATMOS_ITOP: an irradiated top boundary (multiatmos.c:63-64, 186-220) is not ported")


def pyrh_path(explicit=None) -> Path:
    p = explicit or os.environ.get("PYRH_PATH")
    if not p:
        raise RuntimeError("PYRH_PATH is not set (readinput.c:227: the reference finds rh/Atoms/* through it)")
    return Path(p)


def read_elements(path=None, kw=None, atomic_number=None, atomic_abundance=None) -> Elements:
    """readAbundance (rh/abundance.c:72-222): abundance.input, metallicity, the XDR partition-function file."""
    kw = kw or KEYWORD_DEFAULTS
    atoms_dir = pyrh_path(path) / "rh" / "Atoms"
    tab = np.load(DATA / "elements.npz")
    ID, weight = [str(x) for x in tab["ID"]], np.array(tab["weight"], np.float64)
    ne = len(ID)
    abund, aset = np.zeros(ne), np.zeros(ne, bool)
    DEX = False
    reading = 1
    for ln in Path(kw.get("ABUND_FILE") or atoms_dir / "abundance.input").read_text().splitlines():
        if not ln.strip() or ln[0] == "#":                      # getLine(), rh/getline.c
            continue
        f = ln.split()
        if len(f) < 2:
            raise ValueError(f"abundance.input: cannot read '{ln}'")
        eid, a = f[0].upper(), float(f[1])
        if len(eid) == 1:
            eid += " "
        if atomic_number is not None:                            # abundance.c:118-126: position in the file, 1-based
            for n, v in zip(atomic_number, atomic_abundance):
                if int(n) == reading:
                    a = float(v)
                    break
        reading += 1
        for n in range(ne):
            if eid in ID[n]:
                abund[n] = a
                if eid == "H " and a == 12.0:
                    DEX = True
                aset[n] = True
                break
    metallicity = POW10(float(kw.get("METALLICITY", "0.0")))
    raw = Path(kw.get("KURUCZ_PF_DATA") or atoms_dir / "pf_Kurucz.input").read_bytes()
    off = 0

    def xint():
        nonlocal off
        v = struct.unpack_from(">i", raw, off)[0]
        off += 4
        return v

    def xdoubles(n):
        nonlocal off
        v = np.frombuffer(raw, ">f8", n, off).astype(np.float64)
        off += 8 * n
        return v

    npf = xint()
    Tpf = xdoubles(npf)
    total = avg = 0.0
    nstage, ionpot, pf = np.zeros(ne, np.int32), [None] * ne, [None] * ne
    for n in range(ne):
        if aset[n]:
            if DEX:
                abund[n] = POW10(abund[n] - 12.0)
            if metallicity != 1.0 and ID[n] != "H ":
                abund[n] *= metallicity
            total += abund[n]
            avg += abund[n] * weight[n]
        xint()                                                   # pti
        nstage[n] = xint()
        p = xdoubles(nstage[n] * npf).reshape(nstage[n], npf)
        ip = xdoubles(nstage[n])
        if aset[n]:
            ip = np.array([x * ((HPLANCK * CLIGHT) / CM_TO_M) for x in ip])     # `*=`: abundance.c:212
            p = np.array([[math.log(x) for x in row] for row in p])
        ionpot[n], pf[n] = ip, p
    return Elements(ID, weight, abund, aset, nstage, ionpot, pf, Tpf, total, avg, avg / total)


# ------------------------------------------------------------------------------------------- Kurucz lines
def air_to_vacuum(lambda_air: float) -> float:               # rh/vacuumtoair.c (air_to_vacuum)
    if lambda_air >= AIR_TO_VACUUM_LIMIT:
        sqwave = (1.0E+07 / lambda_air) * (1.0E+07 / lambda_air)
        increase = 1.0000834213E+00 + 2.406030E+06 / (1.30E+10 - sqwave) + 1.5997E+04 / (3.89E+09 - sqwave)
        return lambda_air * increase
    return lambda_air


def _gammln(xx: float) -> float:                              # rh/gammafunc.c
    cof = (76.18009172947146, -86.50532032941677, 24.01409824083091, -1.231739572450155, 0.1208650973866179e-2,
           -0.5395239384953e-5)
    x = y = xx
    tmp = x + 5.5
    tmp -= (x + 0.5) * math.log(tmp)
    ser = 1.000000000190015
    for c in cof:
        y += 1
        ser += c / y
    return -tmp + math.log(2.5066282746310005 * ser / x)


# Barklem / Anstee / O'Mara tables (barklem.c:31-49)
BARKLEM = {"SP": ("Barklem_spdata.dat", 21, 18, 1.0, 1.3), "PD": ("Barklem_pddata.dat", 18, 18, 1.3, 2.3),
           "DF": ("Barklem_dfdata.dat", 18, 18, 2.3, 3.3)}
BARKLEM_DELTA_NEFF = 0.1


def read_barklem_table(kind, path=None):
    """readBarklemTable (barklem.c:61-132): 3 header lines, N1 x N2 cross-sections, 2 more lines, N1 x N2 alphas."""
    fname, n1, n2, neff1_0, neff2_0 = BARKLEM[kind]
    lines = (pyrh_path(path) / "rh" / "Atoms" / fname).read_text().splitlines()
    nums = []
    pos = 3
    while len(nums) < n1 * n2:
        nums += [float(x) for x in lines[pos].split()]
        pos += 1
    cross = nums[:n1 * n2]
    pos += 2                                                     # fgets x2 after fscanf left the newline unread: the
    nums = []                                                    # rest of the last data line + one comment line
    tail = lines[pos - 1:]
    for ln in tail:
        try:
            nums += [float(x) for x in ln.split()]
        except ValueError:
            nums = []
        if len(nums) >= n1 * n2:
            break
    alpha = nums[:n1 * n2]
    neff1 = [neff1_0 + n * BARKLEM_DELTA_NEFF for n in range(n1)]
    neff2 = [neff2_0 + n * BARKLEM_DELTA_NEFF for n in range(n2)]
    return dict(N1=n1, N2=n2, neff1=neff1, neff2=neff2, cross=cross, alpha=alpha)


def _cc_kernel(s):                                               # cubeconvol.c:143-151
    return [-s * (1.0 - s * (2.0 - s)) / 2.0, 1.0 + s * s * (3.0 * s - 5.0) / 2.0,
            s * (1.0 + s * (4.0 - 3.0 * s)) / 2.0, s * s * (s - 1.0) / 2.0]


def cubeconvol(Nx, Ny, f, x, y):
    """Cubic-convolution interpolation in a row-major table f[Ny][Nx] at fractional indices (x, y): cubeconvol.c:30-139."""
    if x <= 0.0:
        i, ux = 0, _cc_kernel(0.0)
    elif x >= float(Nx - 1):
        i, ux = Nx - 2, _cc_kernel(1.0)
    else:
        fr, ip = math.modf(x)
        ux, i = _cc_kernel(fr), int(ip)
    if y <= 0.0:
        j, uy = 0, _cc_kernel(0.0)
    elif y >= float(Ny - 1):
        j, uy = Ny - 2, _cc_kernel(1.0)
    else:
        fr, jp = math.modf(y)
        uy, j = _cc_kernel(fr), int(jp)
    base = Nx * (j - 1) + i - 1
    c = [[0.0] * 4 for _ in range(4)]

    def fill(m, row_base):
        if i == 0:
            for n in range(1, 4):
                c[m][n] = f[row_base + n]
            c[m][0] = 3.0 * (c[m][1] - c[m][2]) + c[m][3]
        elif i == Nx - 2:
            for n in range(0, 3):
                c[m][n] = f[row_base + n]
            c[m][3] = 3.0 * (c[m][2] - c[m][1]) + c[m][0]
        else:
            for n in range(4):
                c[m][n] = f[row_base + n]

    if j == 0:
        fc = base + Nx
        for m in range(1, 4):
            fill(m, fc)
            fc += Nx
        for n in range(4):
            c[0][n] = 3.0 * (c[1][n] - c[2][n]) + c[3][n]
    elif j == Ny - 2:
        fc = base
        for m in range(0, 3):
            fill(m, fc)
            fc += Nx
        for n in range(4):
            c[3][n] = 3.0 * (c[2][n] - c[1][n]) + c[0][n]
    else:
        fc = base
        if i != 0 and i != Nx - 2:                               # interior: convolve without copying
            g = 0.0
            for m in range(4):
                for n in range(4):
                    g += f[fc + n] * ux[n] * uy[m]
                fc += Nx
            return g
        for m in range(4):
            fill(m, fc)
            fc += Nx
    g = 0.0
    for m in range(4):
        for n in range(4):
            g += c[m][n] * ux[n] * uy[m]
    return g


def barklem_cross(bs, stage, Ei, Ej, li, lj, ionpot, weight, H_weight):
    """getBarklemcross (barklem.c:139-196): (cross, alpha) or None where the reference falls back to Unsold."""
    if stage > 0:
        return None
    deltaEi, deltaEj = ionpot - Ei, ionpot - Ej
    if deltaEi <= 0.0 or deltaEj <= 0.0:
        return None
    Z = float(stage + 1)
    E_Ryd = E_RYDBERG / (1.0 + M_ELECTRON / (weight * AMU))
    neff1, neff2 = Z * math.sqrt(E_Ryd / deltaEi), Z * math.sqrt(E_Ryd / deltaEj)
    return _barklem_interpolate(bs, neff1, neff2, li, lj, weight, H_weight)


def barklem_active_cross(at, ln, weight, H_weight, path=None):
    """getBarklemactivecross (barklem.c:216-312) for a model-atom line: (cvdWaals[0], cvdWaals[1]) or None where
    readatom.c:313-319 falls back to UNSOLD."""
    from . import zeeman
    i, j, stage, E = ln["i"], ln["j"], at["stage"], at["E"]
    if stage[i] > 0:
        return None
    if at["abo_level"][i] != -1 and at["abo_level"][j] != -1:
        Ll, Lu = at["abo_level"][i], at["abo_level"][j]
    else:
        di, dj = zeeman.determinate(at["label"][i], at["g"][i]), zeeman.determinate(at["label"][j], at["g"][j])
        if not (di[0] and dj[0]):
            return None
        Ll, Lu = di[3], dj[3]
    kind = {frozenset((0, 1)): "SP", frozenset((1, 2)): "PD", frozenset((2, 3)): "DF"}.get(frozenset((Ll, Lu)))
    if kind is None or Ll == Lu:
        return None
    bs = read_barklem_table(kind, path)
    Z = float(stage[j] + 1)
    ic = j + 1
    while stage[ic] < stage[j] + 1:
        ic += 1
    deltaEi, deltaEj = E[ic] - E[i], E[ic] - E[j]
    E_Ryd = E_RYDBERG / (1.0 + M_ELECTRON / (weight * AMU))
    neff1, neff2 = Z * math.sqrt(E_Ryd / deltaEi), Z * math.sqrt(E_Ryd / deltaEj)
    return _barklem_interpolate(bs, neff1, neff2, Ll, Lu, weight, H_weight)


def _barklem_interpolate(bs, neff1, neff2, li, lj, weight, H_weight):
    if li > lj:
        neff1, neff2 = neff2, neff1
    t1, t2 = bs["neff1"], bs["neff2"]
    if neff1 < t1[0] or neff1 > t1[-1]:
        return None

    def locate(arr, v):                                          # hunt.c:92-117, ascending
        lo, hi = 0, len(arr)
        while hi - lo > 1:
            mid = (hi + lo) >> 1
            if v >= arr[mid]:
                lo = mid
            else:
                hi = mid
        return lo
    k = locate(t1, neff1)
    findex1 = float(k) + (neff1 - t1[k]) / BARKLEM_DELTA_NEFF
    if neff2 < t2[0] or neff2 > t2[-1]:
        return None
    k = locate(t2, neff2)
    findex2 = float(k) + (neff2 - t2[k]) / BARKLEM_DELTA_NEFF
    cross = cubeconvol(bs["N2"], bs["N1"], bs["cross"], findex2, findex1)
    alpha = cubeconvol(bs["N2"], bs["N1"], bs["alpha"], findex2, findex1)
    reducedmass = AMU / (1.0 / H_weight + 1.0 / weight)
    meanvelocity = math.sqrt(8.0 * KBOLTZMANN / (PI * reducedmass))
    crossmean = (RBOHR * RBOHR) * math.pow(meanvelocity / 1.0E4, -alpha)
    cross *= 2.0 * math.pow(4.0 / PI, alpha / 2.0) * math.exp(_gammln((4.0 - alpha) / 2.0)) * meanvelocity * crossmean
    return cross, alpha


def read_kurucz_records(cwd, kurucz_data: str):
    """The fixed-length records of every list named in KURUCZ_DATA (kurucz.c:157-184): list entries are opened
    relative to the process cwd in the reference (kurucz.c:160-165); here relative to `cwd`."""
    recs = []
    for ln in (Path(cwd) / kurucz_data).read_text().splitlines():
        if not ln.strip() or ln[0] == "#":
            continue
        text = (Path(cwd) / ln.split()[0]).read_text()
        for phys in text.split("\n"):
            phys_nl = phys + "\n"
            while phys_nl:                                       # fgets(inputLine, RLK_RECORD_LENGTH+1, ...)
                rec, phys_nl = phys_nl[:RLK_RECORD_LENGTH], phys_nl[RLK_RECORD_LENGTH:]
                if rec.strip("\n") == "" and not phys:
                    continue
                if rec[0] != "#":
                    recs.append(rec.rstrip("\n"))
    return recs


def read_kurucz_lines(cwd, kw: dict, el: Elements, loggf_ids=None, loggf_values=None, lam_ids=None,
                      lam_values=None, path=None) -> ll.LineTable:
    """readKuruczLines (kurucz.c:121-431) + the lazily built Zeeman patterns (kurucz.c:832-921) -> LineTable,
    sorted by lambda0 (background.c:292-294)."""
    C = 2.0 * PI * (Q_ELECTRON / EPSILON_0) * (Q_ELECTRON / M_ELECTRON) / CLIGHT
    LS_Lande = _true(kw["LS_LANDE"])
    rows, patterns, used, barklem, si_rows, file_idx = [], [], {}, {}, [], []
    if kw["KURUCZ_DATA"].lower() == "none":
        raise NotImplementedError("KURUCZ_DATA = none: the LTE path needs a Kurucz line list")
    for line_index, rec in enumerate(read_kurucz_records(cwd, kw["KURUCZ_DATA"])):
        f = rec.split()
        lambda_air, gf, elem_code, Ei = float(f[0]), float(f[1]), f[2], float(f[3])
        pt_index, stage = (int(x) for x in elem_code.split("."))
        Ej = _scan_float(rec[53:])[0]
        Ei = abs(Ei) * (HPLANCK * CLIGHT) / CM_TO_M
        Ej = abs(Ej) * (HPLANCK * CLIGHT) / CM_TO_M
        swap = Ej < Ei
        if swap:
            rEi, rEj = Ej, Ei
            labeli, labelj = rec[69:69 + RLK_LABEL_LENGTH], rec[41:41 + RLK_LABEL_LENGTH]
        else:
            rEi, rEj = Ei, Ej
            labeli, labelj = rec[41:41 + RLK_LABEL_LENGTH], rec[69:69 + RLK_LABEL_LENGTH]
        Ji, Jj = _scan_float(rec[35:])[0], _scan_float(rec[63:])[0]
        if swap:
            Ji, Jj = Jj, Ji
        gi, gj = 2 * Ji + 1, 2 * Jj + 1
        if lam_ids is not None:
            for i, v in zip(lam_ids, lam_values):
                if int(i) == line_index:
                    lambda_air += float(v)
        lambda0 = air_to_vacuum(lambda_air)                      # USE_TABULATED_WAVELENGTH, kurucz.c:236-243
        lambda0 *= NM_TO_M
        rEj = rEi + (HPLANCK * CLIGHT) / lambda0
        Aji = C / (lambda0 * lambda0) * POW10(gf) / gj
        if loggf_ids is not None:
            for i, v in zip(loggf_ids, loggf_values):
                if int(i) == line_index:
                    Aji = C / (lambda0 * lambda0) * POW10(float(v)) / gj
        Bji = (lambda0 * lambda0 * lambda0) / (2.0 * HPLANCK * CLIGHT) * Aji
        Bij = (gj / gi) * Bji
        e = pt_index - 1
        if not el.abundance_set[e]:
            continue            # the reference keeps the line but rlk_opacity never adds it (kurucz.c:612-614)
        # ABO columns (kurucz.c:271-294).  The shipped lists end before column 160: the reference then scans
        # whatever its buffer holds there; no information is the only defined reading
        cross = alpha = 0.0
        li = lj = -1
        got_ABO = False
        if len(rec) > 160:
            a, n1 = _scan_float(rec[160:])
            s = _scan_float(rec[160 + n1:])[0] if a is not None else None
            if a is not None and s is not None:
                if a in (0.0, 1.0, 2.0, 3.0) or s in (0.0, 1.0, 2.0, 3.0):
                    li, lj = (int(s), int(a)) if swap else (int(a), int(s))
                else:
                    got_ABO, alpha, cross = True, a, s
        determined, Si, Li, Sj, Lj = zeeman.rlk_determinate(labeli, labelj)
        polarizable = determined                                 # atmos.Stokes is TRUE on the pyrh path (:259)
        G = rec[79:79 + 18]
        Grad, n1 = _scan_float(G)
        GStark, n2 = _scan_float(G[n1:])
        GvdW = _scan_float(G[n1 + n2:])[0]
        GStark = POW10(GStark) * (CM_TO_M * CM_TO_M * CM_TO_M) if GStark != 0.0 else 0.0
        GvdW = POW10(GvdW) * (CM_TO_M * CM_TO_M * CM_TO_M) if GvdW != 0.0 else 0.0
        vdw = None
        if got_ABO:                                              # getABOcross, barklem.c:199-212
            reducedmass = AMU / (1.0 / el.weight[0] + 1.0 / el.weight[e])
            meanvelocity = math.sqrt(8.0 * KBOLTZMANN / (PI * reducedmass))
            crossmean = (RBOHR * RBOHR) * math.pow(meanvelocity / 1.0E4, -alpha)
            cross *= 2.0 * math.pow(4.0 / PI, alpha / 2.0) * math.exp(_gammln((4.0 - alpha) / 2.0)) * meanvelocity * crossmean
            vdw = ll.VDW_BARKLEM
        else:                                                    # kurucz.c:322-332: tables by orbital quantum numbers
            kind = {frozenset((0, 1)): "SP", frozenset((1, 2)): "PD", frozenset((2, 3)): "DF"}.get(frozenset((li, lj)))
            if kind and li != lj:
                if kind not in barklem:
                    barklem[kind] = read_barklem_table(kind, path)
                hit = barklem_cross(barklem[kind], stage, rEi, rEj, li, lj, el.ionpot[e][stage], el.weight[e], el.weight[0])
                if hit:
                    cross, alpha = hit
                    vdw = ll.VDW_BARKLEM
        if vdw is None:                                          # getUnsoldcross, kurucz.c:985-1024
            if stage > el.nstage[e] - 1:
                vdw = ll.VDW_KURUCZ
            else:
                Z = stage + 1
                d1, d2 = E_RYDBERG / (el.ionpot[e][stage] - rEj), E_RYDBERG / (el.ionpot[e][stage] - rEi)
                deltaR = d1 * d1 - d2 * d2
                if deltaR <= 0.0:
                    vdw = ll.VDW_KURUCZ
                else:
                    FOURPIEPS0 = 4.0 * PI * EPSILON_0
                    w = el.weight[e]
                    vrel35_H = math.pow(8.0 * KBOLTZMANN / (PI * AMU * w) * (1.0 + w / el.weight[0]), 0.3)
                    vrel35_He = math.pow(8.0 * KBOLTZMANN / (PI * AMU * w) * (1.0 + w / el.weight[1]), 0.3)
                    ZR = Z * RBOHR
                    C625 = math.pow(2.5 * ((Q_ELECTRON * Q_ELECTRON) / FOURPIEPS0) * (ABARH / FOURPIEPS0) *
                                    2 * PI * (ZR * ZR) / HPLANCK * deltaR, 0.4)
                    cross = 8.08 * (vrel35_H + el.abund[1] * vrel35_He) * C625
                    vdw = ll.VDW_UNSOLD
        Grad = POW10(Grad) if Grad != 0.0 else Aji
        iso_frac = POW10(_scan_float(rec[108:])[0])
        hfs_frac = POW10(_scan_float(rec[117:])[0])
        g1, n1 = _scan_int_w(rec[143:], 5)
        g2 = _scan_int_w(rec[143 + n1:], 5)[0]
        gL_i, gL_j = g1 * MILLI, g2 * MILLI
        if swap:
            gL_i, gL_j = gL_j, gL_i
        if not LS_Lande and gL_i != -99 * MILLI and gL_j != -99 * MILLI:    # kurucz.c:381-385
            polarizable = True
        si_rows.append((lambda0, gi, gj, C / (lambda0 * lambda0) * POW10(gf) / gj))
        r = np.zeros(ll.RL_NFIELD)
        r[ll.RL_LAMBDA0] = lambda0 / NM_TO_M
        r[ll.RL_GI], r[ll.RL_GJ], r[ll.RL_EI], r[ll.RL_EJ] = gi, gj, rEi, rEj
        r[ll.RL_BJI], r[ll.RL_AJI], r[ll.RL_BIJ] = Bji, Aji, Bij
        r[ll.RL_GRAD], r[ll.RL_GSTARK], r[ll.RL_GVDW] = Grad, GStark, GvdW
        r[ll.RL_HFS_FRAC], r[ll.RL_ISO_FRAC] = hfs_frac, iso_frac
        r[ll.RL_CROSS], r[ll.RL_ALPHA] = cross, alpha
        r[ll.RL_POLARIZABLE], r[ll.RL_VDWAALS], r[ll.RL_STAGE] = float(polarizable), vdw, stage
        if e not in used:
            used[e] = len(used)
        r[ll.RL_ELEM] = used[e]
        pat = zeeman.rlk_zeeman(gi, gj, Si, Li, Sj, Lj, gL_i, gL_j, LS_Lande) if polarizable else None
        rows.append(r)
        patterns.append(pat)
        file_idx.append(line_index)
    order = sorted(range(len(rows)), key=lambda i: rows[i][ll.RL_LAMBDA0])       # stable; qsort in the reference
    zq, zs, zst, out = [], [], [], []
    for i in order:
        r, pat = rows[i], patterns[i]
        r[ll.RL_ZOFF] = len(zq)
        if pat is not None:
            r[ll.RL_NCOMP] = len(pat[0])
            zq += list(pat[0]); zs += list(pat[1]); zst += list(pat[2])
        out.append(r)
    elems, pfrows = [], []
    for e, _ in sorted(used.items(), key=lambda kv: kv[1]):
        nst = int(el.nstage[e])
        if nst > ll.RE_MAXSTAGE:
            raise ValueError(f"element {el.ID[e]}: {nst} ionisation stages exceed RE_MAXSTAGE")
        row = np.zeros(ll.RE_NFIELD)
        row[ll.RE_WEIGHT], row[ll.RE_ABUND], row[ll.RE_NSTAGE], row[ll.RE_PFROW] = el.weight[e], el.abund[e], nst, len(pfrows)
        row[ll.RE_IONPOT0:ll.RE_IONPOT0 + nst] = el.ionpot[e]
        pfrows += list(el.pf[e])
        elems.append(row)
    lt = ll.LineTable(lines=np.array(out), zq=np.array(zq, np.int32), zshift=np.array(zs, np.float64),
                      zstrength=np.array(zst, np.float64), elems=np.array(elems), pf=np.array(pfrows), Tpf=el.Tpf,
                      vmicro_char=float(kw["VMICRO_CHAR"]) * 1.0E+03)
    lt.validate()
    lt.elem_rows = dict(used)                                    # periodic-table index - 1 -> row of lt.elems
    lt.file_index = np.array([file_idx[i] for i in order], np.int64)     # table row -> line number in the Kurucz files
    lt.si = [si_rows[i] for i in order]                          # per table row: lambda0 [m], gi, gj, Aji of the file's log gf
    return lt


VDW_NONE, VDW_UNSOLD_A, VDW_RIDDER_A = -1, 0, 1
# passive-line rows of the device table -- include/rhb200.h RHB200_PL_*
(PL_ATOM, PL_LEVEL_I, PL_LEVEL_J, PL_LAMBDA0, PL_QWING, PL_BIJ, PL_BJI, PL_AJI, PL_VOIGT, PL_NCOMP, PL_COMPOFF, PL_GRAD,
 PL_VDW_TYPE, PL_VDW_A, PL_VDW_B, PL_VDW_C, PL_VDW_D, PL_HE_ABUND, PL_STARK_TYPE, PL_STARK_A, PL_STARK_C, PL_STARK_CM,
 PL_LINSTARK_C, PL_IS_H, PL_WEIGHT) = range(25)
PL_NFIELD = 28


def read_atom(atom_file):
    """Levels and bound-bound lines of one model atom: the part of readAtom (rh/readatom.c:100-330) that passive_bb
    (metal.c:174-344), Damping (broad.c:273-314) and rlk_opacity's duplicate check (kurucz.c:617-633) use."""
    data = [ln for ln in Path(atom_file).read_text().splitlines() if ln.strip() and ln[0] != "#"]
    ID = data[0].split()[0][:2].upper().ljust(2)
    nlevel, nline = (int(x) for x in data[1].split()[:2])
    E, g, label, stage, abo_level = [], [], [], [], []
    for ln in data[2:2 + nlevel]:
        head, tail = ln.split("'")[0].split(), ln.split("'")[2].split()
        E.append(float(head[0]) * ((HPLANCK * CLIGHT) / CM_TO_M))           # `*=`, readatom.c:175
        g.append(float(head[1]))
        label.append(ln.split("'")[1])
        stage.append(int(tail[0]))
        abo_level.append(int(tail[2]) if len(tail) > 2 and re.fullmatch(r"[+-]?\d+", tail[2]) else -1)   # readatom.c:166-168
    C = 2 * PI * (Q_ELECTRON / EPSILON_0) * (Q_ELECTRON / M_ELECTRON) / CLIGHT      # readatom.c:213
    lines, pos = [], 2 + nlevel
    for _ in range(nline):
        f = data[pos].split()
        pos += 1
        j, i = int(f[0]), int(f[1])
        i, j = min(i, j), max(i, j)
        lambda0 = (HPLANCK * CLIGHT) / (E[j] - E[i])
        Aji = C / (lambda0 * lambda0) * (g[i] / g[j]) * float(f[2])
        Bji = (lambda0 * lambda0 * lambda0) / (2.0 * HPLANCK * CLIGHT) * Aji
        Bij = (g[j] / g[i]) * Bji
        shape, vdw = f[3], f[8]
        cvdW = [float(x) for x in f[9:13]]
        c_shift, c_fraction = [0.0], [1.0]
        if "COMPOSIT" in shape:
            nc = int(data[pos].split()[0])
            comp = [data[pos + 1 + n].split() for n in range(nc)]
            pos += 1 + nc
            c_shift, c_fraction = [float(x[0]) for x in comp], [float(x[1]) for x in comp]
        if "BARKLEM" in vdw and stage[i] > 0:                    # getBarklemactivecross: ABO tables are for neutrals only
            vdw = "UNSOLD"                                       # (barklem.c:232-233) -> readatom.c:313-319 falls back
        if "UNSOLD" in vdw:
            cvdW[1] = cvdW[3] = 0.0
        lines.append(dict(i=i, j=j, lambda0=lambda0 / NM_TO_M, Aji=Aji, Bji=Bji, Bij=Bij, voigt="GAUSS" not in shape,
                          qwing=float(f[7]), vdw=vdw, cvdW=cvdW, Grad=float(f[13]), cStark=float(f[14]),
                          c_shift=c_shift, c_fraction=c_fraction))
    return dict(ID=ID, E=E, g=g, label=label, stage=stage, abo_level=abo_level, lines=lines)


def read_atom_lines(atom_file):
    """(ID, [(stage of the lower level, lambda0 [nm], qwing), ...]) for rlk_opacity's duplicate check."""
    a = read_atom(atom_file)
    return a["ID"], [(a["stage"][ln["i"]], ln["lambda0"], ln["qwing"]) for ln in a["lines"]]


def passive_line_table(cwd, kw, el: Elements, level_first, path=None, status="PASSIVE"):
    """Device table of the bound-bound lines of the PASSIVE model atoms (passive_bb, metal.c:174-344) with the
    depth-independent factors of Damping() (broad.c:60-314) evaluated here: rows [nline, PL_NFIELD] in the reference's
    order (atoms, then lines), component shifts and fractions.  ``level_first[a]`` = row of atom a's first level in
    the population table.  ``status`` = "ACTIVE" gives the same table for the lines of the ACTIVE atoms (Damping() of
    getProfiles, profile.c:112)."""
    atoms_dir = pyrh_path(path) / "rh" / "Atoms"
    rows, cs, cf = [], [], []
    FOURPIEPS0 = 4.0 * PI * EPSILON_0
    H_weight, He_weight, He_abund = el.weight[0], el.weight[1], el.abund[1]
    for a, (fname, st) in enumerate(_atoms_listed(cwd, kw)):
        if (st == "ACTIVE") != (status == "ACTIVE"):
            continue
        at = read_atom(atoms_dir / fname)
        e = el.ID.index(at["ID"])
        weight, E, stage = el.weight[e], at["E"], at["stage"]
        for ln in at["lines"]:
            i, j, cv = ln["i"], ln["j"], ln["cvdW"]
            r = np.zeros(PL_NFIELD)
            r[PL_ATOM], r[PL_LEVEL_I], r[PL_LEVEL_J] = a, level_first[a] + i, level_first[a] + j
            r[PL_LAMBDA0], r[PL_QWING] = ln["lambda0"], ln["qwing"]
            r[PL_BIJ], r[PL_BJI], r[PL_AJI], r[PL_VOIGT] = ln["Bij"], ln["Bji"], ln["Aji"], float(ln["voigt"])
            r[PL_NCOMP], r[PL_COMPOFF], r[PL_GRAD] = len(ln["c_shift"]), len(cs), ln["Grad"]
            r[PL_WEIGHT], r[PL_IS_H], r[PL_HE_ABUND] = weight, float(at["ID"] == "H "), He_abund
            cs += ln["c_shift"]; cf += ln["c_fraction"]
            r[PL_VDW_TYPE] = VDW_NONE
            vdw_kind = ln["vdw"]
            if "BARKLEM" in vdw_kind:                                        # readatom.c:311-320
                hit = barklem_active_cross(at, ln, weight, H_weight, path)
                if hit:
                    cv = [hit[0], hit[1], 1.0, 0.0]                          # barklem.c:295-310
                else:
                    vdw_kind, cv = "UNSOLD", [cv[0], 0.0, cv[2], 0.0]
            if cv[0] > 0.0 or cv[2] > 0.0:                                   # VanderWaals, broad.c:60-140
                if "UNSOLD" in vdw_kind or "BARKLEM" in vdw_kind:
                    vrel35_He = math.pow(8.0 * KBOLTZMANN / (PI * AMU * weight) * (1.0 + weight / He_weight), 0.3)
                    Z = stage[j] + 1
                    ic = j + 1
                    while stage[ic] < stage[j] + 1:
                        ic += 1
                    d1, d2 = E_RYDBERG / (E[ic] - E[j]), E_RYDBERG / (E[ic] - E[i])
                    deltaR = d1 * d1 - d2 * d2
                    ZR = Z * RBOHR
                    C625 = math.pow(2.5 * ((Q_ELECTRON * Q_ELECTRON) / FOURPIEPS0) * (ABARH / FOURPIEPS0) *
                                    2 * PI * (ZR * ZR) / HPLANCK * deltaR, 0.4)
                    if "BARKLEM" in vdw_kind:                                # broad.c:125-136
                        r[PL_VDW_TYPE] = 2
                        r[PL_VDW_A], r[PL_VDW_B] = cv[0], (1.0 - cv[1]) / 2.0
                        r[PL_VDW_C] = 8.08 * cv[2] * He_abund * vrel35_He * C625
                    else:
                        vrel35_H = math.pow(8.0 * KBOLTZMANN / (PI * AMU * weight) * (1.0 + weight / H_weight), 0.3)
                        r[PL_VDW_TYPE] = VDW_UNSOLD_A
                        r[PL_VDW_A] = 8.08 * (cv[0] * vrel35_H + cv[2] * He_abund * vrel35_He) * C625
                elif "PARAMTR" in vdw_kind:
                    CUBE_CM = CM_TO_M * CM_TO_M * CM_TO_M
                    gH = 1.0E-8 * CUBE_CM * math.pow(1.0 + H_weight / weight, cv[1])
                    gHe = 1.0E-9 * CUBE_CM * math.pow(1.0 + He_weight / weight, cv[3])
                    r[PL_VDW_TYPE] = VDW_RIDDER_A
                    r[PL_VDW_A], r[PL_VDW_B], r[PL_VDW_C], r[PL_VDW_D] = gH * cv[0], cv[1], gHe * cv[2], cv[3]
                else:
                    raise ValueError(f"{fname} line {j}->{i}: invalid van der Waals keyword {vdw_kind} (readatom.c:322)")
            cS = ln["cStark"]
            if cS < 0.0:                                                     # Stark, broad.c:147-215
                r[PL_STARK_TYPE], r[PL_STARK_A] = 1, abs(cS)
            elif cS != 0.0:
                m_electron = M_ELECTRON / AMU
                Cc = 8.0 * KBOLTZMANN / (PI * AMU * weight)
                Cm = math.pow(1.0 + weight / m_electron, 0.16666667) + math.pow(1.0 + weight / 28.0, 0.16666667)
                Z = stage[i] + 1
                ic = i + 1
                while stage[ic] < stage[i] + 1 and ic < len(stage):
                    ic += 1
                E_Ryd = E_RYDBERG / (1.0 + M_ELECTRON / (weight * AMU))
                neff_l = Z * math.sqrt(E_Ryd / (E[ic] - E[i]))
                neff_u = Z * math.sqrt(E_Ryd / (E[ic] - E[j]))
                Z2 = Z * Z
                tu, tl = neff_u * (5.0 * (neff_u * neff_u) + 1.0), neff_l * (5.0 * (neff_l * neff_l) + 1.0)
                C4 = ((Q_ELECTRON * Q_ELECTRON) / (4.0 * PI * EPSILON_0)) * RBOHR * \
                    (2.0 * PI * (RBOHR * RBOHR) / HPLANCK) / (18.0 * Z2 * Z2) * (tu * tu - tl * tl)
                r[PL_STARK_TYPE], r[PL_STARK_A], r[PL_STARK_C], r[PL_STARK_CM] = 2, 11.37 * math.pow(cS * C4, 0.66666667), Cc, Cm
            if at["ID"] == "H ":                                             # StarkLinear, broad.c:222-264
                def nq(lab):
                    m = re.match(r"H I (\d+)", lab)
                    return int(m.group(1))
                n_lower, n_upper = nq(at["label"][i]), nq(at["label"][j])
                a1 = 0.642 if n_upper - n_lower == 1 else 1.0
                r[PL_LINSTARK_C] = a1 * 0.6 * (n_upper * n_upper - n_lower * n_lower) * (CM_TO_M * CM_TO_M)
            rows.append(r)
    return np.array(rows).reshape(-1, PL_NFIELD), np.array(cs), np.array(cf)


def passive_line_windows(cwd, kw, path=None, atoms=True):
    """Wavelength windows [lo, hi] (nm) in which the reference's Background() adds lines this package's fused LTE path
    does not sum yet: passive_bb (lines of the PASSIVE model atoms, metal.c:245-246) and MolecularOpacity (line lists
    of PASSIVE molecules, opacity.c:774-787).  Session refuses grids that touch them instead of silently missing
    opacity."""
    root = pyrh_path(path) / "rh"
    vchar = float(kw["VMICRO_CHAR"]) * 1.0E+03 / CLIGHT
    out = []
    if atoms and _true(kw.get("ALLOW_PASSIVE_BB", "TRUE")):
        for fname, _ in _atoms_listed(cwd, kw):
            ID, lines = read_atom_lines(root / "Atoms" / fname)
            out += [(lam0 - lam0 * qw * vchar, lam0 + lam0 * qw * vchar, f"{ID.strip()} line at {lam0:.3f} nm") for _, lam0, qw in lines]
    for ln in (Path(cwd) / kw["MOLECULES_FILE"]).read_text().splitlines():
        f = ln.split("#", 1)[0].split()
        if len(f) >= 2 and f[0].endswith(".molecule"):
            body = [x.strip() for x in (root / "Molecules" / f[0]).read_text().splitlines() if x.strip() and x[0] != "#"]
            for item in body:
                lst = root / "Molecules" / item
                if "/" in item and lst.is_file():
                    rows = lst.read_text().splitlines()
                    qwing = float(rows[1].split()[1])
                    lam = [float(r[:10]) for r in rows[2:] if r.strip()]
                    lo, hi = min(lam), max(lam)
                    out.append((lo - lo * qwing * vchar, hi + hi * qwing * vchar, f"{f[0]} line list {item}"))
    return out


def model_line_rows(cwd, kw, el: Elements, lt_elem_order, path=None):
    """rows for Context.set_model_lines: the lines of every listed model atom whose element also has Kurucz lines."""
    rows = []
    atoms_dir = pyrh_path(path) / "rh" / "Atoms"
    for fname, _ in _atoms_listed(cwd, kw):
        ID, lines = read_atom_lines(atoms_dir / fname)
        e = el.ID.index(ID)
        if e in lt_elem_order:
            rows += [(lt_elem_order[e], st, lam0, qw) for st, lam0, qw in lines]
    return np.array(rows, np.float64).reshape(-1, 4)


# ------------------------------------------------------------------------------------------- background model
EV = 1.60217733E-19
FIT_TYPES = ("KURUCZ_70", "KURUCZ_85", "SAUVAL_TATUM_84", "IRWIN_81", "TSUJI_73")        # enum fit_type, atom.h:32


def read_atom_continua(atom_file, atom):
    """Bound-free continua of one model atom (readatom.c:372-425): [(j, i, alpha0, hydrogenic, lambda0, lambda[], alpha[])]."""
    data = [ln for ln in Path(atom_file).read_text().splitlines() if ln.strip() and ln[0] != "#"]
    nlevel, nline, ncont = (int(x) for x in data[1].split()[:3])
    pos = 2 + nlevel
    for _ in range(nline):                                       # skip the lines (and their component records)
        shape = data[pos].split()[3]
        pos += 1
        if "COMPOSIT" in shape:
            pos += 1 + int(data[pos].split()[0])
    E, out = atom["E"], []
    for _ in range(ncont):
        f = data[pos].split()
        pos += 1
        j, i = int(f[0]), int(f[1])
        i, j = min(i, j), max(i, j)
        alpha0, nlam, dep, lambdamin = float(f[2]), int(f[3]), f[4], float(f[5])
        lambda0 = ((HPLANCK * CLIGHT) / (E[j] - E[i])) / NM_TO_M
        lam, alp = [0.0] * nlam, [0.0] * nlam
        if "EXPLICIT" in dep:
            for la in range(nlam - 1, -1, -1):                   # the file lists the table from red to blue
                w = data[pos].split()
                pos += 1
                lam[la], alp[la] = float(w[0]), float(w[1])
            hydrogenic = 0.0
        elif "HYDROGENIC" in dep:
            dlamb = (lambda0 - lambdamin) / (nlam - 1)           # getLambdaCont, readatom.c
            lam[0] = lambdamin
            for la in range(1, nlam):
                lam[la] = lam[la - 1] + dlamb
            hydrogenic = 1.0
        else:
            raise ValueError(f"{atom_file}: wavelength dependence {dep}")
        out.append((j, i, alpha0, hydrogenic, lambda0, lam, alp))
    return out


def read_molecule(mol_file, el: Elements):
    """One *.molecule file up to the equilibrium-constant fit (readmolecule.c:60-215)."""
    data = [ln.split("#")[0] for ln in Path(mol_file).read_text().splitlines() if ln.strip() and ln.lstrip()[0] != "#"]
    ID, charge = data[0].split()[0], int(data[1].split()[0])
    pt_index, pt_count = [], []
    for tok in re.split(r"[ ,]+", data[2].strip()):
        m = re.match(r"(\d*)(\S+)", tok)
        cnt, sym = (int(m.group(1)) if m.group(1) else 1), m.group(2).upper()
        if len(sym) == 1:
            sym += " "
        e = next(k for k in range(len(el.ID)) if el.ID[k] in sym)          # strstr(elementID, elements[m].ID)
        if not el.abundance_set[e]:
            raise ValueError(f"{mol_file}: no abundance for {sym}")
        pt_index.append(e); pt_count.append(cnt)
    Ediss = float(data[3].split()[0]) * EV
    fit = next(k for k, nme in enumerate(FIT_TYPES) if nme in data[4].split()[0])
    Tmin, Tmax = (float(x) for x in data[5].split()[:2])
    pfl = data[6].split()
    npf = int(pfl[0])
    pf_coef = [0.0] * npf
    for n in range(npf - 1, -1, -1):                             # readmolecule.c:209-213
        pf_coef[n] = float(pfl[1 + (npf - 1 - n)])
    eq = data[7].split()
    neqc = int(eq[0])
    coef = [0.0] * neqc
    for n in range(neqc - 1, -1, -1):                            # stored last-to-first (readmolecule.c:196-198)
        coef[n] = float(eq[1 + (neqc - 1 - n)])
    weight = 0.0
    for e, cnt in zip(pt_index, pt_count):                       # readmolecule.c:226-229
        weight += cnt * el.weight[e]
    return dict(ID=ID, charge=charge, pt_index=pt_index, pt_count=pt_count, Ediss=Ediss, fit=fit, Tmin=Tmin, Tmax=Tmax,
                eqc=coef, pf_coef=pf_coef, weight=weight, line_lists=[x.split()[0] for x in data[8:]])


(ML_LAMBDA0, ML_EI, ML_GI, ML_BIJ, ML_AJI, ML_BJI, ML_ISO_FRAC, ML_QWING, ML_POLARIZABLE, ML_MOL, ML_ZOFF, ML_NCOMP) = range(12)
ML_NFIELD = 16
MS_NFIELD = 16         # molecule rows of Context.set_molecular_lines: chem index, weight, fit, Tmin, Tmax, Npf, pf_coef[8]


def mol_lande(hund, Lambda, S, sub, J):
    """Lande factor of a molecular level: Hund's case a (MolLande_a, molzeeman.c:107-115; the reference never sets
    Omega, so the factor is 0) or case b (MolLande_b, :120-136) with N = J - S + (sub - 1) (:159, :171)."""
    if hund == "A":
        Omega = 0.0
        return (Lambda + 2.0 * S) * Omega / (J * (J + 1.0))
    N = J - S + (sub - 1)
    if Lambda == 0:
        return 1.0 / (J * (J + 1.0)) * (J * (J + 1.0) - N * (N + 1.0) + S * (S + 1.0))
    return 1.0 / (J * (J + 1.0)) * (Lambda * Lambda / (2 * N * (N + 1.0)) * (J * (J + 1.0) + N * (N + 1.0) - S * (S + 1.0)) +
                                    J * (J + 1.0) - N * (N + 1.0) + S * (S + 1.0))


def mol_zeeman_strength(Ju, Mu, Jl, Ml):
    """MolZeemanStr (molzeeman.c:36-101): strength of the component (Ju, Mu) -> (Jl, Ml), q = Ml - Mu, dJ = Jl - Ju."""
    q, dJ = int(Ml - Mu), int(Jl - Ju)
    if dJ == 1:
        if q == 1:
            return (Ju + 1.0 + Mu) * (Ju + 2.0 + Mu) / (2 * (Ju + 1.0) * (2 * Ju + 1.0) * (2 * Ju + 3.0))
        if q == 0:
            return (Ju + 1.0 + Mu) * (Ju + 1.0 - Mu) / ((Ju + 1.0) * (2 * Ju + 1.0) * (2 * Ju + 3.0))
        return (Ju + 1.0 - Mu) * (Ju + 2.0 - Mu) / (2 * (Ju + 1.0) * (2 * Ju + 1.0) * (2 * Ju + 3.0))
    if dJ == 0:
        if q == 1:
            return (Ju - Mu) * (Ju + 1.0 + Mu) / (2 * Ju * (Ju + 1.0) * (2 * Ju + 1.0))
        if q == 0:
            return Mu * Mu / (Ju * (Ju + 1.0) * (2 * Ju + 1.0))
        return (Ju + Mu) * (Ju + 1.0 - Mu) / (2 * Ju * (Ju + 1.0) * (2 * Ju + 1.0))
    if dJ == -1:
        if q == 1:
            return (Ju - Mu) * (Ju - 1.0 - Mu) / (2 * Ju * (2 * Ju - 1.0) * (2 * Ju + 1.0))
        if q == 0:
            return (Ju + Mu) * (Ju - Mu) / (Ju * (2 * Ju - 1.0) * (2 * Ju + 1.0))
        return (Ju + Mu) * (Ju - 1.0 + Mu) / (2 * Ju * (2 * Ju - 1.0) * (2 * Ju + 1.0))
    raise ValueError(f"MolZeemanStr: invalid dJ {dJ}")


def mol_zeeman(gi, gj, hund_i, Lambda_i, S_i, sub_i, hund_j, Lambda_j, S_j, sub_j):
    """MolZeeman (molzeeman.c:196-319), anomalous splitting: (q, shift, strength) of every component with
    abs(Ml - Mu) <= 1 in the reference's (Ml outer, Mu inner) order, strengths normalised per q."""
    Jl, Ju = np.float64((gi - 1.0) / 2.0), np.float64((gj - 1.0) / 2.0)       # numpy scalars: J = 0 divides like C does
    with np.errstate(all="ignore"):
        gLl, gLu = mol_lande(hund_i, Lambda_i, S_i, sub_i, Jl), mol_lande(hund_j, Lambda_j, S_j, sub_j, Ju)
    Jl, Ju, gLl, gLu = float(Jl), float(Ju), float(gLl), float(gLu)
    q, shift, strength, norm = [], [], [], [0.0, 0.0, 0.0]
    Ml = -Jl
    while Ml <= Jl:
        Mu = -Ju
        while Mu <= Ju:
            if abs(Ml - Mu) <= 1.0:
                q.append(int(Ml - Mu))
                shift.append(gLl * Ml - gLu * Mu)
                strength.append(mol_zeeman_strength(Ju, Mu, Jl, Ml))
                norm[q[-1] + 1] += strength[-1]
            Mu += 1
        Ml += 1
    return q, shift, [s / norm[k + 1] for s, k in zip(strength, q)]


def read_molecular_lines(list_file, mol_sel, zq, zshift, zstrength):
    """One line list of a molecule (readMolecularLines, readmolecule.c:437-770, KURUCZ_NEW / KURUCZ_CD18 formats) ->
    rows [n, ML_NFIELD].  Lines that carry Hund's-case data behind column 71 (:859-912) are polarizable -- pyrh always
    sets atmos.Stokes -- and get their MolZeeman components appended to zq / zshift / zstrength."""
    data = [ln for ln in Path(list_file).read_text().splitlines() if ln.strip() and ln[0] != "#"]
    h = data[0].split()
    nrt, fmt = int(h[0]), h[2]
    if not any(f in fmt for f in ("KURUCZ_NEW", "KURUCZ_CD18")):
        raise NotImplementedError(f"{list_file}: molecular line format {fmt} is not ported")
    sub_col = 56 if "KURUCZ_CD18" in fmt else 57                           # subbranch digit (:790-791, :816-817)
    qwing = float(data[1].split()[1])
    C = 2 * PI * (Q_ELECTRON / EPSILON_0) * (Q_ELECTRON / M_ELECTRON) / CLIGHT
    rows = []
    for ln in data[2:2 + nrt]:
        log_gf, gi, Ei, gj, Ej = float(ln[10:17]), float(ln[17:22]), float(ln[22:32]), float(ln[32:37]), float(ln[37:48])
        Ei = (HPLANCK * CLIGHT) / CM_TO_M * abs(Ei)
        Ej = (HPLANCK * CLIGHT) / CM_TO_M * abs(Ej)
        gi, gj = 2 * gi + 1, 2 * gj + 1
        lambda0 = (HPLANCK * CLIGHT) / (Ej - Ei)
        Aji = C / (lambda0 * lambda0) * POW10(log_gf) / gj
        Bji = (lambda0 * lambda0 * lambda0) / (2.0 * HPLANCK * CLIGHT) * Aji
        Bij = (gj / gi) * Bji
        r = np.zeros(ML_NFIELD)
        r[ML_LAMBDA0], r[ML_EI], r[ML_GI], r[ML_BIJ], r[ML_AJI], r[ML_BJI] = lambda0 / NM_TO_M, Ei, gi, Bij, Aji, Bji
        r[ML_ISO_FRAC], r[ML_QWING], r[ML_MOL] = 1.0, qwing, mol_sel
        if len(ln) + 1 > 71:                                                # strlen(inputLine) counts the newline
            f = ln[71:].split()
            if len(f) < 6 or f[0][0] not in "AB" or f[3][0] not in "AB" or f[1][0] not in "SPDF" or f[4][0] not in "SPDF":
                raise ValueError(f"{list_file}: bad Hund's-case data {ln[71:]!r} (the reference exits, readmolecule.c:867-906)")
            sub = []
            for c0 in (sub_col, sub_col + 8):                              # sscanf(inputLine + c0, "%1d", ...)
                t = ln[c0:].lstrip()
                if not t or not t[0].isdigit():
                    raise ValueError(f"{list_file}: no subbranch digit at column {c0} of {ln!r}: the reference then feeds an "
                                     "uninitialised mrt->subi / subj to MolZeeman (readmolecule.c:790-791, 816-817)")
                sub.append(int(t[0]))
            q, sh, st = mol_zeeman(gi, gj, f[0][0], "SPDF".index(f[1][0]), float(f[2]), sub[0],
                                   f[3][0], "SPDF".index(f[4][0]), float(f[5]), sub[1])
            r[ML_POLARIZABLE], r[ML_ZOFF], r[ML_NCOMP] = 1.0, len(zq), len(q)
            zq += q; zshift += sh; zstrength += st
        rows.append(r)
    return rows


def molecular_line_table(cwd, kw, el: Elements, path=None):
    """(mlines [n, ML_NFIELD], molecules [nsel, MS_NFIELD], (zq, zshift, zstrength)) of the PASSIVE molecules that come
    with line lists, in the order of molecules.input; each molecule's lines ascending in lambda0 (qsort(mrt_ascend),
    readmolecule.c:247); the last item holds the MolZeeman components of the polarizable lines."""
    root = pyrh_path(path) / "rh" / "Molecules"
    rows, sel = [], []
    zq, zshift, zstrength = [], [], []
    k = -1
    for ln in (Path(cwd) / kw["MOLECULES_FILE"]).read_text().splitlines():
        f = ln.split("#", 1)[0].split()
        if len(f) >= 2 and f[0].endswith(".molecule"):
            k += 1
            mo = read_molecule(root / f[0], el)
            if not mo["line_lists"]:
                continue
            mine = []
            for lst in mo["line_lists"]:
                mine += read_molecular_lines(root / lst, len(sel), zq, zshift, zstrength)
            mine.sort(key=lambda r: r[ML_LAMBDA0])
            rows += mine
            if len(mo["pf_coef"]) > 8:
                raise ValueError(f"{f[0]}: more than 8 partition-function coefficients")
            m = np.zeros(MS_NFIELD)
            m[0:6] = k, mo["weight"], mo["fit"], mo["Tmin"], mo["Tmax"], len(mo["pf_coef"])
            m[6:6 + len(mo["pf_coef"])] = mo["pf_coef"]
            sel.append(m)
    return (np.array(rows).reshape(-1, ML_NFIELD), np.array(sel).reshape(-1, MS_NFIELD),
            (np.array(zq, np.int32), np.array(zshift, np.float64), np.array(zstrength, np.float64)))


def read_background_model(cwd, kw, el: Elements, path=None, allow_active=False):
    """The flat background model the device continuum / chemistry kernels take (rhb200_continuum_model,
    rhb200_set_chemistry) from the *.atom and *.molecule files of the working directory's lists: level table,
    bound-free edges with their cross-section tables, the Rayleigh lines of H and He, the chemical network.  The
    published opacity tables RH keeps in its source (H-, H2-, H2+, OH, CH) come from data/background_falc11.npz."""
    root = pyrh_path(path) / "rh"
    listed = _atoms_listed(cwd, kw)
    if any(s != "PASSIVE" for _, s in listed) and not allow_active:
        raise NotImplementedError("ACTIVE atoms: this is the LTE session (host.compute1d routes such directories to NlteSession)")
    tabs = {k: v for k, v in np.load(DATA / "background_falc11.npz").items() if k.startswith("tab_")}
    lev, bf, tl, ta, ray, pt, atoms = [], [], [], [], [], [], []
    l0 = 0
    for m, (fname, st) in enumerate(listed):
        at = read_atom(root / "Atoms" / fname)
        if m == 0 and at["ID"] != "H ":
            raise ValueError("First atomic model is not hydrogen (readatom.c:845-849)")
        atoms.append(at)
        pt.append(el.ID.index(at["ID"]) + 1)
        act = float(st == "ACTIVE")                              # Metal_bf / Hydrogen_bf skip ACTIVE atoms (metal.c:105, hydrogen.c:179)
        for i in range(len(at["E"])):
            lev.append([m, at["E"][i], at["stage"][i], at["g"][i], act])
        for j, i, alpha0, hyd, lambda0, lam, alp in read_atom_continua(root / "Atoms" / fname, at):
            bf.append([m, l0 + i, l0 + j, lambda0, lam[0], hyd, alpha0, len(lam), len(tl), act])
            tl += lam; ta += alp
        if m < 2:                                                # Rayleigh(): lines from the ground level of H and He
            for ln in at["lines"]:
                if ln["i"] == 0:
                    ray.append([m, ln["lambda0"], ln["qwing"], ln["Aji"], at["g"][ln["j"]], at["g"][0], l0, at["stage"][0]])
        l0 += len(at["E"])
    mols = []
    for ln in (Path(cwd) / kw["MOLECULES_FILE"]).read_text().splitlines():
        f = ln.split("#", 1)[0].split()
        if len(f) >= 2 and f[0].endswith(".molecule"):
            if f[1].upper() != "PASSIVE":
                raise NotImplementedError("ACTIVE molecules are not implemented")
            mols.append(read_molecule(root / "Molecules" / f[0], el))
    nuc_elems = sorted({e for mo in mols for e in mo["pt_index"]})
    atom_of = {p - 1: a for a, p in enumerate(pt)}
    nuclei = [[e, atom_of.get(e, -1)] for e in nuc_elems]
    if any(a < 0 for _, a in nuclei):
        raise NotImplementedError("a nucleus bound in molecules has no model atom (chemequil.c:222-228 then uses getfjk)")
    ce_mol = np.zeros((len(mols), 32))
    for r, mo in zip(ce_mol, mols):
        r[0:8] = mo["fit"], mo["charge"], sum(mo["pt_count"]), len(mo["pt_index"]), len(mo["eqc"]), mo["Tmin"], mo["Tmax"], mo["Ediss"]
        r[8:8 + len(mo["eqc"])] = mo["eqc"]
        for j, (e, cnt) in enumerate(zip(mo["pt_index"], mo["pt_count"])):
            r[16 + j], r[20 + j] = nuc_elems.index(e), cnt
        r[24], r[25], r[26] = mo["ID"] == "H2", mo["ID"] == "OH", mo["ID"] == "CH"
    ids = [mo["ID"] for mo in mols]
    any_active = any(st == "ACTIVE" for _, st in listed)
    hdr = np.array([len(listed), len(lev), len(bf), len(tl), len(ray), float(listed[0][1] == "ACTIVE"), float(len(pt) > 1 and pt[1] == 2),
                    float("OH" in ids), float("CH" in ids), float("H2" in ids), float(any_active), float(kw["VMICRO_CHAR"]) * 1.0E+03, 0.0,
                    len(atoms[0]["E"]), 1.0, 1.0])
    out = dict(ct_hdr=hdr, ct_lev=np.array(lev), ct_bf=np.array(bf).reshape(-1, 10), ct_tab_lambda=np.array(tl),
               ct_tab_alpha=np.array(ta), ct_ray=np.array(ray).reshape(-1, 8), ce_nuclei=np.array(nuclei, np.float64),
               ce_mol=ce_mol, atom_files=np.array([a for a, _ in listed]), atom_pt_index=np.array(pt, np.int32))
    out.update(tabs)
    return out


# ------------------------------------------------------------------------------------------- session
def _atoms_listed(cwd, kw):
    out = []
    for ln in (Path(cwd) / kw["ATOMS_FILE"]).read_text().splitlines():
        f = ln.split("#", 1)[0].split()
        if len(f) >= 2 and f[0].endswith(".atom"):
            out.append((f[0], f[1].upper()))
    return out


def sort_lambda(wave, lambda_ref):
    """spectrum.lambda of an all-PASSIVE run (sortlambda.c:180-210): user grid + lambda_ref, ascending, unique."""
    return np.unique(np.concatenate([np.asarray(wave, np.float64), [float(lambda_ref)]]))


class Session:
    """Everything of one working directory + wavelength grid that does not depend on the column, resident on one
    GPU: SURVEY 8(b)'s ``rhb200_open(cwd, ...)``."""

    def __init__(self, cwd, wave, device=0, path=None, loggf_ids=None, loggf_values=None, lam_ids=None,
                 lam_values=None, fudge_wave=None, fudge_value=None, atomic_number=None, atomic_abundance=None):
        from . import api, continuum
        self.cwd = Path(cwd)
        kw = self.kw = read_keywords(cwd)
        # OPACITY_FUDGE names a file the pyrh build never opens (the reader is commented out, background.c:181-208): the
        # factors come from compute1d's fudge_wave / fudge_value alone.  DO_FUDGE = TRUE without them leaves the
        # reference interpolating in an empty table (undefined) -- refused
        if _true(kw.get("DO_FUDGE", "FALSE")) and fudge_wave is None:
            raise NotImplementedError("DO_FUDGE = TRUE without fudge_wave / fudge_value is undefined in the reference")
        refuse_unported_keywords(kw)
        if _true(kw["MAGNETO_OPTICAL"]):
            raise NotImplementedError("MAGNETO_OPTICAL = TRUE is refused (the reference overflows chip_c there, readj.c:328)")
        listed = _atoms_listed(cwd, kw)
        self.el = read_elements(path, kw, atomic_number, atomic_abundance)
        bg = self.background = read_background_model(cwd, kw, self.el, path)
        self.lt = read_kurucz_lines(cwd, kw, self.el, loggf_ids, loggf_values, lam_ids, lam_values, path)
        self.lambda_ref = float(kw["LAMBDA_REF"])
        self.lam = sort_lambda(wave, self.lambda_ref)
        mlines, msel, mzee = molecular_line_table(cwd, kw, self.el, path)      # MolZeeman patterns of polarizable lists included
        self.ctx = api.Context(device)
        self.ctx.set_lines(self.lt, magneto_optical=False, rlkscatter=_true(kw["RLK_SCATTER"]))
        if len(mlines):                                                  # MolecularOpacity, opacity.c:711-839
            self.ctx.set_molecular_lines(mlines, msel, *mzee)
        if _true(kw.get("ALLOW_PASSIVE_BB", "TRUE")):                       # passive_bb, metal.c:174-344
            lev = bg["ct_lev"]
            first = [int(np.flatnonzero(lev[:, 0] == a)[0]) for a in range(len(listed))]
            self.ctx.set_passive_lines(*passive_line_table(cwd, kw, self.el, first, path))
        self.model_lines = model_line_rows(cwd, kw, self.el, self.lt.elem_rows, path)
        self.ctx.set_model_lines(self.model_lines)
        self.ctx.set_scatter(int(kw["N_MAX_SCATTER"]), float(kw.get("ITER_LIMIT", "1.0E-2")))   # pyrh_compute1dray.c:332-337
        self.stokes_mode = kw["STOKES_MODE"].upper()
        # without ACTIVE atoms Formal() solves for Q, U, V only when StokesMode == FULL_STOKES (formal.c:94-95) and
        # adjustStokesMode() leaves the mode alone (zeeman.c:315-317): FIELD_FREE and POLARIZATION_FREE are NO_STOKES here
        self.ctx.set_stokes_mode("FULL_STOKES" if self.stokes_mode == "FULL_STOKES" else "NO_STOKES")
        self.ctx.set_wavelengths(self.lam)
        self.ctx.set_solvers(kw["S_INTERPOLATION"], kw["S_INTERPOLATION_STOKES"])
        abundance = np.array([self.el.abund[int(p) - 1] for p in bg["atom_pt_index"]])
        self.model = continuum.ContinuumModel(bg, fudge_wave, fudge_value)
        self.ctx.set_continuum(self.model, abundance)
        self.ctx.set_chemistry(bg["ce_nuclei"][:, 1].astype(np.int32), bg["ce_mol"])
        self.vmacro_tresh = float(kw["VMACRO_TRESH"])
        self.ctx.set_gravity(self.el.totalAbund)
        # get_atomic_rfs: parameter p <-> the line loggf_ids[p] names (RLK_Line.loggf_rf_ind = the LAST p that names
        # it, kurucz.c:250-259); input.n_atomic_pars = Nloggf + Nlam, and no line carries a parameter beyond Nloggf
        ids = [] if loggf_ids is None else [int(i) for i in loggf_ids]
        row_of = {int(f): r for r, f in enumerate(self.lt.file_index)}
        self.loggf_rows = [row_of.get(i, -1) if i not in ids[p + 1:] else -1 for p, i in enumerate(ids)]
        self._n_lam_pars = 0 if lam_ids is None else len(lam_ids)
        self.n_atomic_pars = len(ids) + self._n_lam_pars
        self._loggf_now = {} if loggf_ids is None else {int(i): float(v) for i, v in zip(loggf_ids, loggf_values)}

    def set_loggf(self, loggf_ids=None, loggf_values=None):
        """Apply pyrh.compute1d's log gf overrides to the RESIDENT line table (kurucz.c:247-257: the last entry naming a
        line wins; lines not named get the file's value back): Aji, Bji, Bij of the affected rows are recomputed with
        readKuruczLines' expressions and patched on the device (rhb200_update_line_strengths) -- no re-parse, no new
        context."""
        import ctypes as C_
        from . import _lib
        Cc = 2.0 * PI * (Q_ELECTRON / EPSILON_0) * (Q_ELECTRON / M_ELECTRON) / CLIGHT
        want = {}
        if loggf_ids is not None:
            for i, v in zip(loggf_ids, loggf_values):
                want[int(i)] = float(v)
        row_of = {int(f): r for r, f in enumerate(self.lt.file_index)}
        current = self.__dict__.setdefault("_loggf_now", {})
        rows, A, Bji_, Bij_ = [], [], [], []
        for fid in set(want) | set(current):
            r = row_of.get(fid)
            if r is None or want.get(fid) == current.get(fid):
                continue
            lambda0, gi, gj, Aji_file = self.lt.si[r]
            Aji = Cc / (lambda0 * lambda0) * POW10(want[fid]) / gj if fid in want else Aji_file
            Bji = (lambda0 * lambda0 * lambda0) / (2.0 * HPLANCK * CLIGHT) * Aji
            rows.append(r); A.append(Aji); Bji_.append(Bji); Bij_.append((gj / gi) * Bji)
            self.lt.lines[r, ll.RL_AJI], self.lt.lines[r, ll.RL_BJI], self.lt.lines[r, ll.RL_BIJ] = Aji, Bji, (gj / gi) * Bji
        if rows:
            ri = np.ascontiguousarray(rows, np.int32)
            a, b, c = (np.ascontiguousarray(x, np.float64) for x in (A, Bji_, Bij_))
            _lib.check(self.ctx.lib.rhb200_update_line_strengths(self.ctx.h, len(rows), ri.ctypes.data_as(_lib.ip),
                                                                 a.ctypes.data_as(_lib.dp), b.ctypes.data_as(_lib.dp),
                                                                 c.ctypes.data_as(_lib.dp)))
        self._loggf_now = dict(want)
        ids = [] if loggf_ids is None else [int(i) for i in loggf_ids]
        self.loggf_rows = [row_of.get(i, -1) if i not in ids[p + 1:] else -1 for p, i in enumerate(ids)]
        self.n_atomic_pars = len(ids) + self._n_lam_pars
        return len(rows)

    def compute_rf(self, atmosphere, mu=1.0, atm_scale=0):
        """``compute`` with ``get_atomic_rfs``: ``(stokes [.., 4, nlambda], rfs [.., n_atomic_pars, nlambda])``."""
        if not self.loggf_rows:                                      # no log gf parameter: rfs is [n_atomic_pars][nlw] of
            st = self.compute(atmosphere, mu=mu, atm_scale=atm_scale)      # columns no line carries
            return st, np.zeros(st.shape[:-2] + (self.n_atomic_pars, st.shape[-1]))
        if self.ctx.lrf_npar != len(self.loggf_rows):
            self.ctx.set_loggf_rf(self.loggf_rows)
        a = np.asarray(atmosphere, np.float64)
        single = a.ndim == 2
        st, rf = self.ctx.compute1d_rf_batch(a[None] if single else a, mu=mu, atm_scale=atm_scale, lambda_ref=self.lambda_ref,
                                             wght_per_H=self.el.wght_per_H, vmacro_tresh=self.vmacro_tresh)
        full = np.zeros((rf.shape[0], self.n_atomic_pars, rf.shape[1]))
        full[:, :rf.shape[2]] = rf.transpose(0, 2, 1)
        return (st[0], full[0]) if single else (st, full)

    def compute(self, atmosphere, mu=1.0, atm_scale=0, get_scales=False):
        """``atmosphere`` [9+, ndep] or [ncol, 9+, ndep] (pyrh units) -> Stokes [.., 4, nlambda] on ``self.wavelengths``."""
        a = np.asarray(atmosphere, np.float64)
        single = a.ndim == 2
        r = self.ctx.compute1d_batch(a[None] if single else a, mu=mu, atm_scale=atm_scale, lambda_ref=self.lambda_ref,
                                     wght_per_H=self.el.wght_per_H, vmacro_tresh=self.vmacro_tresh, get_scales=get_scales)
        if get_scales:
            return (r[0][0], r[1][0]) if single else r
        return r[0] if single else r

    @property
    def wavelengths(self):
        return self.lam[self.lam != self.lambda_ref]

    def close(self):
        self.ctx.close()


class MultiSession:
    """One host batch over several GPUs of the box from ONE process (BASELINE configs[1]: column-sharded over 1/2/4/8
    GPUs): a Session per device with the same tables; ``compute`` cuts the columns into contiguous blocks and runs them
    concurrently (``rhb200_compute1d_batch_multi``: one host thread per device, no collective -- columns are
    independent).  Results are bit-identical to a single-device run."""

    def __init__(self, cwd, wave, devices, **kwargs):
        if len(set(devices)) != len(devices) or not devices:
            raise ValueError("devices must be a non-empty list of distinct CUDA device indices")
        self.sessions = [Session(cwd, wave, device=d, **kwargs) for d in devices]
        self.devices = list(devices)

    @property
    def wavelengths(self):
        return self.sessions[0].wavelengths

    def compute(self, atmosphere, mu=1.0, atm_scale=0, out=None):
        """``atmosphere`` [ncol, 9+, ndep] (pyrh units; page-locked memory keeps the devices' copies concurrent) ->
        Stokes [ncol, 4, nlambda] on ``wavelengths``."""
        import ctypes as C
        from . import _lib, api
        s0 = self.sessions[0]
        a = np.ascontiguousarray(atmosphere, np.float64)
        ncol, nrow, ndep = a.shape
        nl = len(s0.lam)
        iref = int(np.flatnonzero(s0.lam == s0.lambda_ref)[0])
        st = np.empty((ncol, 4, nl)) if out is None else out
        handles = (C.c_void_p * len(self.sessions))(*[s.ctx.h for s in self.sessions])
        _lib.check(s0.ctx.lib.rhb200_compute1d_batch_multi(
            len(self.sessions), handles, ncol, ndep, nrow, float(mu), int(atm_scale), C.c_void_p(a.ctypes.data), iref,
            float(s0.el.wght_per_H), float(s0.vmacro_tresh) * api.KM_TO_M, _lib.BC_ZERO, _lib.BC_THERMALIZED,
            C.c_void_p(st.ctypes.data), None))
        return st if out is not None else np.delete(st, iref, axis=2)

    def close(self):
        for s in self.sessions:
            s.close()


class ScalesSession:
    """The parsed working directory of ``get_scales`` resident on one GPU (continuum model at ``lam_ref``, chemistry)."""

    def __init__(self, cwd, lam_ref, device=0, atomic_number=None, atomic_abundance=None):
        from . import api, continuum
        kw = read_keywords(cwd)
        self.el = el = read_elements(None, kw, atomic_number, atomic_abundance)
        bg = read_background_model(cwd, kw, el)
        self.lam_ref, self.vmacro_tresh = float(lam_ref), float(kw["VMACRO_TRESH"])
        self.ctx = ctx = api.Context(device)
        empty = ll.LineTable(lines=np.zeros((0, ll.RL_NFIELD)), zq=np.zeros(0, np.int32), zshift=np.zeros(0),
                             zstrength=np.zeros(0), elems=np.zeros((0, ll.RE_NFIELD)), pf=np.zeros((0, len(el.Tpf))),
                             Tpf=el.Tpf, vmicro_char=float(kw["VMICRO_CHAR"]) * 1.0E+03)
        ctx.set_lines(empty)
        ctx.set_wavelengths(np.array([self.lam_ref]))
        ctx.set_continuum(continuum.ContinuumModel(bg), np.array([el.abund[int(p) - 1] for p in bg["atom_pt_index"]]))
        ctx.set_chemistry(bg["ce_nuclei"][:, 1].astype(np.int32), bg["ce_mol"])

    def get_scales(self, atm_scale, atmosphere):
        """``atmosphere`` [ncol, 9, ndep] (row 0 = the depth scale) -> [ncol, 3, ndep]: height, tau_ref, column mass."""
        return self.ctx.get_scales_batch(atmosphere, atm_scale, self.lam_ref, self.el.wght_per_H, self.el.totalAbund,
                                         vmacro_tresh=self.vmacro_tresh)

    def close(self):
        self.ctx.close()


class NeSession:
    """atmos.elements[] resident on one GPU for ``get_ne_from_nH`` (Solve_ne over all elements)."""

    def __init__(self, cwd, device=0):
        from . import api
        kw = read_keywords(cwd)
        el = read_elements(None, kw)
        if not el.abundance_set.all():
            raise NotImplementedError("elements without an abundance: the reference feeds their raw partition functions "
                                      "into Solve_ne (abundance.c:207-215); not reproduced")
        self.ctx = api.Context(device)
        self.ctx.set_elements(*element_table(el), el.Tpf)

    def close(self):
        self.ctx.close()


def _cached(kind, cwd, extra, factory):
    """The same bounded LRU as compute1d's sessions, for the helpers pyrh users call once per pixel (hse, get_scales,
    get_ne_from_nH): keyed on the directory, PYRH_PATH, the mtimes of every input file read and the call's overrides."""
    tob = lambda x: None if x is None else np.asarray(x).tobytes()   # noqa: E731
    key = _session_key(cwd, np.zeros(0), (kind,) + tuple(tob(x) if not isinstance(x, (int, float, str)) else x for x in extra))
    s = _SESSIONS.get(key)
    if s is not None:
        _SESSIONS.move_to_end(key)
        return s
    s = factory()
    _SESSIONS[key] = s
    while len(_SESSIONS) > MAX_SESSIONS:
        _, old = _SESSIONS.popitem(last=False)
        old.close()
    return s


def get_scales(cwd, atm_scale, scale, atmosphere, lam_ref, atomic_number=None, atomic_abundance=None, device=0):
    """Drop-in for ``pyrh.get_scales`` (pyrh.pyx:491-534): ``(tau, height [m], cmass [kg m^-2])`` of one column from
    Background() at ``lam_ref`` and convertScales().  Like the reference it looks at no Kurucz line
    (pyrh_hse.c:441) and takes the depth scale from ``scale`` (``atmosphere`` row 0 is ignored).  The parsed directory
    stays resident between calls (``close_sessions()`` releases it)."""
    s = _cached("get_scales", cwd, (float(lam_ref), atomic_number, atomic_abundance, device),
                lambda: ScalesSession(cwd, lam_ref, device, atomic_number, atomic_abundance))
    a = np.array(atmosphere, np.float64)[None, :9].copy()
    a[0, 0] = scale
    sc = s.get_scales(atm_scale, a)[0]
    return sc[1], sc[0], sc[2]


def element_table(el: Elements):
    """All elements as RHB200_RE_* rows + their ln U rows (for Context.set_elements)."""
    rows, pfrows = [], []
    for e in range(len(el.ID)):
        nst = int(el.nstage[e])
        if nst > ll.RE_MAXSTAGE:
            raise ValueError(f"element {el.ID[e]}: {nst} ionisation stages exceed RE_MAXSTAGE")
        r = np.zeros(ll.RE_NFIELD)
        r[ll.RE_WEIGHT], r[ll.RE_ABUND], r[ll.RE_NSTAGE], r[ll.RE_PFROW] = el.weight[e], el.abund[e], nst, len(pfrows)
        r[ll.RE_IONPOT0:ll.RE_IONPOT0 + nst] = el.ionpot[e]
        pfrows += list(el.pf[e])
        rows.append(r)
    return np.array(rows), np.array(pfrows)


class HseSession:
    """The parsed working directory of ``hse`` resident on one GPU (elements, the 500 nm continuum model, chemistry)."""

    def __init__(self, cwd, device=0, atomic_number=None, atomic_abundance=None, fudge_wave=None, fudge_value=None):
        from . import api, continuum
        kw = read_keywords(cwd)
        self.el = el = read_elements(None, kw, atomic_number, atomic_abundance)
        bg = read_background_model(cwd, kw, el)
        if not el.abundance_set.all():
            raise NotImplementedError("elements without an abundance are not supported by the electron-density solver")
        self.ctx = ctx = api.Context(device)
        empty = ll.LineTable(lines=np.zeros((0, ll.RL_NFIELD)), zq=np.zeros(0, np.int32), zshift=np.zeros(0),
                             zstrength=np.zeros(0), elems=np.zeros((0, ll.RE_NFIELD)), pf=np.zeros((0, len(el.Tpf))),
                             Tpf=el.Tpf, vmicro_char=float(kw["VMICRO_CHAR"]) * 1.0E+03)
        ctx.set_lines(empty)
        ctx.set_wavelengths(np.array([500.0]))                              # pyrh_hse.c:197-200
        ctx.set_elements(*element_table(el), el.Tpf)
        self.model = continuum.ContinuumModel(bg, fudge_wave, fudge_value)
        ctx.set_continuum(self.model, np.array([el.abund[int(p) - 1] for p in bg["atom_pt_index"]]))
        ctx.set_chemistry(bg["ce_nuclei"][:, 1].astype(np.int32), bg["ce_mol"])

    def hse(self, atm_scale, scale, temp, pg_top=0.1):
        """``scale``, ``temp`` [ndep] or [ncol, ndep] -> ne, nHtot, rho, pg (SI, like the reference's arrays)."""
        scale, temp = np.asarray(scale, np.float64), np.asarray(temp, np.float64)
        single = temp.ndim == 1
        out = self.ctx.hse_batch(np.atleast_2d(scale), np.atleast_2d(temp), pg_top, atm_scale, self.el.wght_per_H,
                                 self.el.totalAbund)
        return tuple(o[0] for o in out) if single else out

    def close(self):
        self.ctx.close()


def hse(cwd, atm_scale, scale, temp, pg_top=0.1, fudge_wave=None, fudge_value=None, atomic_number=None,
        atomic_abundance=None, full_output=False, device=0):
    """Drop-in for ``pyrh.hse`` (pyrh.pyx:427-489): ``(ne, nHtot)`` or, with ``full_output``, ``(ne, nHtot, rho, pg)``."""
    s = _cached("hse", cwd, (atomic_number, atomic_abundance, fudge_wave, fudge_value, device),
                lambda: HseSession(cwd, device, atomic_number, atomic_abundance, fudge_wave, fudge_value))
    ne, nH, rho, pg = s.hse(atm_scale, scale, temp, pg_top)
    return (ne, nH, rho, pg) if full_output else (ne, nH)


def get_ne_from_nH(cwd, atm_scale, scale, temperature, nH, device=0):
    """Drop-in for ``pyrh.get_ne_from_nH`` (pyrh.pyx:396-425): electron density [cm^-3] from temperature [K] and total
    hydrogen density [cm^-3] by the LTE ionisation equilibrium of all elements (Solve_ne from scratch, hydrogen in
    LTE).  ``atm_scale`` / ``scale`` are accepted like the reference's and, like there, do not enter the result."""
    s = _cached("get_ne_from_nH", cwd, (device,), lambda: NeSession(cwd, device))
    CUBE_CM = 1.0E-02 * 1.0E-02 * 1.0E-02
    nHtot = np.array([x / CUBE_CM for x in np.asarray(nH, np.float64)])              # pyrh_hse.c:626
    ne = s.ctx.solve_ne(np.asarray(temperature, np.float64), nHtot)
    return np.array([x * CUBE_CM for x in ne])                                       # pyrh_hse.c:647


_SESSIONS: OrderedDict = OrderedDict()   # per-process LRU of resident sessions (bounded: MAX_SESSIONS)
MAX_SESSIONS = 4


class Populations:
    """pyrh.pyx:170-176: what compute1d(get_populations=True) returns per ACTIVE atom."""

    def __init__(self, ID, nlevel, nz, n, nstar):
        self.ID, self.nlevel, self.nz, self.n, self.nstar = ID, nlevel, nz, n, nstar


def _input_files(cwd, kw, path=None):
    """Every file a session reads: a change of any of them (or of PYRH_PATH) must not hit a stale session."""
    cwd = Path(cwd)
    files = [cwd / "keyword.input", cwd / kw["ATOMS_FILE"], cwd / kw["MOLECULES_FILE"]]
    root = pyrh_path(path) / "rh"
    files += [Path(kw.get("ABUND_FILE") or root / "Atoms" / "abundance.input"),
              Path(kw.get("KURUCZ_PF_DATA") or root / "Atoms" / "pf_Kurucz.input")]
    if kw["KURUCZ_DATA"].lower() != "none" and (cwd / kw["KURUCZ_DATA"]).exists():
        files.append(cwd / kw["KURUCZ_DATA"])
        for ln in (cwd / kw["KURUCZ_DATA"]).read_text().splitlines():
            if ln.strip() and ln[0] != "#":
                files.append(cwd / ln.split()[0])
    files += [root / "Atoms" / f for f, _ in _atoms_listed(cwd, kw)]
    if files[2].exists():
        for ln in files[2].read_text().splitlines():
            f = ln.split("#", 1)[0].split()
            if len(f) >= 2 and f[0].endswith(".molecule"):
                files.append(root / "Molecules" / f[0])
    return files


def _session_key(cwd, wave, extra, path=None):
    kw = read_keywords(cwd)
    st = tuple((str(p), p.stat().st_mtime_ns if p.exists() else 0) for p in _input_files(cwd, kw, path))
    return (str(Path(cwd).resolve()), str(pyrh_path(path)), st, np.asarray(wave, np.float64).tobytes(), extra)


def close_sessions():
    """Release every cached session (GPU context, device tables, page-locked staging)."""
    for s in _SESSIONS.values():
        s.close()
    _SESSIONS.clear()


def _get_session(cwd, wave, device, overrides):
    """LRU lookup; the least recently used session is closed when the cache is full, so a caller that varies log gf or
    an abundance from call to call holds at most MAX_SESSIONS contexts."""
    tob = lambda x: None if x is None else np.asarray(x).tobytes()   # noqa: E731
    # log gf overrides are applied IN PLACE to the resident session (Session.set_loggf): not part of the key
    structural = {k: v for k, v in overrides.items() if k not in ("loggf_ids", "loggf_values")}
    key = _session_key(cwd, wave, tuple(tob(structural[k]) for k in sorted(structural)) + (device,))
    s = _SESSIONS.get(key)
    if s is not None:
        _SESSIONS.move_to_end(key)
        if isinstance(s, Session):
            s.set_loggf(overrides.get("loggf_ids"), overrides.get("loggf_values"))
        elif overrides.get("loggf_ids") is not None and s.loggf_key != (tob(overrides["loggf_ids"]), tob(overrides["loggf_values"])):
            s = None                                                # NLTE sessions: rebuilt (the LRU bounds them)
        if s is not None:
            return s
    kw = read_keywords(cwd)
    if any(st == "ACTIVE" for _, st in _atoms_listed(cwd, kw)):
        from . import nlte_host
        s = nlte_host.NlteSession(cwd, wave, device, None, **overrides)
    else:
        s = Session(cwd, wave, device, None, **overrides)
    stale = _SESSIONS.pop(key, None)
    if stale is not None:
        stale.close()
    _SESSIONS[key] = s
    while len(_SESSIONS) > MAX_SESSIONS:
        _, old = _SESSIONS.popitem(last=False)
        old.close()
    return s


def compute1d(cwd, mu, atm_scale, atmosphere, wave, loggf_ids=None, loggf_values=None, lam_ids=None, lam_values=None,
              fudge_wave=None, fudge_value=None, atomic_number=None, atomic_abundance=None, get_atomic_rfs=False,
              get_populations=False, device=0):
    """Drop-in for ``pyrh.compute1d`` (pyrh.pyx:537-668): returns ``(sI, sQ, sU, sV, lam)``; with ACTIVE atoms in
    ``atoms.input`` the NLTE problem is solved first (initScatter, Iterate, the final pass at ``mu``) and
    ``get_populations`` adds a tuple of ``Populations`` (one per ACTIVE atom).  The parsed working directory stays
    resident on the GPU between calls (bounded per-process LRU keyed on the directory, PYRH_PATH, the mtimes of every
    input file read, the wavelength grid and the per-call line / abundance / fudge overrides)."""
    s = _get_session(cwd, wave, device, dict(loggf_ids=loggf_ids, loggf_values=loggf_values, lam_ids=lam_ids,
                                             lam_values=lam_values, fudge_wave=fudge_wave, fudge_value=fudge_value,
                                             atomic_number=atomic_number, atomic_abundance=atomic_abundance))
    nlte = not isinstance(s, Session)
    populations = ()
    if nlte:
        if get_atomic_rfs:
            raise NotImplementedError("get_atomic_rfs with ACTIVE atoms is not ported")
        res = s.compute(atmosphere, mu=mu, atm_scale=atm_scale)
        # atmos.Stokes is TRUE: zeros with STOKES_MODE NO_STOKES, the full Stokes solution after FIELD_FREE iterations
        output = (res["I"], res["Q"], res["U"], res["V"], s.wavelengths)
        if get_populations:
            populations = tuple(Populations(ID, n.shape[0], n.shape[1], n, ns) for ID, n, ns in s.populations(res))
    else:
        if get_atomic_rfs:                                       # pyrh.pyx:658-660: (output, rf.T); lam_ids columns stay 0
            st, rf = s.compute_rf(atmosphere, mu=mu, atm_scale=atm_scale)
        else:
            st = s.compute(atmosphere, mu=mu, atm_scale=atm_scale)
        # rhf1d() sets atmos.Stokes = TRUE unconditionally (pyrh_compute1dray.c:258), so _solveray always fills sQ/sU/sV
        # and sets spec->stokes (pyrh_solveray.c:137-142): in NO_STOKES mode they are arrays of zeros, never None
        output = (st[0], st[1], st[2], st[3], s.wavelengths)
        if get_atomic_rfs:
            output = (output, rf)
    if get_populations:                                          # pyrh.pyx:654-673
        return output, populations
    return output

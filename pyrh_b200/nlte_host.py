"""NLTE front end: builds the flat MALI problem (include/rhb200.h, rhb200_nlte_plan / rhb200_nlte_front) from a working
directory whose ``atoms.input`` marks model atoms ACTIVE -- what the reference does between ``readAtomicModels()`` and
``Iterate()`` on the host, once per directory and wavelength grid instead of once per column.

Reference routines restated here (all integer / table work, bit-identical to the reference's parsed state):

* readAtom for ACTIVE atoms            rh/readatom.c:100-760 (lines with their grids, continua, the collisional section)
* getLambda / getwlambda_line / _cont  rh/getlambda.c:44-219 (line wavelength quadrature, integration weights)
* SortLambda                           rh/sortlambda.c:42-561 (merged grid, Nblue, active sets; cross-sections of the
                                       bound-free continua: hydrogenic Gaunt_bf, or natural spline of the table)
* CollisionRate, reading part          rh/collision.c:450-676 (TEMP / OMEGA / CE / CI / CP / CH / CH0 / CH+ records; the
                                       spline coefficients of splineCoef, rh/spline.c:31-66) -- the rates themselves are
                                       evaluated on the device per column
* getAngleQuad / GaussLeg              rh/rhf1d/anglequad.c:30-55, rh/gaussleg.c:30-65

Per-column work (LTE populations, collisional rates, damping, background, initScatter, Iterate, the final single-mu pass)
runs on the GPU: ``rhb200_nlte_compute1d_batch``.
"""
from __future__ import annotations

import math
from pathlib import Path

import numpy as np

from . import host as H
from . import zeeman

TR_NFIELD = 16
(TR_ATOM, TR_TYPE, TR_I, TR_J, TR_NBLUE, TR_NLAMBDA, TR_AJI, TR_BJI, TR_BIJ, TR_ISOFRAC, TR_WOFF,
 TR_PHIROW, TR_KR, TR_LINEIDX, TR_LAMBDA0) = range(15)

# collision records of the device table -- include/rhb200.h RHB200_CO_*
(CO_ATOM, CO_TYPE, CO_I, CO_J, CO_NT, CO_TOFF, CO_DE) = range(7)
CO_NFIELD = 8
CO_OMEGA, CO_CE, CO_CI, CO_CP, CO_CH, CO_CH0, CO_CHPLUS = range(7)
_CO_KEYS = {"OMEGA": CO_OMEGA, "CE": CO_CE, "CI": CO_CI, "CP": CO_CP, "CH": CO_CH, "CH0": CO_CH0, "CH+": CO_CHPLUS}
_CO_UNPORTED = ("AR85-CHP", "AR85-CHH", "AR85-CEA", "BURGESS", "SHULL82", "BADNELL", "SUMMERS", "AR85-CDI")


def _sq(x):
    return x * x


def _cube(x):
    return x * x * x


def _data_lines(path):
    """getLine(): lines that hold something and do not start with '#' in column 0 (rh/getline.c:31-52)."""
    return [ln for ln in Path(path).read_text().splitlines() if ln.strip() and ln[0] != "#"]


def gauss_leg(x1, x2, n):
    """GaussLeg (rh/gaussleg.c:30-65)."""
    x, w = [0.0] * n, [0.0] * n
    m = (n + 1) // 2
    xm, xl = 0.5 * (x2 + x1), 0.5 * (x2 - x1)
    for i in range(m):
        zz = math.cos(H.PI * (i + 0.75) / (n + 0.5))
        while True:
            p1, p2 = 1.0, 0.0
            for j in range(1, n + 1):
                p3 = p2
                p2 = p1
                p1 = (2.0 * (j - 0.5) * zz * p2 - (j - 1.0) * p3) / j
            pp = n * (zz * p1 - p2) / (zz * zz - 1.0)
            z1 = zz
            zz = z1 - p1 / pp
            dz = abs(p1 / pp)
            if not dz > 3.0E-14:
                break
        x[i] = xm - xl * zz
        w[i] = 2.0 * xl / ((1.0 - zz * zz) * pp * pp)
        x[n - 1 - i] = xm + xl * zz
        w[n - 1 - i] = w[i]
    return np.array(x), np.array(w)


def locate(arr, v):
    """Locate (rh/hunt.c:92-117) for an ascending table; also Hunt() without a usable starting guess."""
    lo, hi = 0, len(arr)
    while hi - lo > 1:
        mid = (hi + lo) >> 1
        if v >= arr[mid]:
            lo = mid
        else:
            hi = mid
    return lo


def hunt(arr, v, ilow):
    """Hunt (rh/hunt.c:17-78) for an ascending table, started from the guess ``ilow``.  Not the same function as
    Locate: when it hunts DOWN from the guess and ``v`` equals a table entry exactly, the bracket closes below that
    entry and the index returned is one less -- SortLambda's Nred inherits this (a line whose red-most wavelength is
    found that way loses its last point), so it is reproduced, not repaired."""
    n = len(arr)
    if ilow <= 0 or ilow > n - 1:
        return locate(arr, v)
    inc = 1
    if v >= arr[ilow]:
        ihigh = ilow + inc
        if ilow == n - 1:
            return ilow
        while v >= arr[ihigh]:
            ilow = ihigh
            inc += inc
            ihigh = ilow + inc
            if ihigh >= n:
                ihigh = n
                break
    else:
        ihigh = ilow
        if ilow == 0:
            return ilow
        while v <= arr[ilow]:
            ihigh = ilow
            inc += inc
            ilow = ihigh - inc
            if ilow <= 0:
                ilow = 0
                break
    while ihigh - ilow > 1:
        mid = (ihigh + ilow) >> 1
        if v >= arr[mid]:
            ilow = mid
        else:
            ihigh = mid
    return ilow


def spline_coef(x, y):
    """splineCoef (rh/spline.c:31-66): second derivatives M of the natural cubic spline."""
    N = len(x)
    q, u, M = [0.0] * N, [0.0] * N, [0.0] * N
    hj = x[1] - x[0]
    D = (y[1] - y[0]) / hj
    for j in range(1, N - 1):
        hj1 = x[j + 1] - x[j]
        mu = hj / (hj + hj1)
        D1 = (y[j + 1] - y[j]) / hj1
        p = mu * q[j - 1] + 2
        q[j] = (mu - 1) / p
        u[j] = ((D1 - D) * 6 / (hj + hj1) - mu * u[j - 1]) / p
        hj, D = hj1, D1
    M[N - 1] = 0.0
    for j in range(N - 2, -1, -1):
        M[j] = q[j] * M[j + 1] + u[j]
    return M


def spline_eval(xt, yt, M, xs):
    """splineEval (rh/spline.c:70-100) for an ascending or descending table."""
    N = len(xt)
    ascend = xt[1] > xt[0]
    xmin, xmax = (xt[0], xt[N - 1]) if ascend else (xt[N - 1], xt[0])
    out = []
    for x in xs:
        if x <= xmin:
            out.append(yt[0] if ascend else yt[N - 1])
        elif x >= xmax:
            out.append(yt[N - 1] if ascend else yt[0])
        else:
            if ascend:
                j = locate(xt, x)
            else:
                lo, hi = 0, N
                while hi - lo > 1:
                    mid = (hi + lo) >> 1
                    if x <= xt[mid]:
                        lo = mid
                    else:
                        hi = mid
                j = lo
            hj = xt[j + 1] - xt[j]
            fx = (x - xt[j]) / hj
            fx1 = 1 - fx
            out.append(fx1 * yt[j] + fx * yt[j + 1] +
                       (fx1 * (_sq(fx1) - 1) * M[j] + fx * (_sq(fx) - 1) * M[j + 1]) * _sq(hj) / 6.0)
    return out


def gaunt_bf(lam, n_eff, charge):
    """Gaunt_bf (rh/hydrogen.c:269-281)."""
    x = ((H.HPLANCK * H.CLIGHT) / (lam * H.NM_TO_M)) / (H.E_RYDBERG * _sq(charge))
    x3 = math.pow(x, 0.33333333)
    nsqx = 1.0 / (_sq(n_eff) * x)
    return 1.0 + 0.1728 * x3 * (1.0 - 2.0 * nsqx) - 0.0496 * _sq(x3) * (1.0 - (1.0 - nsqx) * 0.66666667 * nsqx)


def lande(S, L, J):
    """Lande (rh/zeeman.c:138-144)."""
    return 0.0 if J == 0.0 else 1.5 + (S * (S + 1.0) - L * (L + 1)) / (2.0 * J * (J + 1.0))


# ------------------------------------------------------------------------------------------- active atoms
def read_active_atom(atom_file, kw, moving=True):
    """readAtom() of an ACTIVE atom (rh/readatom.c:100-760): levels, lines incl. their wavelength quadrature
    (getLambda), continua, and the raw records of the collisional section."""
    data = _data_lines(atom_file)
    base = H.read_atom(atom_file)
    nlevel, nline, ncont, nfixed = (int(x) for x in data[1].split()[:4])
    if nfixed > 0:
        raise NotImplementedError(f"{atom_file}: fixed transitions (FixedRate, rh/fixedrate.c) are not ported")
    E, g, label, stage = base["E"], base["g"], base["label"], base["stage"]
    if stage[nlevel - 1] != stage[nlevel - 2] + 1:
        raise ValueError(f"{atom_file}: atomic model does not have an overlying continuum (readatom.c:177-182)")
    prd_on = int(kw.get("PRD_N_MAX_ITER", "3")) > 0
    vmicro_char = float(kw["VMICRO_CHAR"]) * 1.0E+03
    B_char = float(kw.get("B_STRENGTH_CHAR", "0.0"))
    pos = 2 + nlevel
    lines = []
    for kr in range(nline):
        f = data[pos].split()
        pos += 1
        ln = dict(base["lines"][kr])
        ln.update(kr=kr, Nlambda_in=int(f[4]), symmetric="ASYMM" not in f[5], qcore=float(f[6]),
                  g_Lande_eff=float(f[15]) if len(f) > 15 and H._scan_float(f[15])[0] is not None else 0.0,
                  isotope_frac=1.0)
        ln["PRD"] = bool("PRD" in f[3] and prd_on)                       # readatom.c:255-258
        if "COMPOSIT" in f[3]:
            pos += 1 + int(data[pos].split()[0])
        i, j = ln["i"], ln["j"]
        # atmos.Stokes is TRUE on the pyrh path (pyrh_compute1dray.c:258): readatom.c:340-360
        di, dj = zeeman.determinate(label[i], g[i]), zeeman.determinate(label[j], g[j])
        ln["polarizable"] = bool(ln["g_Lande_eff"] != 0.0 or (di[0] and dj[0] and abs(dj[4] - di[4]) <= 1.0))
        if ln["polarizable"] and len(ln["c_shift"]) > 1:
            raise ValueError(f"{atom_file}: cannot treat composite line {j}->{i} with polarization (readatom.c:346-350)")
        ln["det"] = (di, dj)
        lines.append(ln)
    conts = []
    for kr, (j, i, alpha0, hyd, lambda0, lam, alp) in enumerate(H.read_atom_continua(atom_file, base)):
        conts.append(dict(kr=kr, i=i, j=j, alpha0=alpha0, hydrogenic=bool(hyd), lambda0=lambda0, lam=list(lam),
                          alpha=list(alp), isotope_frac=1.0))
    # skip the continuum records to reach the collisional section
    for _ in range(ncont):
        f = data[pos].split()
        pos += 1
        if "EXPLICIT" in f[4]:
            pos += int(f[3])
    for ln in lines:                                                   # readatom.c:497
        _get_lambda(ln, moving, vmicro_char, B_char)
    return dict(ID=base["ID"], E=E, g=g, label=label, stage=stage, abo_level=base["abo_level"], lines=lines,
                continua=conts, coll=data[pos:], file=str(atom_file))


def _effective_lande(ln):
    """effectiveLande (rh/zeeman.c:106-133)."""
    if ln["g_Lande_eff"] != 0.0:
        return ln["g_Lande_eff"]
    di, dj = ln["det"]
    if di[0] and dj[0]:
        g_l, g_u = lande(di[2], di[3], di[4]), lande(dj[2], dj[3], dj[4])
        return 0.5 * (g_u + g_l) + 0.25 * (g_u - g_l) * (dj[4] * (dj[4] + 1.0) - di[4] * (di[4] + 1.0))
    return 0.0


def _get_lambda(ln, moving, vmicro_char, B_char):
    """getLambda (rh/getlambda.c:44-166): fills ln['lam'] and may clear ln['symmetric']."""
    if ln["qcore"] <= 0.0 or ln["qwing"] <= 0.0:
        raise ValueError(f"line {ln['j']}->{ln['i']}: qcore or qwing is negative or zero (getlambda.c:55-59)")
    ncomp = len(ln["c_shift"])
    if ln["polarizable"] or moving or ncomp > 1:                        # atmos.Stokes && line->polarizable
        ln["symmetric"] = False
    nin = ln["Nlambda_in"]
    if ln["symmetric"]:
        Nlambda = nin if nin % 2 else nin + 1
    else:
        Nlambda = nin // 2 if nin % 2 else (nin + 1) // 2
    beta = 1.0 if ln["qwing"] <= 2.0 * ln["qcore"] else ln["qwing"] / (2.0 * ln["qcore"])
    y = beta + math.sqrt(_sq(beta) + (beta - 1.0) * Nlambda + 2.0 - 3.0 * beta)
    b = 2.0 * math.log(y) / (Nlambda - 1)
    a = ln["qwing"] / (Nlambda - 2.0 + _sq(y))
    q = [a * (la + (math.exp(b * la) - 1.0)) for la in range(Nlambda)]
    if ln["polarizable"]:
        g_eff = _effective_lande(ln)
        qB_char = g_eff * (H.Q_ELECTRON / (4.0 * H.PI * H.M_ELECTRON)) * (ln["lambda0"] * H.NM_TO_M) * B_char / vmicro_char
        NB = locate(q, qB_char / 2.0)
        qB_shift = 2 * q[NB]
        q = q + [0.0] * (2 * NB)
        for la in range(NB + 1, 2 * NB + 1):
            q[la] = qB_shift - a * (2 * NB - la + (math.exp(b * (2 * NB - la)) - 1.0))
        for la in range(2 * NB + 1, Nlambda + 2 * NB):
            q[la] = qB_shift + a * (la - 2 * NB + (math.exp(b * (la - 2 * NB)) - 1.0))
        Nlambda += 2 * NB
    q_to_lambda = ln["lambda0"] * (vmicro_char / H.CLIGHT)
    if ln["symmetric"]:
        lam = [ln["lambda0"] + q_to_lambda * q[la] for la in range(Nlambda)]
    else:
        n_tot = (2 * Nlambda - 1) * ncomp
        lam = [0.0] * n_tot
        for n in range(ncomp):
            Nmid = n * (2 * Nlambda - 1) + Nlambda - 1
            lambda0 = ln["lambda0"] + ln["c_shift"][n]
            lam[Nmid] = lambda0
            for la in range(1, Nlambda):
                dl = q_to_lambda * q[la]
                lam[Nmid - la] = lambda0 - dl
                lam[Nmid + la] = lambda0 + dl
        if ncomp > 1:
            lam.sort()
    ln["lam"] = lam


def _wlambda(lam, la):
    n = len(lam)
    if la == 0:
        return 0.5 * (lam[la + 1] - lam[la])
    if la == n - 1:
        return 0.5 * (lam[la] - lam[la - 1])
    return 0.5 * (lam[la + 1] - lam[la - 1])


# ------------------------------------------------------------------------------------------- collisions
def collision_table(atoms):
    """The collisional sections of the ACTIVE atoms as device records (rows [n, CO_NFIELD] in file order, which is the
    order of the reference's += into atom->C) plus the concatenated {T, coefficient, spline M} tables."""
    rows, tT, tC, tM = [], [], [], []
    for a, at in enumerate(atoms):
        T = None
        nlev = len(at["E"])
        for raw in at["coll"]:
            tok = [t for t in raw.split(" ") if t.strip()]            # strtok(inputLine, " ")
            if not tok:
                continue
            key = tok[0].strip()
            if key == "TEMP":
                nitem = int(tok[1])
                T = [float(x) for x in tok[2:2 + nitem]]
                if len(T) != nitem:
                    raise ValueError(f"{at['file']}: TEMP record with {len(T)} of {nitem} items")
            elif key in _CO_KEYS:
                if T is None:
                    raise ValueError(f"{at['file']}: {key} record before any TEMP record")
                i1, i2 = int(tok[1]), int(tok[2])
                coef = [float(x) for x in tok[3:3 + len(T)]]
                if len(coef) != len(T):
                    raise ValueError(f"{at['file']}: read {len(coef)}, not {len(T)} items (keyword = {key})")
                i, j = min(i1, i2), max(i1, i2)
                if not (0 <= i < j < nlev):
                    raise ValueError(f"{at['file']}: collision record {key} {i1} {i2}: level out of range")
                M = spline_coef(T, coef) if len(T) > 2 else [0.0] * len(T)
                r = np.zeros(CO_NFIELD)
                r[CO_ATOM], r[CO_TYPE], r[CO_I], r[CO_J], r[CO_NT], r[CO_TOFF] = a, _CO_KEYS[key], i, j, len(T), len(tT)
                r[CO_DE] = at["E"][j] - at["E"][i]
                if key == "OMEGA":                                    # factors the device kernel divides / multiplies by:
                    r[7] = at["g"][j]                                   # atom->g[j], collision.c:693
                elif key == "CE":
                    r[7] = at["g"][i] / at["g"][j]                      # gij, :700
                tT += T; tC += coef; tM += M
                rows.append(r)
            elif "END" in key:
                break
            elif key in _CO_UNPORTED:
                raise NotImplementedError(f"{at['file']}: collision keyword {key} is not ported (collision.c:516-936)")
            else:
                raise ValueError(f"{at['file']}: unknown collision keyword !{key}! (collision.c:660)")
    return (np.array(rows).reshape(-1, CO_NFIELD), np.array(tT, np.float64), np.array(tC, np.float64),
            np.array(tM, np.float64))


# ------------------------------------------------------------------------------------------- SortLambda
def build_plan(atoms, wave, lambda_ref, nrays, kw, mu_single=None):
    """SortLambda (rh/sortlambda.c:42-561) + the flat tables of rhb200_nlte_plan.  ``atoms``: the ACTIVE atoms in the
    order of atoms.input (= atmos.activeatoms).  Returns a dict of numpy arrays with the keys of NlteProblem."""
    spect = ([float(lambda_ref)] if lambda_ref > 0.0 else []) + [float(x) for x in wave]
    for at in atoms:
        for c in at["continua"]:
            spect += c["lam"]
        for ln in at["lines"]:
            spect += ln["lam"]
    spect.sort()                                                        # qsort(qsascend), :198
    lam = [spect[0]]
    for x in spect[1:]:                                                 # :202-208
        if x > lam[-1]:
            lam.append(x)
    Ns = len(lam)
    as_lists = [[] for _ in range(Ns)]
    rows, wl, wlam, alpha = [], [], [], []
    phirow = nline = 0
    trans_index = {}
    # the device table lists each atom's lines first, then its continua; the active sets keep SortLambda's order
    # (continua first, then lines, atom by atom: :297-409)
    for a, at in enumerate(atoms):
        Nred = 0                                                        # :299; carried from transition to transition
        for c in at["continua"]:                                        # Hunt() with its starting guesses, :311-315, 375-379
            c["Nblue"] = hunt(lam, c["lam"][0], 0)
            Nred = hunt(lam, c["lam"][-1], Nred)
            c["Nlambda"] = Nred - c["Nblue"] + 1
        for ln in at["lines"]:
            ln["Nblue"] = hunt(lam, ln["lam"][0], 0)
            Nred = hunt(lam, ln["lam"][-1], Nred)
            ln["Nlambda"] = Nred - ln["Nblue"] + 1
        for ln in at["lines"]:
            Nblue, Nla = ln["Nblue"], ln["Nlambda"]
            grid = lam[Nblue:Nblue + Nla]
            dopp = H.CLIGHT / ln["lambda0"]
            if ln["symmetric"]:
                dopp *= 2.0
            r = np.zeros(TR_NFIELD)
            r[[TR_ATOM, TR_TYPE, TR_I, TR_J, TR_NBLUE, TR_NLAMBDA]] = [a, 0, ln["i"], ln["j"], Nblue, Nla]
            r[[TR_AJI, TR_BJI, TR_BIJ, TR_ISOFRAC]] = [ln["Aji"], ln["Bji"], ln["Bij"], ln["isotope_frac"]]
            r[TR_WOFF], r[TR_PHIROW], r[TR_KR], r[TR_LINEIDX], r[TR_LAMBDA0] = len(wl), phirow, ln["kr"], nline, ln["lambda0"]
            wl += grid
            wlam += [_wlambda(grid, la) * dopp for la in range(Nla)]      # getwlambda_line, getlambda.c:170-189
            alpha += [0.0] * Nla
            phirow += 2 * nrays * Nla
            nline += 1
            trans_index[(a, 0, ln["kr"])] = len(rows)
            rows.append(r)
        for c in at["continua"]:
            Nblue, Nla = c["Nblue"], c["Nlambda"]
            grid = lam[Nblue:Nblue + Nla]
            if c["hydrogenic"]:                                         # :332-347
                Z = at["stage"][c["j"]]
                n_eff = Z * math.sqrt(H.E_RYDBERG / (at["E"][c["j"]] - at["E"][c["i"]]))
                gbf_0 = gaunt_bf(c["lambda0"], n_eff, Z)
                alp = [c["alpha0"] * gaunt_bf(x, n_eff, Z) / gbf_0 * _cube(x / c["lambda0"]) for x in grid]
            else:                                                       # :348-360
                alp = spline_eval(c["lam"], c["alpha"], spline_coef(c["lam"], c["alpha"]), grid)
            r = np.zeros(TR_NFIELD)
            r[[TR_ATOM, TR_TYPE, TR_I, TR_J, TR_NBLUE, TR_NLAMBDA]] = [a, 1, c["i"], c["j"], Nblue, Nla]
            r[TR_WOFF], r[TR_PHIROW], r[TR_KR], r[TR_LINEIDX] = len(wl), -1, c["kr"], -1
            wl += grid
            alpha += alp
            wlam += [_wlambda(grid, la) for la in range(Nla)]             # getwlambda_cont, getlambda.c:193-206
            trans_index[(a, 1, c["kr"])] = len(rows)
            rows.append(r)
        for c in at["continua"]:
            for ns in range(c["Nblue"], c["Nblue"] + c["Nlambda"]):
                as_lists[ns].append(trans_index[(a, 1, c["kr"])])
        for ln in at["lines"]:
            for ns in range(ln["Nblue"], ln["Nblue"] + ln["Nlambda"]):
                as_lists[ns].append(trans_index[(a, 0, ln["kr"])])
    N_MAX_OVERLAP = 100                                                 # spectrum.h:17
    for ns, lst in enumerate(as_lists):
        for a in range(len(atoms)):
            if sum(1 for t in lst if rows[t][TR_ATOM] == a) >= N_MAX_OVERLAP:
                raise ValueError(f"too many overlapping transitions at wavelength {ns} (sortlambda.c:318-324)")
    first = [0]
    flat = []
    for lst in as_lists:
        flat += lst
        first.append(len(flat))
    muz, wmu = gauss_leg(0.0, 1.0, nrays) if mu_single is None else (np.array([mu_single]), np.array([1.0]))
    return dict(lam=np.array(lam), muz=muz, wmu=wmu, atom_nlevel=np.array([len(at["E"]) for at in atoms], np.int32),
                trans=np.array(rows).reshape(-1, TR_NFIELD), tr_lambda=np.array(wl), tr_wlambda=np.array(wlam),
                tr_alpha=np.array(alpha), as_first=np.array(first, np.int32), as_trans=np.array(flat, np.int32),
                nphirow=phirow, nline=nline)


def single_mu_plan(plan, mu):
    """The plan of _solveray()'s final pass (pyrh_solveray.c:84-106): one ray at ``mu`` with weight 1; the profile rows
    shrink to 2 * Nlambda per line."""
    p = dict(plan)
    tr = plan["trans"].copy()
    row = 0
    for t in tr:
        if t[TR_TYPE] == 0:
            t[TR_PHIROW] = row
            row += 2 * int(t[TR_NLAMBDA])
    p.update(trans=tr, muz=np.array([float(mu)]), wmu=np.array([1.0]), nphirow=row)
    return p


# ------------------------------------------------------------------------------------------- session
class NlteSession:
    """A working directory with ACTIVE atoms + a wavelength grid, resident on one GPU: the parsed state ``rhf1d()``
    rebuilds on every call (readAtomicModels, SortLambda, the collisional data, the Kurucz / passive / molecular line
    tables and the continuum model of the merged grid).  ``compute`` then runs everything that depends on the column
    on the device (``rhb200_nlte_compute1d_batch``).  ``exact_rates=True`` accumulates the radiative rates in the
    reference's own order (``rhb200_nlte_set_exact_rates``): populations are then bit-identical to ``rhf1d()`` on every
    column; the default fixed-partition sums agree to ~1e-9 on columns that converge normally and drift (up to 1e-4)
    only on columns the MALI/Ng iteration itself struggles with for 50+ iterations."""

    def __init__(self, cwd, wave, device=0, path=None, loggf_ids=None, loggf_values=None, lam_ids=None, lam_values=None,
                 fudge_wave=None, fudge_value=None, atomic_number=None, atomic_abundance=None, exact_rates=False):
        import ctypes as C
        from . import api, continuum, nlte, _lib
        self.cwd = Path(cwd)
        self.exact_rates = bool(exact_rates)
        tob = lambda x: None if x is None else np.asarray(x).tobytes()   # noqa: E731
        self.loggf_key = (tob(loggf_ids), tob(loggf_values))
        kw = self.kw = H.read_keywords(cwd)
        self.stokes_mode = kw["STOKES_MODE"].upper()
        if self.stokes_mode not in ("NO_STOKES", "FIELD_FREE", "FULL_STOKES", "POLARIZATION_FREE"):
            raise ValueError("STOKES_MODE = %s (readvalue.c:262-283 knows NO_STOKES, FIELD_FREE, POLARIZATION_FREE, FULL_STOKES)"
                             % self.stokes_mode)
        H.refuse_unported_keywords(kw)
        if H._true(kw["MAGNETO_OPTICAL"]):
            raise NotImplementedError("MAGNETO_OPTICAL = TRUE is refused (the reference overflows chip_c there, readj.c:328)")
        if H._true(kw.get("DO_FUDGE", "FALSE")) and fudge_wave is None:
            raise NotImplementedError("DO_FUDGE = TRUE without fudge_wave / fudge_value is undefined in the reference")
        if H._true(kw["RLK_SCATTER"]):
            raise NotImplementedError("RLK_SCATTER = TRUE with ACTIVE atoms is not ported")
        if float(kw["VMACRO_TRESH"]) > 0.0:
            raise NotImplementedError("VMACRO_TRESH > 0 with ACTIVE atoms (static columns: symmetric line grids) is not ported")
        for key in ("N_MAX_ITER", "ITER_LIMIT", "NRAYS"):
            if key not in kw:
                raise ValueError(f"keyword.input: {key} is required (readinput.c:58-97)")
        listed = H._atoms_listed(cwd, kw)
        self.el = el = H.read_elements(path, kw, atomic_number, atomic_abundance)
        root = H.pyrh_path(path) / "rh" / "Atoms"
        self.active_index = [m for m, (_, st) in enumerate(listed) if st == "ACTIVE"]
        self._check_initial_solution(cwd, kw)
        self.atoms = [read_active_atom(root / listed[m][0], kw) for m in self.active_index]
        self.lambda_ref = float(kw["LAMBDA_REF"])
        self.nrays = int(kw["NRAYS"])
        self.plan = build_plan(self.atoms, wave, self.lambda_ref, self.nrays, kw)
        self.lam = self.plan["lam"]
        self.iref = int(np.flatnonzero(self.lam == self.lambda_ref)[0]) if self.lambda_ref > 0.0 else -1
        if self.iref < 0:
            raise NotImplementedError("LAMBDA_REF = 0: convertScales needs the reference wavelength")
        bg = self.background = H.read_background_model(cwd, kw, el, path, allow_active=True)
        self.lt = H.read_kurucz_lines(cwd, kw, el, loggf_ids, loggf_values, lam_ids, lam_values, path)
        mlines, msel, mzee = H.molecular_line_table(cwd, kw, el, path)
        self.ctx = ctx = api.Context(device)
        if self.exact_rates:      # rate sums in the reference's order: bit-identical populations on every column, ~2x slower
            _lib.check(ctx.lib.rhb200_nlte_set_exact_rates(ctx.h, 1))
        ctx.set_lines(self.lt, magneto_optical=False, rlkscatter=False)
        if len(mlines):
            ctx.set_molecular_lines(mlines, msel, *mzee)
        lev = bg["ct_lev"]
        first = [int(np.flatnonzero(lev[:, 0] == a)[0]) for a in range(len(listed))]
        if H._true(kw.get("ALLOW_PASSIVE_BB", "TRUE")):
            ctx.set_passive_lines(*H.passive_line_table(cwd, kw, el, first, path))
        ctx.set_model_lines(H.model_line_rows(cwd, kw, el, self.lt.elem_rows, path))
        ctx.set_stokes_mode("NO_STOKES")
        ctx.set_wavelengths(self.lam)
        ctx.set_solvers(kw["S_INTERPOLATION"], kw["S_INTERPOLATION_STOKES"])
        abundance = np.array([el.abund[int(p) - 1] for p in bg["atom_pt_index"]])
        self.model = continuum.ContinuumModel(bg, fudge_wave, fudge_value)
        ctx.set_continuum(self.model, abundance)
        ctx.set_chemistry(bg["ce_nuclei"][:, 1].astype(np.int32), bg["ce_mol"])
        flags = np.zeros(len(self.lam), np.int32)
        _lib.check(ctx.lib.rhb200_get_wavelength_flags(ctx.h, flags.ctypes.data_as(_lib.ip)))
        self.plan["bg_hasline"] = (flags & 1).astype(np.int32)
        # the ACTIVE lines' Damping() constants, and the collisional records
        self.line_rows = np.ascontiguousarray(H.passive_line_table(cwd, kw, el, first, path, status="ACTIVE")[0])
        assert len(self.line_rows) == self.plan["nline"]
        self.coll, self.coll_T, self.coll_C, self.coll_M = collision_table(self.atoms)
        self.hdr = dict(Nspect=len(self.lam), Nrays=self.nrays, Natom=len(self.atoms), moving=1,
                        Ngorder=int(kw.get("NG_ORDER", "0")), Ngdelay=int(kw.get("NG_DELAY", "0")),
                        Ngperiod=int(kw.get("NG_PERIOD", "1")), isum=int(kw.get("I_SUM", "0")),
                        NmaxIter=int(kw["N_MAX_ITER"]), iterLimit=float(kw["ITER_LIMIT"]),
                        NmaxScatter=int(kw["N_MAX_SCATTER"]), bc_top=_lib.BC_ZERO, bc_bottom=_lib.BC_THERMALIZED)
        self._C, self._nlte, self._lib = C, nlte, _lib
        self.vmacro_tresh = float(kw["VMACRO_TRESH"])
        self.ctx.set_gravity(self.el.totalAbund)
        self.IDs = [at["ID"] for at in self.atoms]
        # angle-averaged PRD (redistribute.c, scatter.c:51-290): line->PRD in line-index order + the keywords
        self.line_prd = np.ascontiguousarray([int(ln["PRD"]) for at in self.atoms for ln in at["lines"]], np.int32)
        self.prd_nmax, self.prd_limit = int(kw.get("PRD_N_MAX_ITER", "3")), float(kw.get("PRD_ITER_LIMIT", "1.0E-2"))
        if self.line_prd.any():
            if H._true(kw.get("PRD_ANGLE_DEP", "FALSE")) or H._true(kw.get("XRD", "FALSE")):
                raise NotImplementedError("PRD_ANGLE_DEP / XRD = TRUE (angle-dependent and cross redistribution) are not ported")
            if int(kw.get("PRD_NG_ORDER", "0")) > 0:
                raise NotImplementedError("PRD_NG_ORDER > 0 (Ng acceleration of the PRD profile ratio) is not ported")
        # FIELD_FREE: Zeeman patterns of the polarizable ACTIVE lines (Zeeman(), zeeman.c:186-281), line-index order
        self.line_pol, self.line_zoff, zq, zs, zt = [], [0], [], [], []
        for at in self.atoms:
            for ln in at["lines"]:
                pol = bool(ln["polarizable"]) and self.stokes_mode != "NO_STOKES"
                if pol:
                    if len(ln["c_shift"]) > 1:
                        raise ValueError("cannot treat composite line with polarization (readatom.c:355-359)")
                    q, sh, st = zeeman.zeeman(at["label"][ln["i"]], at["g"][ln["i"]], at["label"][ln["j"]], at["g"][ln["j"]],
                                              ln["g_Lande_eff"])
                    zq += list(q); zs += list(sh); zt += list(st)
                self.line_pol.append(int(pol))
                self.line_zoff.append(len(zq))
        assert len(self.line_pol) == self.plan["nline"]
        self.line_pol = np.ascontiguousarray(self.line_pol, np.int32)
        self.line_zoff = np.ascontiguousarray(self.line_zoff, np.int32)
        self.zq = np.ascontiguousarray(zq if zq else [0], np.int32)
        self.zshift = np.ascontiguousarray(zs if zs else [0.0], np.float64)
        self.zstrength = np.ascontiguousarray(zt if zt else [0.0], np.float64)

    @staticmethod
    def _check_initial_solution(cwd, kw):
        for ln in (Path(cwd) / kw["ATOMS_FILE"]).read_text().splitlines():
            f = ln.split("#", 1)[0].split()
            if len(f) >= 3 and f[0].endswith(".atom") and f[1].upper() == "ACTIVE" and "LTE_POPULATIONS" not in f[2].upper():
                raise NotImplementedError(f"{f[0]}: initial solution {f[2]} (initial_xdr.c:293-350) is not ported; "
                                          "use LTE_POPULATIONS")

    @property
    def wavelengths(self):
        return self.lam[self.lam != self.lambda_ref]

    def _plan_struct(self, plan, ndep, keep):
        nl, ip, dp = self._nlte, self._lib.ip, self._lib.dp
        f64 = lambda x: np.ascontiguousarray(x, np.float64)   # noqa: E731
        i32 = lambda x: np.ascontiguousarray(x, np.int32)     # noqa: E731
        a = dict(lam=f64(plan["lam"]), muz=f64(plan["muz"]), wmu=f64(plan["wmu"]), atom_nlevel=i32(plan["atom_nlevel"]),
                 trans=f64(plan["trans"]), tr_lambda=f64(plan["tr_lambda"]), tr_wlambda=f64(plan["tr_wlambda"]),
                 tr_alpha=f64(plan["tr_alpha"]), as_first=i32(plan["as_first"]), as_trans=i32(plan["as_trans"]),
                 bg_hasline=i32(self.plan["bg_hasline"]))
        keep.append(a)
        ptr = lambda x: x.ctypes.data_as(ip if x.dtype == np.int32 else dp)   # noqa: E731
        h = self.hdr
        return nl.PlanStruct(h["Nspect"], len(a["muz"]), int(ndep), h["Natom"], a["trans"].shape[0], h["moving"],
                             h["Ngorder"], h["Ngdelay"], h["Ngperiod"], h["isum"], h["bc_top"], h["bc_bottom"],
                             len(a["tr_lambda"]), int(plan["nphirow"]), int(plan["nline"]),
                             *[ptr(a[k]) for k in ("lam", "muz", "wmu", "atom_nlevel", "trans", "tr_lambda", "tr_wlambda",
                                                   "tr_alpha", "as_first", "as_trans", "bg_hasline")])

    def compute(self, atmosphere, mu=1.0, atm_scale=0, get_scales=False):
        """``atmosphere`` [9+, ndep] or [ncol, 9+, ndep] (pyrh units) -> dict(I [.., Nspect-1] on ``wavelengths``,
        n, nstar [.., sum Nlevel, ndep] (ACTIVE atoms concatenated in the order of atoms.input), niter [..])."""
        C, nl, lib = self._C, self._nlte, self._lib
        a = np.ascontiguousarray(atmosphere, np.float64)
        single = a.ndim == 2
        if single:
            a = a[None]
        ncol, nrow, ndep = a.shape
        cache = self.__dict__.setdefault("_structs", {})          # the C structs (and the arrays they point into) per (mu, ndep)
        if (float(mu), ndep) not in cache:
            keep = []
            plan = self._plan_struct(self.plan, ndep, keep)
            plan1 = self._plan_struct(single_mu_plan(self.plan, mu), ndep, keep)
            model = np.ascontiguousarray(self.active_index, np.int32)
            tabs = [np.ascontiguousarray(x, np.float64) for x in (self.coll, self.coll_T, self.coll_C, self.coll_M)]
            fr = nl.FrontStruct(model.ctypes.data_as(lib.ip), len(self.coll), len(self.coll_T),
                                *[x.ctypes.data_as(lib.dp) for x in tabs],
                                self.line_rows.ctypes.data_as(lib.dp), self.hdr["NmaxScatter"], self.hdr["NmaxIter"],
                                self.hdr["iterLimit"], C.pointer(plan1), {"NO_STOKES": 0, "FIELD_FREE": 1, "FULL_STOKES": 2, "POLARIZATION_FREE": 3}[self.stokes_mode],
                                self.line_pol.ctypes.data_as(lib.ip), self.line_zoff.ctypes.data_as(lib.ip),
                                self.zq.ctypes.data_as(lib.ip), self.zshift.ctypes.data_as(lib.dp),
                                self.zstrength.ctypes.data_as(lib.dp), self.line_prd.ctypes.data_as(lib.ip), self.prd_nmax,
                                self.prd_limit)
            cache.clear()
            cache[(float(mu), ndep)] = (plan, plan1, fr, keep, model, tabs)
        plan, plan1, fr = cache[(float(mu), ndep)][:3]
        Ns, nlev = len(self.lam), int(np.sum(self.plan["atom_nlevel"]))
        spec = np.zeros((ncol, Ns)); n = np.zeros((ncol, nlev, ndep)); nstar = np.zeros((ncol, nlev, ndep))
        niter, passes = np.zeros(ncol, np.int32), np.zeros((ncol, 2), np.int32)
        scales = np.zeros((ncol, 3, ndep)) if get_scales else None
        vp = lambda x: None if x is None else C.c_void_p(x.ctypes.data)   # noqa: E731
        quv = np.zeros((ncol, 3, Ns))
        lib.check(self.ctx.lib.rhb200_nlte_compute1d_stokes_batch(
            self.ctx.h, C.byref(plan), C.byref(fr), ncol, ndep, nrow, float(mu), int(atm_scale), vp(a), self.iref,
            float(self.el.wght_per_H), self.vmacro_tresh, vp(spec), vp(quv), vp(n), vp(nstar), vp(niter), vp(passes), vp(scales)))
        user = self.lam != self.lambda_ref
        I = spec[:, user]
        out = dict(I=I, Q=quv[:, 0][:, user], U=quv[:, 1][:, user], V=quv[:, 2][:, user], n=n, nstar=nstar, niter=niter, passes=passes)
        if get_scales:
            out["scales"] = scales
        if single:
            out = {k: v[0] for k, v in out.items()}
        return out

    def ray_points(self, res, ndep):
        """Formal-solution ray-points of a ``compute`` result (SURVEY 8(d) unit of work): depth steps of every ray of every
        solveSpectrum pass -- per wavelength Nrays x 2 where the wavelength is angle dependent (line present: both
        directions), else Nrays (Feautrier) -- times initScatter passes + iterations + passes after Iterate(), plus the
        single-mu final pass."""
        hasline = np.zeros(len(self.lam), bool)
        hasline[:] = self.plan["bg_hasline"] != 0
        tr = self.plan["trans"]
        for t in tr[tr[:, TR_TYPE] == 0]:
            hasline[int(t[TR_NBLUE]):int(t[TR_NBLUE]) + int(t[TR_NLAMBDA])] = True
        per_pass = int(np.sum(np.where(hasline, 2, 1)))
        npass = np.atleast_1d(res["niter"]).astype(np.int64) + np.atleast_2d(res["passes"]).sum(axis=1)
        return int(np.sum(npass * per_pass * self.nrays + per_pass) * ndep)

    def populations(self, res):
        """``res`` of one column -> tuple of (ID, n [Nlevel, ndep], nstar [Nlevel, ndep]) per ACTIVE atom: what
        pyrh.compute1d hands back with get_populations (pyrh_compute1dray.h:4-10, pyrh.pyx:654-673)."""
        out, l0 = [], 0
        for ID, nlv in zip(self.IDs, self.plan["atom_nlevel"]):
            out.append((ID, res["n"][..., l0:l0 + nlv, :], res["nstar"][..., l0:l0 + nlv, :]))
            l0 += int(nlv)
        return tuple(out)

    def debug(self, which, shape):
        out = np.zeros(shape)
        self._lib.check(self.ctx.lib.rhb200_nlte_front_debug(self.ctx.h, int(which), out.ctypes.data_as(self._lib.dp), out.size))
        return out

    def close(self):
        self.ctx.close()

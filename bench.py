#!/usr/bin/env python
"""bench.py -- formal-solution ray-points/s of the LTE polarised hot path on B200.

Workload (BASELINE.json configs[1]): a batch of 16 384 perturbed FAL-C columns (70 depths,
B field), LTE FULL_STOKES DELO-Bezier3 synthesis of the Hinode Fe I 6301/6302 window
(301 wavelengths, 2 Kurucz lines, mu = 1), column-sharded: every rank gets its own 16 384
columns ("weak" scaling, no data-path collective).

One step = one pass of the hot path (prep + line opacity + DELO-Bezier3) over the rank's batch.
  value : ray-points/s (ncol x nlambda x ndep useful up-ray depth steps per step, summed over
          ranks) with inputs resident in HBM, timed with CUDA events on the launching stream.
  e2e   : the same workload through rhb200_compute1d_batch() (= pyrh.compute1d per column) with pinned
          HOST buffers: the nine pyrh rows per column go in, the spectra come back, copies inside the
          timed region; everything rhf1d() derives per column runs on the device.
  --impl reference : the unmodified reference rhf1d() (oracle/_ref) on the host cores
          (fork farm over all cores: the only parallel mode the reference supports).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "formal-solution ray-points/s"
NCOL_DEFAULT, NDEP, NLAMBDA = 16384, 70, 301

# algorithmic FP64 work per ray-point (SURVEY.md 8(d) counting rule)
FLOP_DELO = 400.0                              # structure-exploiting DELO-Bezier3 step (reference formulation: 1170)
FLOP_DELO_REFERENCE_FORMULATION = 1170.0
FLOP_OPACITY = 16 * 32.8 + 2 * 45.0 + 30.0     # 16 Humlicek evals (measured region mix) + 2 line preambles + combine
BYTES_DELO = 64.0                              # chi, K'(3), S(4) read per ray-point


# ------------------------------------------------------------------ synthetic inputs
def smooth_noise(rng, n, ndep, sigma_pts=7.0):
    x = rng.standard_normal((n, ndep))
    d = (np.arange(ndep)[:, None] - np.arange(ndep)[None, :]) / sigma_pts
    w = np.exp(-0.5 * d * d)
    w /= w.sum(axis=1, keepdims=True)
    y = x @ w.T
    return y / np.sqrt((w * w).sum(axis=1))[None, :]


def synth_inputs(ncol, rank=0, pinned=True):
    """Per-column inputs of the hot path.  The three golden 70-depth columns (whose background
    opacities, proton densities and heights were produced by the reference's own host code)
    are tiled and perturbed: T, ne, nH and the background by smooth +-1-2 % factors, and every
    column gets its own B (0-2500 G), inclination, azimuth, v_los (sigma 1.5 km/s) and v_mic."""
    from pyrh_b200 import api
    rng = np.random.default_rng(20261017 + 1000 * rank)
    gs = [dict(np.load(ROOT / "tests" / "golden" / f"synth70_c{c}.npz")) for c in range(3)]
    k = gs[0]["lam_keep"]
    lam = gs[0]["lam_spect"][k]
    A = api.AT
    base_at = np.zeros((3, len(A), NDEP))
    for j, g in enumerate(gs):
        for f, i in A.items():
            base_at[j, i] = g["col_" + f]
    base_chi = np.stack([g["chi_ai"][k] for g in gs])
    base_eta = np.stack([g["eta_ai"][k] for g in gs])
    alloc = api.pinned_empty if pinned else (lambda s: np.empty(s))
    at = alloc((ncol, len(A), NDEP))
    chi = alloc((ncol, NLAMBDA, NDEP))
    eta = alloc((ncol, NLAMBDA, NDEP))
    sel = np.arange(ncol) % 3
    blk = 1024
    for c0 in range(0, ncol, blk):
        s = slice(c0, min(ncol, c0 + blk))
        n = s.stop - s.start
        j = sel[s]
        a = base_at[j].copy()
        fT = 1.0 + 0.01 * smooth_noise(rng, n, NDEP)
        a[:, A["T"]] *= fT
        a[:, A["ne"]] /= fT
        a[:, A["nHtot"]] /= fT
        a[:, A["np"]] /= fT
        a[:, A["vel"]] = 1.5e3 * smooth_noise(rng, n, NDEP)
        a[:, A["vturb"]] = np.clip(1.0 + 0.5 * smooth_noise(rng, n, NDEP), 0.2, 3.0) * 1e3
        a[:, A["B"]] = np.maximum(rng.uniform(0, 2500, (n, 1)) * (1 + 0.2 * smooth_noise(rng, n, NDEP)), 0) / 1e4
        gam = rng.uniform(0, np.pi, (n, 1)) + 0.1 * smooth_noise(rng, n, NDEP)
        azi = rng.uniform(0, np.pi, (n, 1)) + 0.1 * smooth_noise(rng, n, NDEP)
        a[:, A["cos_gamma"]] = np.cos(gam)
        a[:, A["cos_2chi"]] = np.cos(2 * azi)
        a[:, A["sin_2chi"]] = np.sin(2 * azi)
        at[s] = a
        fc = (1.0 + 0.02 * smooth_noise(rng, n, NDEP))[:, None, :]
        np.multiply(base_chi[j], fc, out=chi[s])
        np.multiply(base_eta[j], fc, out=eta[s])
    return gs[0], lam, at, chi, eta


def synth_chem(ncol, rank=0, pinned=True):
    """Per-column inputs of the path with the continuum on the device: the population factors of the chemical
    equilibrium and nHmin / nH2 / nOH / nCH of the three base columns (produced by the reference's host code),
    tiled like synth_inputs() and perturbed by smooth +-2 %."""
    from pyrh_b200 import api
    rng = np.random.default_rng(77 + 1000 * rank)
    sc = dict(np.load(ROOT / "tests" / "golden" / "synth70_chem.npz"))
    base = sc["chem"]                                   # [3, natom+4, NDEP]
    natom = base.shape[1] - 4
    alloc = api.pinned_empty if pinned else (lambda s: np.empty(s))
    chem = alloc((ncol,) + base.shape[1:])
    sel = np.arange(ncol) % 3
    blk = 4096
    for c0 in range(0, ncol, blk):
        s = slice(c0, min(ncol, c0 + blk))
        n = s.stop - s.start
        b = base[sel[s]].copy()
        b[:, natom:] *= (1.0 + 0.02 * smooth_noise(rng, n, NDEP))[:, None, :]
        chem[s] = b
    return chem, sc["abundance"]


# ------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.idx, self.p = gpu_index, None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
            out, _ = self.p.communicate()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        load = [s for s in sm if s > 0.5 * max(mx)] if sm and mx else sm
        return {"sm_mhz": statistics.median(load or sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------- reference arm
_REF = {}


def _ref_init(ndep):
    from oracle import refdriver as rd
    _REF["rd"] = rd
    _REF["base"] = np.load(ROOT / "tests" / "golden" / "falc_base.npy")
    _REF["wave"] = rd.hinode_wave(NLAMBDA)
    _REF["cwd"] = rd.make_workdir("benchmark")
    _REF["ndep"] = ndep


def _ref_step(args):
    first, n = args
    from pyrh_b200 import synthetic
    rd = _REF["rd"]
    atms = synthetic.perturbed_batch(_REF["base"], n, ndep=_REF["ndep"], first=first)
    t0 = time.perf_counter()
    for a in atms:
        rd.rhf1d(a, _REF["wave"], _REF["cwd"])
    return time.perf_counter() - t0


def _ref_spectrum(column):
    from pyrh_b200 import synthetic
    a = synthetic.perturbed_batch(_REF["base"], 1, ndep=_REF["ndep"], first=column)[0]
    o = _REF["rd"].rhf1d(a, _REF["wave"], _REF["cwd"])
    return np.array([o[k] for k in "IQUV"])


def reference_throughput(steps=5, warmup=3, cols_per_proc=2, procs=None, variant="scalar", probe_cols=()):
    """rhf1d() of the unmodified reference (oracle/_ref), one process per host core (the only
    parallel mode the reference supports).  One step = every process synthesises `cols_per_proc`
    columns of the benchmark workload; `warmup` untimed + `steps` timed steps (a step lasts as
    long as its slowest process; synthetic-atmosphere generation is outside the timed section).  Returns dict(rps, sps, procs, ncols, wall)."""
    import multiprocessing as mp
    from oracle import refdriver as rd
    if not rd.available(variant):
        raise RuntimeError("oracle/_ref is not built (make -C oracle ref where /root/reference exists)")
    procs = procs or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    with ctx.Pool(procs, initializer=_ref_init, initargs=(NDEP,)) as pool:
        col = 1000
        for _ in range(warmup):
            pool.map(_ref_step, [(col + p * cols_per_proc, cols_per_proc) for p in range(procs)], chunksize=1)
            col += procs * cols_per_proc
        wall = 0.0
        for _ in range(steps):
            t = pool.map(_ref_step, [(col + p * cols_per_proc, cols_per_proc) for p in range(procs)], chunksize=1)
            wall += max(t)       # all processes run concurrently: a step lasts as long as its slowest process
            col += procs * cols_per_proc
        spectra = pool.map(_ref_spectrum, list(probe_cols), chunksize=1) if len(probe_cols) else []
    ncols = steps * procs * cols_per_proc
    return dict(spectra=spectra, rps=ncols * NLAMBDA * NDEP / wall, sps=ncols / wall, procs=procs, ncols=ncols, wall=wall,
                cols_per_step=procs * cols_per_proc)


# ------------------------------------------------------------------ NLTE records (BASELINE configs 4 and 5)
NLTE_KW = dict(N_MAX_ITER=100, N_MAX_SCATTER=2, NG_ORDER=2, NG_DELAY=10, NG_PERIOD=3, ITER_LIMIT="1.0E-4", PRD_N_MAX_ITER=0,
               STOKES_MODE="NO_STOKES")
NLTE_CASES = {
    # configs[3]: H (6 levels) + Ca II (5 levels + continuum) ACTIVE, CRD, Ng acceleration, to convergence
    "config4": dict(kw=dict(NLTE_KW, NRAYS=3, HYDROGEN_LTE="FALSE"), active=("H_6.atom", "CaII.atom"),
                    wave=(630.25, 630.5, 21), ncol=512),
    # configs[4] sample: Ca II 8542, ~1000 wavelengths, 5 mu, columns of the synthetic cube (per GPU)
    "config5_sample": dict(kw=dict(NLTE_KW, NRAYS=5, HYDROGEN_LTE="TRUE"), active=("CaII.atom",),
                           wave=(853.5, 855.5, 600), ncol=4096),
}


def _pyrh_data_path():
    """$PYRH_PATH, else the atom / molecule DATA files staged beside the oracle build (files only: nothing is imported)."""
    if not os.environ.get("PYRH_PATH"):
        os.environ["PYRH_PATH"] = str(ROOT / "oracle" / "_ref" / "pyrh_path")
    return os.environ["PYRH_PATH"]


def _nlte_workdir(case):
    import tempfile
    from pyrh_b200 import workdir
    c = NLTE_CASES[case]
    return workdir.stage(tempfile.mkdtemp(prefix=f"rhb200_{case}_"), c["kw"], active=c["active"], extra_atoms=("CaII.atom",))


_NLTE_SAMPLE = {}      # case -> first columns of the GPU arm's batch (rank 0), for the parity record of the reference leg


def nlte_records(device, rank, ncol_scale=1.0, world=1, barrier=None, maxreduce=None):
    """NLTE through the drop-in call (NlteSession.compute = rhb200_nlte_compute1d_batch): host atmosphere rows in,
    spectra + populations out, everything between on the device.  One record per case: atmospheres/s and formal-solution
    ray-points/s of a batch of perturbed 70-depth columns (and the latency of one FAL-C atmosphere for config 4)."""
    from pyrh_b200 import nlte_host, synthetic
    _pyrh_data_path()
    base = np.load(ROOT / "tests" / "golden" / "falc_base.npy")
    out = {}
    for case, c in NLTE_CASES.items():
        ncol = max(8, int(c["ncol"] * ncol_scale))
        # under N ranks the barrier and the max-reduction are collectives: every rank reaches both whatever happens to it
        s, res, err, dt = None, None, None, float("inf")
        try:
            s = nlte_host.NlteSession(_nlte_workdir(case), np.linspace(*c["wave"]), device)
            atm = synthetic.perturbed_batch(base, ncol, ndep=NDEP, first=10000 + rank * ncol)
            s.compute(atm[:min(64, ncol)])                              # warm-up: allocations, first launches
            s.ctx.synchronize()
        except Exception as e:          # noqa: BLE001
            err = e
        if barrier:
            barrier()
        times = []
        if err is None:
            try:
                for _ in range(2):      # two timed calls of the same batch, the faster one is reported (both are kept):
                    t0 = time.perf_counter()    # a call has one host read-back per MALI iteration and is sensitive to host noise
                    res = s.compute(atm)
                    s.ctx.synchronize()
                    times.append(time.perf_counter() - t0)
                dt = min(times)
            except Exception as e:      # noqa: BLE001
                err = e
        if maxreduce:                                                   # every rank solves its own ncol columns (weak scaling)
            dt = maxreduce(dt)
        if err is not None or not np.isfinite(dt):
            out[case] = {"unavailable": f"{type(err).__name__}: {err}" if err is not None else "another rank failed"}
            if s is not None:
                s.close()
            continue
        if rank == 0:                   # the columns the reference leg recomputes (column ids 10000 ...)
            _NLTE_SAMPLE[case] = {"I": np.array(res["I"][:64]), "n": np.array(res["n"][:64]), "niter": np.array(res["niter"][:64])}
        finite = np.isfinite(res["I"]).all(axis=tuple(range(1, res["I"].ndim))) & np.isfinite(res["n"]).all(axis=tuple(range(1, res["n"].ndim)))
        conv = (res["niter"] < int(c["kw"]["N_MAX_ITER"])) & finite
        rec = {"workload": f"{ncol} perturbed FAL-C columns x {NDEP} depths, {len(s.lam)} wavelengths, NRAYS {s.nrays}, "
                           f"ACTIVE {'+'.join(a.split('.')[0] for a in c['active'])}, CRD, Ng 2/10/3, ITER_LIMIT 1e-4",
               "ncol": ncol, "nspect": int(len(s.lam)), "nrays": s.nrays, "seconds": dt, "seconds_each_call": times,
               "atmospheres_per_s": ncol / dt,
               "n_gpus": world, "atmospheres_per_s_all_gpus": world * ncol / dt,        # columns sharded over the ranks, no collective
               "ray_points_per_s": s.ray_points(res, NDEP) / dt, "ray_points": s.ray_points(res, NDEP),
               "iterations_median": float(np.median(res["niter"])), "iterations_max": int(res["niter"].max()),
               "converged_columns": int(conv.sum()), "all_finite": bool(finite.all()),
               # the MALI/Ng iteration of the reference diverges on these perturbed columns (negative, then NaN
               # populations; or exit() from LUdecomp, "Singular matrix"): with the reference's summation order
               # (RHB200_NLTE_EXACT=1) the same columns fail here, bit for bit; see cpu_baseline.columns_the_reference_*
               "nonfinite_columns": int((~finite).sum()),
               "h2d_bytes": int(atm.nbytes), "d2h_bytes": int(res["I"].nbytes + res["n"].nbytes + res["nstar"].nbytes)}
        if case == "config4":                                           # the single FAL-C atmosphere configs[3] names
            one = base.copy()
            s.compute(one)
            t0 = time.perf_counter()
            for _ in range(3):
                r1 = s.compute(one)
            rec["single_falc_atmosphere_ms"] = 1e3 * (time.perf_counter() - t0) / 3
            rec["single_falc_iterations"] = int(r1["niter"])
        # per-kernel-family split of one smaller batch (CUDA events around every launch: serialising, separate pass)
        s.ctx.timing(True)
        s.compute(atm[:min(256, ncol)])
        rec["kernel_ms_per_256_columns"] = {n: ms for n, (ms, cnt) in s.ctx.timing_get().items() if cnt}
        s.ctx.timing(False)
        s.close()
        out[case] = rec
    return out


def nlte_ref_worker(case, column):
    """One column of an NLTE case through the unmodified reference, in a process of its own (bench.py --nlte-ref-worker):
    prints the seconds rhf1d() took.  The reference exit()s on columns whose statistical equilibrium turns singular."""
    from oracle import refdriver as rd
    from pyrh_b200 import synthetic
    _pyrh_data_path()
    a = synthetic.perturbed_batch(np.load(ROOT / "tests" / "golden" / "falc_base.npy"), 1, ndep=NDEP, first=column)[0]
    cwd = _nlte_workdir(case)
    t0 = time.perf_counter()
    o = rd.rhf1d(a, np.linspace(*NLTE_CASES[case]["wave"]), cwd, get_populations=True)
    dt = time.perf_counter() - t0
    ok = bool(np.isfinite(o["I"]).all() and all(np.isfinite(v["n"]).all() for v in o.get("pops", {}).values()))
    dump = os.environ.get("RHB200_NLTE_REF_DUMP")
    if dump:                            # spectrum + populations (atoms in the reference's order) for the parity record
        np.savez(Path(dump) / f"{case}_{column}.npz", I=o["I"], n=np.concatenate([v["n"] for v in o["pops"].values()]))
    print(f"NLTE_REF_FINITE {int(ok)}", flush=True)          # the reference also RETURNS NaN populations on some columns
    print(f"NLTE_REF_SECONDS {dt:.6f}", flush=True)


def nlte_reference_baseline(procs=None, limit_s=120.0):
    """The unmodified reference's rhf1d() on the same NLTE workloads: one process per host core, one column each, all
    started together (bounded sample).  Every column runs in its OWN process with a time limit: the reference calls
    exit() from LUdecomp ("Singular matrix") on some perturbed columns and would take a worker pool down with it."""
    import tempfile
    procs = procs or os.cpu_count() or 1
    out = {}
    dump = tempfile.mkdtemp(prefix="rhb200_nlte_ref_")
    env = dict(os.environ, RHB200_NLTE_REF_DUMP=dump)
    for case in NLTE_CASES:
        ps = [subprocess.Popen([sys.executable, str(ROOT / "bench.py"), "--nlte-ref-worker", case, str(10000 + p)],
                               stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, env=env) for p in range(procs)]
        t, nan_cols, t_end = [], 0, time.perf_counter() + limit_s
        for q in ps:
            try:
                o, _ = q.communicate(timeout=max(1.0, t_end - time.perf_counter()))
            except subprocess.TimeoutExpired:
                q.kill()
                q.communicate()
                continue
            t += [float(ln.split()[1]) for ln in o.splitlines() if ln.startswith("NLTE_REF_SECONDS")]
            nan_cols += sum(1 for ln in o.splitlines() if ln.startswith("NLTE_REF_FINITE 0"))
        if not t:
            out[case] = {"atmospheres_per_s": None, "cores": procs, "kind": "reference", "sample": "no column finished"}
            continue
        out[case] = {"atmospheres_per_s": len(t) / max(t), "cores": procs, "kind": "reference",
                     "seconds_per_atmosphere_per_core": float(np.mean(t)), "columns_finished": len(t),
                     "columns_the_reference_aborted_on": procs - len(t),
                     "columns_the_reference_returned_nan_for": nan_cols,
                     "sample": f"{procs} perturbed columns started together, one rhf1d(get_populations) process per core; "
                               f"{len(t)} finished, slowest {max(t):.1f} s"}
        # parity of the GPU arm on the very same columns (default rate accumulation: fixed-partition sums, so the bar is
        # north_star's 1e-6 on populations, not bits; RHB200_NLTE_EXACT=1 reproduces the reference's order bit for bit)
        g = _NLTE_SAMPLE.get(case)
        if g is not None:
            errn, errI, ncmp, both_nan, one_nan, per_col = 0.0, 0.0, 0, 0, 0, []
            for p in range(min(procs, len(g["I"]))):
                f = Path(dump) / f"{case}_{10000 + p}.npz"
                if not f.exists():
                    continue
                r = np.load(f)
                fin_r = bool(np.isfinite(r["I"]).all() and np.isfinite(r["n"]).all())
                fin_g = bool(np.isfinite(g["I"][p]).all() and np.isfinite(g["n"][p]).all())
                if not fin_r or not fin_g:
                    both_nan += int(not fin_r and not fin_g)
                    one_nan += int(fin_r != fin_g)
                    continue
                ncmp += 1
                en = float(np.max(np.abs(g["n"][p].reshape(r["n"].shape) / r["n"] - 1.0)))
                eI = float(np.max(np.abs(g["I"][p] / r["I"] - 1.0)))
                per_col.append({"column": 10000 + p, "iterations": int(g["niter"][p]), "max_rel_err_populations": en, "max_rel_err_I": eI})
                errn, errI = max(errn, en), max(errI, eI)
            out[case]["parity_vs_reference"] = {"columns_compared": ncmp, "max_rel_err_populations": errn, "max_rel_err_I": errI,
                                                "columns_nan_in_both": both_nan, "columns_nan_in_one_only": one_nan,
                                                "tolerance": 1.0e-6, "within_tolerance": bool(ncmp > 0 and errn <= 1.0e-6 and errI <= 1.0e-6),
                                                "columns_within_tolerance": sum(1 for c in per_col if c["max_rel_err_populations"] <= 1.0e-6),
                                                "per_column": per_col,
                                                "note": "columns that converge in the usual ~30 iterations agree to ~1e-9; a column the MALI/Ng "
                                                        "iteration struggles with (50+ iterations) amplifies the summation-order difference of the "
                                                        "default fixed-partition rate sums; RHB200_NLTE_EXACT=1 (reference order) is bit-identical"}
    return out


# ------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ncol", type=int, default=NCOL_DEFAULT)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-cols-per-proc", type=int, default=3)
    ap.add_argument("--nlte-ref-worker", nargs=2, metavar=("CASE", "COLUMN"), help=argparse.SUPPRESS)
    ap.add_argument("--no-nlte", action="store_true", help="skip the NLTE (configs 4 / 5) records")
    ap.add_argument("--nlte-scale", type=float, default=1.0, help="scale the NLTE batch sizes (profiling runs)")
    args = ap.parse_args()
    if args.nlte_ref_worker:
        nlte_ref_worker(args.nlte_ref_worker[0], int(args.nlte_ref_worker[1]))
        return

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = {"workload": f"configs[1]: {args.ncol} perturbed FAL-C columns/GPU x {NDEP} depths x {NLAMBDA} "
                          "wavelengths (Hinode Fe I 6301/6302, 2 Kurucz lines, 16 Zeeman components), LTE "
                          "FULL_STOKES DELO-Bezier3, mu=1",
              "ncol_per_gpu": args.ncol, "ndep": NDEP, "nlambda": NLAMBDA, "parallelism": f"columns x{world}",
              "l2": "inputs (5.5 GB) and ray-point workspace (>5 GB) far larger than the 126 MB L2"}

    if args.impl == "reference":
        if rank != 0:
            return
        r = reference_throughput(args.steps, args.warmup, args.ref_cols_per_proc)
        rps, sps, procs, ncols, wall = r["rps"], r["sps"], r["procs"], r["ncols"], r["wall"]
        line = {"impl": "reference", "metric": METRIC, "value": rps, "unit": "ray-points/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config, "spectra_per_s": sps,
                "cpu_baseline": {"value": rps, "unit": "ray-points/s", "cores": procs, "kind": "reference",
                                 "sample": f"{ncols} perturbed FAL-C columns in {args.steps} steps of {r['cols_per_step']} "
                                           f"({args.ref_cols_per_proc}/process x {procs} processes), {wall:.1f} s wall; full "
                                           "rhf1d() per column incl. input parsing + continuum (the reference cannot "
                                           "run the hot path alone)"},
                "e2e": {"value": rps, "unit": "ray-points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from pyrh_b200 import api
    from pyrh_b200.linelist import LineTable
    ctx = api.Context(local_rank)
    g0, lam, at, chi, eta = synth_inputs(args.ncol, rank)
    ctx.set_lines(LineTable.from_npz(g0))
    ctx.set_wavelengths(lam)
    ncol = args.ncol
    stokes = api.pinned_empty((ncol, 4, NLAMBDA))
    units = ncol * NLAMBDA * NDEP

    def barrier():
        ctx.synchronize()
        if dist is not None:
            dist.barrier()
        ctx.synchronize()

    def maxreduce(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm
    d_at, d_chi, d_eta, d_st = (ctx.dev_alloc(x.nbytes) for x in (at, chi, eta, stokes))
    ctx.h2d(d_at, at); ctx.h2d(d_chi, chi); ctx.h2d(d_eta, eta)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        ctx.lte_stokes_batch_dev(ncol, NDEP, d_at, d_chi, d_eta, d_st)
    ctx.timing(False)
    barrier()
    ctx.timer_begin()
    for _ in range(args.steps):
        ctx.lte_stokes_batch_dev(ncol, NDEP, d_at, d_chi, d_eta, d_st)
    ms_dev = ctx.timer_end()
    launches = sum(v[1] for v in ctx.timing_get().values())
    barrier()
    ms_dev = maxreduce(ms_dev)
    dev_out = np.empty((ncol, 4, NLAMBDA))
    ctx.d2h(dev_out, d_st)

    # ---- per-kernel durations (CUDA events around every launch; serialising, so separate pass)
    ctx.timing(True)
    for _ in range(max(1, min(args.steps, 3))):
        ctx.lte_stokes_batch_dev(ncol, NDEP, d_at, d_chi, d_eta, d_st)
    kt = {n: (ms / max(cnt, 1), cnt) for n, (ms, cnt) in ctx.timing_get().items() if cnt}
    ctx.timing(False)
    nlaunch_pass = {n: v[1] // max(1, min(args.steps, 3)) for n, v in kt.items()}
    for p in (d_at, d_chi, d_eta, d_st):
        ctx.dev_free(p)

    # ---- end-to-end arm A: host (pinned) buffers through the C ABI, background opacities supplied by the host
    for _ in range(max(1, args.warmup - 1)):
        ctx.lte_stokes_batch(at, chi, eta, out=stokes)
    barrier()
    t0 = time.perf_counter()
    ctx.timer_begin()
    for _ in range(args.steps):
        ctx.lte_stokes_batch(at, chi, eta, out=stokes)
    ms_e2e_dev = ctx.timer_end()
    ms_e2e_bg = max(ms_e2e_dev, 1e3 * (time.perf_counter() - t0))
    barrier()
    ms_e2e_bg = maxreduce(ms_e2e_bg)
    same = bool(np.array_equal(dev_out, stokes))

    # ---- end-to-end arm B (the headline): the whole LTE column on the device -- LTE populations, chemical
    #      equilibrium, continuum, line opacity, DELO-Bezier3; the host sends the atmosphere rows only
    from pyrh_b200 import continuum
    full = dict(np.load(ROOT / "tests" / "golden" / "falc_full.npz"))
    abundance = np.load(ROOT / "tests" / "golden" / "synth70_chem.npz")["abundance"]
    model = continuum.ContinuumModel(full)
    #      These are the columns of pyrh_b200.synthetic (SURVEY 8(d) recipe) in pyrh's own units -- exactly what
    #      pyrh.compute1d takes and what the reference arm feeds rhf1d() -- through rhb200_compute1d_batch.
    from pyrh_b200 import synthetic
    scales = dict(np.load(ROOT / "tests" / "golden" / "pyrh_scales.npz"))
    wght_per_H = float(scales["tau_abund_sums"][0])
    lam_spect = g0["lam_spect"]                       # spectrum.lambda: the 301 user wavelengths + lambda_ref
    ctx.set_wavelengths(lam_spect)
    ctx.set_continuum(model, abundance)
    ctx.set_chemistry(full["ce_nuclei"][:, 1].astype(np.int32), full["ce_mol"])
    base = np.load(ROOT / "tests" / "golden" / "falc_base.npy")
    pyrh_atm = api.pinned_empty((ncol, 9, NDEP))
    pyrh_atm[:] = synthetic.perturbed_batch(base, ncol, ndep=NDEP, first=rank * ncol)
    stokes_b = api.pinned_empty((ncol, 4, NLAMBDA + 1))
    ctx.timing(False)
    run_b = lambda: ctx.compute1d_batch(pyrh_atm, wght_per_H=wght_per_H, out=stokes_b, keep_lambda_ref=True)
    for _ in range(max(1, args.warmup - 1)):
        run_b()
    launches_b0 = sum(v[1] for v in ctx.timing_get().values())
    barrier()
    t0 = time.perf_counter()
    ctx.timer_begin()
    for _ in range(args.steps):
        run_b()
    ms_e2e_dev = ctx.timer_end()
    ms_e2e = max(ms_e2e_dev, 1e3 * (time.perf_counter() - t0))
    barrier()
    ms_e2e = maxreduce(ms_e2e)
    launches_b = sum(v[1] for v in ctx.timing_get().values()) - launches_b0
    # ---- strong scaling of the same call: the 16 384 columns of configs[1] in TOTAL, column-sharded over the ranks
    #      (each rank runs its contiguous block; no collective on the data path)
    lo, cnt = rank * ncol // world, (rank + 1) * ncol // world - rank * ncol // world
    run_s = lambda: ctx.compute1d_batch(pyrh_atm[lo:lo + cnt], wght_per_H=wght_per_H, out=stokes_b[lo:lo + cnt], keep_lambda_ref=True)
    run_s()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run_s()
    ctx.synchronize()
    ms_strong = 1e3 * (time.perf_counter() - t0)
    barrier()
    ms_strong = maxreduce(ms_strong)
    # ---- the same from ONE process driving every visible GPU (rhb200_compute1d_batch_multi): only when not under torchrun
    multi = None
    if world == 1 and args.gpus > 1:
        import ctypes as C
        from pyrh_b200 import _lib
        ctxs = [ctx]
        for d in range(1, args.gpus):
            cd = api.Context(d)
            cd.set_lines(LineTable.from_npz(g0)); cd.set_wavelengths(lam_spect); cd.set_continuum(model, abundance)
            cd.set_chemistry(full["ce_nuclei"][:, 1].astype(np.int32), full["ce_mol"])
            ctxs.append(cd)
        handles = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
        iref = int(np.flatnonzero(lam_spect == 500.0)[0])
        run_m = lambda: _lib.check(ctx.lib.rhb200_compute1d_batch_multi(
            len(ctxs), handles, ncol, NDEP, 9, 1.0, 0, C.c_void_p(pyrh_atm.ctypes.data), iref, wght_per_H, 0.0,
            _lib.BC_ZERO, _lib.BC_THERMALIZED, C.c_void_p(stokes_b.ctypes.data), None))
        ref_b = stokes_b.copy()
        run_m(); run_m()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            run_m()
        ms_multi = 1e3 * (time.perf_counter() - t0)
        multi = {"devices": args.gpus, "ncol_total": ncol, "ms_per_step": ms_multi / args.steps,
                 "value": units * args.steps / (ms_multi * 1e-3), "unit": "ray-points/s",
                 "speedup_vs_one_gpu": (ms_e2e / ms_multi), "bitwise_equal_to_one_gpu": bool(np.array_equal(ref_b, stokes_b)),
                 "call": "rhb200_compute1d_batch_multi: one process, one host thread per device, contiguous column blocks"}
        for cd in ctxs[1:]:
            cd.close()
    ctx.timing(True)
    run_b()
    kt_b = {n: (ms / max(cnt, 1), cnt) for n, (ms, cnt) in ctx.timing_get().items() if cnt}
    ctx.timing(False)
    clocks = sampler.stop()          # sampled from the first warm-up step to the end of the e2e regions
    finite_b = bool(np.isfinite(stokes_b).all())

    fma_tf, nofma_tf = ctx.fp64_peak()
    nlte = None
    if not args.no_nlte:
        try:
            nlte = nlte_records(local_rank, rank, args.nlte_scale, world, barrier, maxreduce)
        except Exception as e:          # noqa: BLE001  -- the headline must survive a missing data directory
            nlte = {"unavailable": f"{type(e).__name__}: {e}"}

    if rank != 0:
        ctx.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    value = world * units * args.steps / (ms_dev * 1e-3)
    e2e_val = world * units * args.steps / (ms_e2e * 1e-3)
    # dominant kernel
    dom = max(kt, key=lambda n: kt[n][0] * nlaunch_pass[n])
    dom_ms = kt[dom][0]
    cols_per_launch = ncol / max(1, nlaunch_pass[dom])
    pts_per_launch = cols_per_launch * NLAMBDA * NDEP
    flop_pt = {"delo": FLOP_DELO, "opacity": FLOP_OPACITY}.get(dom, 0.0)
    achieved_tf = pts_per_launch * flop_pt / (dom_ms * 1e-3) / 1e12
    traffic = None            # dram__bytes_read.sum + dram__bytes_write.sum of the committed ncu capture
    try:
        import csv
        for row in csv.DictReader(open(ROOT / "profiles" / "r1_lte_kernels.csv")):
            if {"delo": "delo_raypts", "opacity": "opacity_fused"}.get(dom, "@") in row["kernel"]:
                gb = float(row["dram__bytes_read.sum"]) + float(row["dram__bytes_write.sum"])
                traffic = gb * 1e9 * (cols_per_launch / 4096.0)      # capture was taken at 4096 columns/launch
    except (OSError, KeyError, ValueError):
        pass
    roofline = {"kernel": {"delo": "delo_raypts_kernel", "opacity": "opacity_fused_kernel",
                           "prep": "prep_kernel"}.get(dom, dom),
                "bound": "fp64", "achieved": achieved_tf, "peak": fma_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / fma_tf if fma_tf else None, "traffic": traffic,
                # the arithmetic contract (-fmad=false: every product and sum rounds like the reference's x86-64 build)
                # caps this code at one flop per FP64 instruction; the headline fraction stays the one against the DFMA peak
                "frac_of_no_fma_ceiling": achieved_tf / nofma_tf if nofma_tf else None,
                "algorithmic_bytes": pts_per_launch * (BYTES_DELO if dom == "delo" else 80.0),
                "peak_source": "rhb200_fp64_peak(): DFMA micro-benchmark run in this process (2 flop/FMA); "
                               f"non-FMA FP64 issue rate {nofma_tf:.1f} Tinst/s",
                "algorithmic_flop_per_raypoint": flop_pt, "ms_per_launch": dom_ms,
                "delo": {"ms_per_launch": kt.get("delo", (0, 0))[0],
                         "achieved_tflops": (pts_per_launch * FLOP_DELO / (kt["delo"][0] * 1e-3) / 1e12) if "delo" in kt else None,
                         "algorithmic_flop_per_raypoint": FLOP_DELO,
                         "reference_formulation_flop_per_raypoint": FLOP_DELO_REFERENCE_FORMULATION},
                "hbm": {"achieved": pts_per_launch * (BYTES_DELO if dom == "delo" else 80.0) / (dom_ms * 1e-3) / 1e9,
                        "peak": hbm_peak, "unit": "GB/s",
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s"},
                "kernels_ms_per_launch": {n: v[0] for n, v in kt.items()},
                "kernel_launches_per_step": nlaunch_pass}
    line = {"metric": METRIC, "value": value, "unit": "ray-points/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "spectra_per_s": world * ncol * args.steps / (ms_dev * 1e-3),
            "e2e": {"value": e2e_val, "unit": "ray-points/s",
                    "call": "rhb200_compute1d_batch (= pyrh.compute1d per column): the nine pyrh rows in pyrh units in, "
                            "Stokes spectra out; unit conversion, Bproject, LTE populations, chemical equilibrium, "
                            "proton density, background continuum, line opacity, tau500->height (convertScales) and "
                            "DELO-Bezier3 all on the device; ray-points counted on the 301 user wavelengths only "
                            "(the lambda_ref ray is computed too, like the reference, and not counted)",
                    "h2d_bytes_per_step": int(pyrh_atm.nbytes),
                    "d2h_bytes_per_step": int(stokes_b.nbytes), "ms_per_step": ms_e2e / args.steps,
                    "spectra_per_s": world * ncol * args.steps / (ms_e2e * 1e-3),
                    "gpu_launches": int(launches_b), "all_finite": finite_b,
                    "kernels_ms_per_launch": {n: v[0] for n, v in kt_b.items()}},
            "e2e_background_from_host": {"value": world * units * args.steps / (ms_e2e_bg * 1e-3), "unit": "ray-points/s",
                    "call": "rhb200_lte_stokes_batch: chi_ai / eta_ai computed by the RH host and copied in",
                    "h2d_bytes_per_step": int(at.nbytes + chi.nbytes + eta.nbytes),
                    "d2h_bytes_per_step": int(stokes.nbytes), "ms_per_step": ms_e2e_bg / args.steps,
                    "bitwise_equal_to_device_resident_run": same},
            "strong_scaling": {"ncol_total": ncol, "n_gpus": world, "ms_per_step": ms_strong / args.steps,
                               "value": units * args.steps / (ms_strong * 1e-3), "unit": "ray-points/s",
                               "note": "the SAME 16 384-column batch split over the ranks (e2e call, host buffers); "
                                       "efficiency = value / (n_gpus x the n_gpus = 1 value)"},
            "single_process_multi_gpu": multi,
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "nlte": nlte}
    if world == 1 and not args.no_cpu_baseline:
        try:
            probe_cols = [c for c in (0, 1000, 4097, ncol - 1) if c < ncol]
            r = reference_throughput(5, 2, args.ref_cols_per_proc, probe_cols=probe_cols)
            rps, sps, procs, ncols, wall = r["rps"], r["sps"], r["procs"], r["ncols"], r["wall"]
            # the checker, not the product: the same columns went through rhb200_compute1d_batch above
            keep = lam_spect != 500.0
            line["e2e"]["parity_vs_reference"] = {
                "columns": probe_cols,
                "bitwise_equal": [bool(np.array_equal(stokes_b[c][:, keep], sp)) for c, sp in zip(probe_cols, r["spectra"])],
                "max_rel_err_I": float(max(np.max(np.abs(stokes_b[c][0, keep] / sp[0] - 1)) for c, sp in zip(probe_cols, r["spectra"])))}
            line["cpu_baseline"] = {"value": rps, "unit": "ray-points/s", "cores": procs, "kind": "reference",
                                    "spectra_per_s": sps,
                                    "sample": f"{ncols} perturbed FAL-C columns (70 depths, 301 wavelengths), one rhf1d() "
                                              f"process per core x {procs} cores, {wall:.1f} s; whole call (parsing + "
                                              "continuum + line opacity + formal solution)"}
        except Exception as e:          # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "ray-points/s", "cores": 0, "kind": "reference",
                                    "sample": f"unavailable: {e}"}
        if isinstance(nlte, dict) and "config4" in nlte:
            try:
                for case, b in nlte_reference_baseline().items():
                    nlte[case]["cpu_baseline"] = b
                    if b["atmospheres_per_s"]:
                        nlte[case]["speedup_vs_reference_all_cores"] = nlte[case]["atmospheres_per_s"] / b["atmospheres_per_s"]
            except Exception as e:          # noqa: BLE001
                nlte["cpu_baseline_unavailable"] = f"{type(e).__name__}: {e}"
    print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Stage a pyrh working directory: the four text inputs ``rhf1d()`` reads from ``cwd`` (``keyword.input``,
``atoms.input``, ``molecules.input``, ``kurucz.input`` + the Kurucz line files it names) in the reference's own
formats (rh/readinput.c:43-215, rh/readatom.c:797-903, rh/readmolecule.c:933-1010, rh/kurucz.c:157-184).  The model
atoms and molecules themselves are looked up under ``$PYRH_PATH/rh/Atoms`` / ``rh/Molecules`` like in the reference.

Used by bench.py and the examples to set up the BASELINE workloads without copying a directory around; a user's own
directory works the same way.
"""
from __future__ import annotations

from pathlib import Path

# Fe I 630.15 / 630.25 nm in Kurucz's fixed-column format (the two records of the reference's benchmark/fe6300 and
# tests/fe6300 line list; public Kurucz atomic data, no orbital-number columns)
FE6300 = (
    "  630.1500 -0.71  26.00   45333.875  2.0 4s6D5s e5D   29469.024  2.0 5Dsp3P z5P  8.08 -5.42 -7.54DRLP 0 0  0 0.000  0 0.000                     1503 1835     0 0 1\n"
    "  630.2493 -0.969 26.00   45595.086  0.0 4s6D5s e5D   29732.736  1.0 5Dsp3P z5P  8.08 -5.40 -7.54K17  0 0  0 0.000  0 0.000                        0 2487     0 0 1\n")

ATOMS_STANDARD = ("H_6.atom", "He.atom", "C.atom", "N.atom", "O.atom", "S.atom", "Fe.atom", "Si.atom", "Al.atom", "Na.atom",
                  "Mg.atom")
MOLECULES_STANDARD = ("H2.molecule", "H2+.molecule", "C2.molecule", "N2.molecule", "O2.molecule", "CH.molecule",
                      "CO.molecule", "CN.molecule", "NH.molecule", "NO.molecule", "OH.molecule", "H2O.molecule")
KEYWORDS_STANDARD = {
    "NRAYS": 1, "ATOMS_FILE": "atoms.input", "MOLECULES_FILE": "molecules.input", "N_MAX_SCATTER": 0, "I_SUM": -1,
    "N_MAX_ITER": 1, "ITER_LIMIT": "1.0E-2", "NG_ORDER": 0, "NG_DELAY": 10, "NG_PERIOD": 3, "PRD_N_MAX_ITER": 0,
    "PRD_ITER_LIMIT": "1.0E-2", "J_FILE": "J.dat", "STARTING_J": "NEW_J", "BACKGROUND_FILE": "background.dat",
    "OLD_BACKGROUND": "FALSE", "KURUCZ_DATA": "kurucz.input", "SOLVE_NE": "NONE", "RLK_SCATTER": "FALSE",
    "HYDROGEN_LTE": "TRUE", "VMICRO_CHAR": 5.0, "VMACRO_TRESH": 0, "S_INTERPOLATION": "S_BEZIER3",
    "S_INTERPOLATION_STOKES": "DELO_BEZIER3", "VACUUM_TO_AIR": "FALSE", "STOKES_MODE": "NO_STOKES",
    "MAGNETO_OPTICAL": "FALSE", "LIMIT_MEMORY": "FALSE", "PRINT_CPU": "FALSE"}


def stage(path, keywords=None, active=(), extra_atoms=(), atoms=ATOMS_STANDARD, molecules=MOLECULES_STANDARD,
          kurucz_records=FE6300, kurucz_name="lines.kur"):
    """Write the input files into ``path`` (created if missing) and return it as a string.  ``keywords`` override
    KEYWORDS_STANDARD; atoms named in ``active`` are ACTIVE (``extra_atoms`` are appended to the standard list)."""
    p = Path(path)
    p.mkdir(parents=True, exist_ok=True)
    kw = dict(KEYWORDS_STANDARD)
    kw.update(keywords or {})
    (p / "keyword.input").write_text("".join(f"  {k} = {v}\n" for k, v in kw.items()))
    listed = list(atoms) + [a for a in extra_atoms if a not in atoms]
    rows = [f"  {a:<16s} {'ACTIVE ' if a in active else 'PASSIVE'}     LTE_POPULATIONS   pops.{a.split('.')[0]}.out\n" for a in listed]
    (p / "atoms.input").write_text(f"   {len(listed)}\n" + "".join(rows))
    (p / "molecules.input").write_text(f"{len(molecules)}\n" + "".join(f"  {m:<16s} PASSIVE    LTE_POPULATIONS\n" for m in molecules))
    (p / kurucz_name).write_text(kurucz_records)
    (p / "kurucz.input").write_text(kurucz_name + "\n")
    return str(p)

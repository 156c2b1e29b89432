"""Host-side Zeeman machinery of librhb200.so (no GPU needed) against vectors produced by the
unmodified reference routines (oracle/gen_golden_zeeman.py): RLKdeterminate, RLKZeeman (kurucz.c:832-969),
determinate, Zeeman, Lande (zeeman.c:37-281).  Integer outputs (determined flag, S/L, component count,
q and component order) and the double shifts/strengths must be identical."""
import numpy as np
import pytest

from conftest import GOLD


@pytest.fixture(scope="module")
def g():
    return dict(np.load(GOLD / "zeeman.npz"))


def test_lande_bit_exact(g):
    from pyrh_b200 import zeeman
    got = np.array([zeeman.lande(S, int(L), J) for S, L, J in g["lande_in"]])
    assert np.array_equal(got, g["lande"])


def test_rlk_determinate_and_zeeman_on_kurucz_labels(g):
    from pyrh_b200 import zeeman
    off = 0
    ndet = 0
    for i in range(len(g["rlk_det"])):
        det, Si, Li, Sj, Lj = zeeman.rlk_determinate(str(g["rlk_labeli"][i]), str(g["rlk_labelj"][i]))
        assert det == bool(g["rlk_det"][i])
        if det:
            assert [Si, Li, Sj, Lj] == list(g["rlk_SL"][i])
            ndet += 1
        if g["rlk_LS"][i] < 0:
            continue
        SL = g["rlk_SL"][i]
        q, sh, st = zeeman.rlk_zeeman(g["rlk_gi"][i], g["rlk_gj"][i], SL[0], int(SL[1]), SL[2], int(SL[3]),
                                      g["rlk_gLi"][i], g["rlk_gLj"][i], LS_Lande=bool(g["rlk_LS"][i]))
        n = int(g["rlk_ncomp"][i])
        assert len(q) == n
        assert np.array_equal(q, g["rlk_q"][off:off + n])
        assert np.array_equal(sh, g["rlk_shift"][off:off + n])
        assert np.array_equal(st, g["rlk_strength"][off:off + n], equal_nan=True)
        off += n
    assert off == len(g["rlk_q"]) and ndet >= 30
    # Fe I 6301.5 / 6302.5 of the benchmark list: 13 and 3 components
    assert {13, 3} <= set(int(x) for x in g["rlk_ncomp"])


def test_rlk_zeeman_quantum_number_sweep(g):
    from pyrh_b200 import zeeman
    off = 0
    for row in g["sweep"]:
        gi, gj, Sl, Ll, Su, Lu, gLi, gLj, LS, n = row
        q, sh, st = zeeman.rlk_zeeman(gi, gj, Sl, int(Ll), Su, int(Lu), gLi, gLj, LS_Lande=bool(LS))
        n = int(n)
        assert len(q) == n and np.array_equal(q, g["sweep_q"][off:off + n])
        assert np.array_equal(sh, g["sweep_shift"][off:off + n])
        assert np.array_equal(st, g["sweep_strength"][off:off + n], equal_nan=True)   # J=0->0: 0/0 in both
        for qq in (-1, 0, 1):                       # strengths are normalised per q
            if np.any(q == qq):
                assert np.isnan(st).any() or abs(st[q == qq].sum() - 1.0) < 1e-14
        off += n
    assert off == len(g["sweep_q"])


def test_model_atom_determinate_and_zeeman(g):
    from pyrh_b200 import zeeman
    for lab, gg, det, ref in zip(g["atom_label"], g["atom_g"], g["atom_det"], g["atom_nSLJ"]):
        d, n, S, L, J = zeeman.determinate(str(lab), gg)
        if det < 0:                                 # no parity letter / unscannable term: not determined
            assert not d
            continue
        assert d == bool(det)
        if det:
            assert [n, S, L, J] == list(ref)
    off = 0
    for i in range(len(g["zl_li"])):
        q, sh, st = zeeman.zeeman(str(g["zl_li"][i]), g["zl_gi"][i], str(g["zl_lj"][i]), g["zl_gj"][i],
                                  g["zl_geff"][i])
        n = int(g["zl_ncomp"][i])
        assert len(q) == n and np.array_equal(q, g["zl_q"][off:off + n])
        assert np.array_equal(sh, g["zl_shift"][off:off + n])
        assert np.array_equal(st, g["zl_strength"][off:off + n], equal_nan=True)
        off += n
    assert off == len(g["zl_q"])


def test_patterns_of_the_benchmark_line_table(g, golden_falc):
    """The Zeeman components the reference attached to the two Fe I lines of the benchmark run
    (fixture falc_B1kG, recorded from atmos.rlk_lines[].zm) are what the host code derives from the
    labels/J/Lande columns of benchmark/fe6300 (LS_LANDE as in benchmark/keyword.input)."""
    from pyrh_b200 import zeeman
    from pyrh_b200.linelist import RL_GI, RL_GJ, RL_ZOFF, RL_NCOMP
    f = golden_falc
    for row in f["lt_lines"]:
        o, n = int(row[RL_ZOFF]), int(row[RL_NCOMP])
        hit = [i for i in range(len(g["rlk_det"])) if g["rlk_det"][i] and g["rlk_gi"][i] == row[RL_GI]
               and g["rlk_gj"][i] == row[RL_GJ] and g["rlk_ncomp"][i] == n and "e5D" in str(g["rlk_labelj"][i])]
        assert hit
        ok = False
        for i in hit:
            SL = g["rlk_SL"][i]
            q, sh, st = zeeman.rlk_zeeman(row[RL_GI], row[RL_GJ], SL[0], int(SL[1]), SL[2], int(SL[3]),
                                          g["rlk_gLi"][i], g["rlk_gLj"][i], LS_Lande=bool(g["rlk_LS"][i]))
            ok |= (np.array_equal(q, f["lt_zq"][o:o + n]) and np.array_equal(sh, f["lt_zshift"][o:o + n])
                   and np.array_equal(st, f["lt_zstrength"][o:o + n]))
        assert ok

"""Generate the lookup tables of the table-driven double-precision libm algorithms that glibc
2.39 ships, so the device code can reproduce glibc's exp / pow / sin / cos bit for bit
(DESIGN.md "libm contract").  Everything is computed from first principles with 100-digit
decimal arithmetic following the published construction rules:

 exp (e_exp.c, S. Nagy / ARM optimized-routines), N = 128:
     2^(k/N) = H[k](1 + T[k]);  tab[2k] = asuint64(T[k]), tab[2k+1] = asuint64(H[k]) - (k<<52)/N
 pow's log (e_pow_log_data.c), N = 128, z in [OFF, 2 OFF), OFF = 0x3fe6955500000000:
     invc = j/128 (c < 1) or j/256 (c >= 1), nearest to 1/center of subinterval i
     logc = round(2^43 log c) / 2^43,  logctail = RN(log c - logc)
 sin/cos (IBM Accurate Mathematical Library, s_sin.c / sincostab.c): for x_k = k/128, k = 0..109
     {RN(sin x_k), RN(sin x_k - hi), RN(cos x_k), RN(cos x_k - hi)}

`--check` compares every generated word with the tables inside this machine's libm.so.6
(located by searching for the neighbouring constants); used once at development time and by
tests/test_math_cpu.py when that libm is glibc 2.39.

Writes pyrh_b200/csrc/rhb200_math_tables.inc
"""
import struct
import sys
from decimal import Decimal, getcontext, ROUND_HALF_EVEN
from fractions import Fraction
from pathlib import Path

getcontext().prec = 110
ROOT = Path(__file__).resolve().parent.parent
N = 128
OFF = 0x3FE6955500000000


def asuint(f: float) -> int:
    return struct.unpack("<Q", struct.pack("<d", f))[0]


def asdouble(u: int) -> float:
    return struct.unpack("<d", struct.pack("<Q", u & (2**64 - 1)))[0]


def dec(f: float) -> Decimal:
    fr = Fraction(f)
    return Decimal(fr.numerator) / Decimal(fr.denominator)


def rn(x: Decimal) -> float:
    return float(x)            # correctly rounded (round-half-even) conversion


def exp_table():
    tab = []
    for k in range(N):
        v = Decimal(2) ** (Decimal(k) / Decimal(N))
        H = rn(v)
        T = rn(v / dec(H) - 1)
        tab += [asuint(T), (asuint(H) - ((k << 52) // N)) & (2**64 - 1)]
    return tab


def dsin(x: Decimal) -> Decimal:
    term, s, n = x, x, 1
    while abs(term) > Decimal(10) ** -105:
        term = -term * x * x / ((n + 1) * (n + 2))
        s += term
        n += 2
    return s


def dcos(x: Decimal) -> Decimal:
    term, s, n = Decimal(1), Decimal(1), 0
    while abs(term) > Decimal(10) ** -105:
        term = -term * x * x / ((n + 1) * (n + 2))
        s += term
        n += 2
    return s


# Low words of IBM's published sincostab.c that are NOT the correctly rounded residual
# (they differ from RN(f(x_k) - hi) by 1-40 ulp of the low word, i.e. < 2^-100 of the value);
# kept as published so the table is word-identical to glibc's.
SINCOS_EXCEPTIONS = {
    9: "-0x1.2ab639a9f0777p-63", 41: "-0x1.921915299468cp-58", 93: "-0x1.32c5c8b81c940p-66",
    107: "0x1.e3a0d3e03b1d5p-57", 109: "-0x1.9883b57d6cdebp-58", 133: "-0x1.9b8c29dfd8ec8p-56",
    137: "-0x1.9fb0a0c93e2b5p-56", 145: "0x1.46076fe0dcff5p-56", 161: "0x1.03d5504878398p-63",
    179: "-0x1.660aec7ef636cp-58", 283: "0x1.8ff7947027a16p-58", 301: "-0x1.f190c70cbb5ffp-58",
    303: "-0x1.b83d607cd5070p-63", 319: "0x1.95e25736c0358p-60", 341: "-0x1.97653a7d2f07bp-56",
    361: "0x1.0da05738cc59ap-61", 377: "0x1.c843b4d0fb198p-58", 429: "0x1.ad1197ccd0393p-59",
}


def sincos_table():
    out = []
    for k in range(110):
        x = Decimal(k) / Decimal(128)
        s, c = dsin(x), dcos(x)
        sh, ch = rn(s), rn(c)
        out += [sh, rn(s - dec(sh)), ch, rn(c - dec(ch))]
    for i, v in SINCOS_EXCEPTIONS.items():
        out[i] = float.fromhex(v)
    return out


def powlog_table():
    out = []
    for i in range(N):
        zlo, zhi = asdouble(OFF + (i << 45)), asdouble(OFF + ((i + 1) << 45))
        # z = asdouble(iz) with the exponent of x removed: values >= 2^0 boundary wrap handled by asdouble
        center = (dec(zlo) + dec(zhi)) / 2
        den = 128 if center < 1 else 256
        j = int((Decimal(den) / center).to_integral_value(rounding=ROUND_HALF_EVEN))
        invc = j / den
        logc_exact = -(Decimal(j) / Decimal(den)).ln()
        logc = float((logc_exact * (2 ** 43)).to_integral_value(rounding=ROUND_HALF_EVEN)) / 2.0 ** 43
        logctail = rn(logc_exact - dec(logc))
        out += [invc, 0.0, logc, logctail]
    return out


def find(b: bytes, *vals) -> int:
    key = b"".join(struct.pack("<d", v) for v in vals)
    off = b.find(key)
    if off < 0:
        raise RuntimeError("constants not found in libm")
    return off


def check(exp_t, sc_t, pl_t, libm="/lib/x86_64-linux-gnu/libm.so.6"):
    b = open(libm, "rb").read()
    ok = True
    off = find(b, float.fromhex("0x1.71547652b82fep+7"), float.fromhex("0x1.8p+52")) + 0xB0
    ref = list(struct.unpack_from("<256Q", b, off))
    n = sum(a == r for a, r in zip(exp_t, ref))
    print(f"exp table      : {n}/256 words identical to libm")
    ok &= n == 256
    off = find(b, float.fromhex("0x1.62e42fefa3800p-1"), float.fromhex("0x1.ef35793c76730p-45"), -0.5) + 0x48
    ref = list(struct.unpack_from("<512d", b, off))
    bad = [i for i in range(512) if asuint(ref[i]) != asuint(pl_t[i])]
    print(f"pow log table  : {512 - len(bad)}/512 words identical to libm", bad[:8])
    ok &= not bad
    # sincostab: starts with sin(0)=0,0,cos(0)=1,0 then sin(1/128)...
    off = find(b, 0.0, 0.0, 1.0, 0.0, sc_t[4])
    ref = list(struct.unpack_from("<440d", b, off))
    bad = [i for i in range(440) if asuint(ref[i]) != asuint(sc_t[i])]
    print(f"sincos table   : {440 - len(bad)}/440 words identical to libm", bad[:8])
    ok &= not bad
    return ok


def fmt_u64(words, per=4):
    return ",\n".join("  " + ", ".join(f"0x{w:016x}ull" for w in words[i:i + per]) for i in range(0, len(words), per))


def fmt_f64(vals, per=4):
    return ",\n".join("  " + ", ".join(f"{v.hex()}" for v in vals[i:i + per]) for i in range(0, len(vals), per))


def main():
    e, s, p = exp_table(), sincos_table(), powlog_table()
    if "--check" in sys.argv:
        if not check(e, s, p):
            print("MISMATCH against libm")
            sys.exit(1)
    txt = ("// generated by tools/gen_math_tables.py -- do not edit\n"
           "// exp: 2^(k/128) = H(1+T): {asuint64(T), asuint64(H) - (k<<52)/128}\n"
           "#define RH_EXP_TABLE \\\n" + fmt_u64(e).replace("\n", " \\\n") + "\n\n"
           "// sin/cos: {sin hi, sin lo, cos hi, cos lo} at x = k/128, k = 0..109\n"
           "#define RH_SINCOS_TABLE \\\n" + fmt_f64(s).replace("\n", " \\\n") + "\n\n"
           "// pow: {invc, 0, logc, logctail} for the 128 subintervals of [OFF, 2 OFF)\n"
           "#define RH_POWLOG_TABLE \\\n" + fmt_f64(p).replace("\n", " \\\n") + "\n")
    (ROOT / "pyrh_b200" / "csrc" / "rhb200_math_tables.inc").write_text(txt)
    print("written pyrh_b200/csrc/rhb200_math_tables.inc")


if __name__ == "__main__":
    main()

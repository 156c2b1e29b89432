// tests/helpers/mathcheck.cpp -- host build of pyrh_b200/csrc/rhb200_math.cuh compared bit
// for bit with this machine's libm (glibc 2.39: exp/pow/sin/cos/log/log10).  Prints one line per
// (function, range): "<name> <n> <mismatches> <max_ulp>".
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <random>
#include "rhb200_math.cuh"

static uint64_t bits(double x) { uint64_t u; std::memcpy(&u, &x, 8); return u; }
static long ulpdiff(double a, double b)
{
  if (a == b) return 0;
  if (std::isnan(a) && std::isnan(b)) return 0;
  int64_t ia = (int64_t) bits(a), ib = (int64_t) bits(b);
  if (ia < 0) ia = INT64_MIN - ia;
  if (ib < 0) ib = INT64_MIN - ib;
  int64_t d = ia - ib;
  return d < 0 ? -d : d;
}

template <class F, class G, class R>
static long run1(const char *name, long n, F f, G g, R rnd)
{
  long bad = 0, mx = 0;
  for (long i = 0; i < n; i++) {
    double x = rnd();
    double a = f(x), b = g(x);
    if (bits(a) != bits(b) && !(std::isnan(a) && std::isnan(b))) {
      bad++;
      long u = ulpdiff(a, b);
      if (u > mx) mx = u;
      if (bad <= 3) fprintf(stderr, "  %s(%a): got %a want %a\n", name, x, a, b);
    }
  }
  printf("%s %ld %ld %ld\n", name, n, bad, mx);
  return bad;
}

int main(int argc, char **argv)
{
  long n = argc > 1 ? atol(argv[1]) : 2000000;
  std::mt19937_64 gen(12345);
  auto uni = [&](double lo, double hi) { return [&gen, lo, hi]() { return std::uniform_real_distribution<double>(lo, hi)(gen); }; };
  auto logu = [&](double lo, double hi) { return [&gen, lo, hi]() { double e = std::uniform_real_distribution<double>(std::log(lo), std::log(hi))(gen); double s = (gen() & 1) ? -1.0 : 1.0; return s * std::exp(e); }; };
  long bad = 0;
  auto E = [](double x) { return rhm::rh_exp(x); };   auto Eg = [](double x) { return std::exp(x); };
  auto S = [](double x) { return rhm::rh_sin(x); };   auto Sg = [](double x) { return std::sin(x); };
  auto Cc = [](double x) { return rhm::rh_cos(x); };  auto Cg = [](double x) { return std::cos(x); };
  bad += run1("exp[-1,1]", n, E, Eg, uni(-1, 1));
  bad += run1("exp[-60,0]", n, E, Eg, uni(-60, 0));
  bad += run1("exp[0,160]", n, E, Eg, uni(0, 160));
  bad += run1("exp[-760,720]", n, E, Eg, uni(-760, 720));
  bad += run1("exp[log 1e-320..1e3]", n, E, Eg, logu(1e-320, 1e3));
  bad += run1("sin[-0.9,0.9]", n, S, Sg, uni(-0.9, 0.9));
  bad += run1("sin[-12,12]", n, S, Sg, uni(-12, 12));
  bad += run1("sin[-1e4,1e4]", n, S, Sg, uni(-1e4, 1e4));
  bad += run1("sin[log 1e-300..1e8]", n, S, Sg, logu(1e-300, 1.0e8));
  bad += run1("cos[-0.9,0.9]", n, Cc, Cg, uni(-0.9, 0.9));
  bad += run1("cos[-12,12]", n, Cc, Cg, uni(-12, 12));
  bad += run1("cos[-1e4,1e4]", n, Cc, Cg, uni(-1e4, 1e4));
  bad += run1("cos[log 1e-300..1e8]", n, Cc, Cg, logu(1e-300, 1.0e8));
  auto L = [](double x) { return rhm::rh_log(x); };    auto Lg = [](double x) { return std::log(x); };
  auto L10 = [](double x) { return rhm::rh_log10(x); }; auto L10g = [](double x) { return std::log10(x); };
  auto pos = [&](double lo, double hi) { return [&gen, lo, hi]() { double e = std::uniform_real_distribution<double>(std::log(lo), std::log(hi))(gen); return std::exp(e); }; };
  bad += run1("log[0.9,1.1]", n, L, Lg, uni(0.9, 1.1));
  bad += run1("log[0.93,1.07]", n, L, Lg, uni(0.93, 1.07));
  bad += run1("log[1e3,2e4]", n, L, Lg, uni(1e3, 2e4));
  bad += run1("log[log 1e-320..1e300]", n, L, Lg, pos(1e-320, 1e300));
  auto AT = [](double x) { return rhm::rh_atan(x); }; auto ATg = [](double x) { return std::atan(x); };
  bad += run1("atan[-1,1]", n, AT, ATg, uni(-1.0, 1.0));
  bad += run1("atan[-20,20]", n, AT, ATg, uni(-20.0, 20.0));
  bad += run1("atan[log 1e-12..1e20]", n, AT, ATg, logu(1e-12, 1e20));
  bad += run1("log10[0.2,3]", n, L10, L10g, uni(0.2, 3.0));
  bad += run1("log10[log 1e-320..1e300]", n, L10, L10g, pos(1e-320, 1e300));
  {
    const double ys[] = {0.3, 0.38, 0.375, -1.5, 1.5, 0.5, 2.0, -0.25};
    for (double y : ys) {
      char nm[64]; snprintf(nm, sizeof nm, "pow[x in 1e-2..1e6]^%g", y);
      bad += run1(nm, n / 4, [y](double x) { return rhm::rh_pow(x, y); }, [y](double x) { return std::pow(x, y); },
                  [&]() { return std::fabs(logu(1e-2, 1e6)()); });
    }
    std::uniform_real_distribution<double> yd(-40, 40);
    bad += run1("pow[x in 1e-30..1e30]^[-40,40]", n, [&](double x) { double y = yd(gen); return bits(rhm::rh_pow(x, y)) == bits(std::pow(x, y)) ? 0.0 : 1.0; },
                [](double) { return 0.0; }, [&]() { return std::fabs(logu(1e-30, 1e30)()); });
  }
  return bad ? 1 : 0;
}

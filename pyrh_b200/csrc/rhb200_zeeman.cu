// rhb200_zeeman.cu -- host-side Zeeman machinery (runs once per line list; no device code).
//
// Reference: RLKdeterminate  rh/kurucz.c:925-969     term labels -> S, L of both levels
//            RLKZeeman       rh/kurucz.c:832-921     component list (q, shift, strength), normalised per q
//            determinate     rh/zeeman.c:37-85       model-atom level label -> n, S, L, J
//            Zeeman          rh/zeeman.c:186-281     model-atom line: effective triplet or anomalous pattern
//            Lande, ZeemanStrength, getOrbital, getWords   rh/zeeman.c:87-184, 362-397
// Integer work (component count and order, q) is bit-exact by construction; shifts and strengths
// use the reference's expressions in its order (plain IEEE double, host compiler without contraction).
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "rhb200_common.cuh"

namespace {

const double MILLI = 1.0e-3;

// getOrbital, zeeman.c:362-397; -1 = "Invalid orbital" (the reference aborts with ERROR_LEVEL_2)
int get_orbital(char orbit)
{
  static const char letters[] = "SPDFGHIJKLMNOQRTUVWX";
  const char *p = orbit ? strchr(letters, orbit) : nullptr;
  return p ? (int) (p - letters) : -1;
}

// getWords(label, " ", &count), zeeman.c:87-105: strtok semantics (runs of blanks are one separator)
std::vector<std::string> get_words(const std::string &label)
{
  std::vector<std::string> w;
  size_t i = 0;
  while (i < label.size()) {
    while (i < label.size() && label[i] == ' ') i++;
    size_t j = i;
    while (j < label.size() && label[j] != ' ') j++;
    if (j > i) w.emplace_back(label.substr(i, j - i));
    i = j;
  }
  return w;
}

// sscanf(s, "%d%1s", &multiplicity, orbit): optional blanks, signed decimal, blanks, one non-blank char
int scan_mult_orbit(const char *s, int *mult, char *orbit)
{
  int consumed = 0;
  if (sscanf(s, "%d%n", mult, &consumed) != 1) return 0;
  const char *p = s + consumed;
  while (*p && isspace((unsigned char) *p)) p++;
  if (!*p) return 1;
  *orbit = *p;
  return 2;
}

// one level of RLKdeterminate (kurucz.c:934-949)
bool rlk_level(const char *label, double *S, int *L)
{
  std::string l(label);
  if (l.size() > 10) l.resize(10);                       // RLK_LABEL_LENGTH
  const std::vector<std::string> words = get_words(l);
  if (words.empty()) return false;
  const std::string &last = words.back();
  if (last.size() < 2) return false;                     // the reference reads before the word here (UB)
  int mult = 0; char orbit = 0;
  const int nread = scan_mult_orbit(last.c_str() + last.size() - 2, &mult, &orbit);
  if (nread != 2 || !isupper((unsigned char) orbit)) return false;
  const int Lq = get_orbital(orbit);
  if (Lq < 0) return false;
  *L = Lq;
  *S = (mult - 1) / 2.0;
  return true;
}

double lande(double S, int L, double J)                  // zeeman.c:139-146
{
  if (J == 0.0) return 0.0;
  return 1.5 + (S*(S + 1.0) - L*(L + 1)) / (2.0*J*(J + 1.0));
}

#define SQ(x) ((x)*(x))
// ZeemanStrength, zeeman.c:148-184; returns false on "Invalid dJ"
bool zeeman_strength(double Ju, double Mu, double Jl, double Ml, double *s)
{
  const int q = (int) (Ml - Mu), dJ = (int) (Ju - Jl);
  switch (dJ) {
  case 0:
    switch (q) {
    case  0: *s = 2.0 * SQ(Mu); break;
    case -1: *s = (Ju + Mu) * (Ju - Mu + 1.0); break;
    case  1: *s = (Ju - Mu) * (Ju + Mu + 1.0); break;
    }
    return true;
  case 1:
    switch (q) {
    case  0: *s = 2.0 * (SQ(Ju) - SQ(Mu)); break;
    case -1: *s = (Ju + Mu) * (Ju + Mu - 1.0); break;
    case  1: *s = (Ju - Mu) * (Ju - Mu - 1.0); break;
    }
    return true;
  case -1:
    switch (q) {
    case  0: *s = 2.0 * (SQ(Ju + 1.0) - SQ(Mu)); break;
    case -1: *s = (Ju - Mu + 1.0) * (Ju - Mu + 2.0); break;
    case  1: *s = (Ju + Mu + 1.0) * (Ju + Mu + 2.0); break;
    }
    return true;
  }
  return false;
}

// anomalous pattern shared by RLKZeeman (kurucz.c:856-917) and Zeeman (zeeman.c:240-277)
int anomalous(double Jl, double Ju, double gLl, double gLu, int cap, int *q, double *shift, double *strength)
{
  int nc = 0;
  for (double Ml = -Jl; Ml <= Jl; Ml++)
    for (double Mu = -Ju; Mu <= Ju; Mu++)
      if (fabs(Mu - Ml) <= 1.0) nc++;
  if (nc > cap) return nc;
  double norm[3] = {0.0, 0.0, 0.0};
  int n = 0;
  for (double Ml = -Jl; Ml <= Jl; Ml++) {
    for (double Mu = -Ju; Mu <= Ju; Mu++) {
      if (fabs(Mu - Ml) <= 1.0) {
        q[n] = (int) (Ml - Mu);
        shift[n] = gLl*Ml - gLu*Mu;
        if (!zeeman_strength(Ju, Mu, Jl, Ml, &strength[n])) return -1;
        norm[q[n]+1] += strength[n];
        n++;
      }
    }
  }
  for (n = 0; n < nc; n++) strength[n] /= norm[q[n]+1];
  return nc;
}

// determinate, zeeman.c:37-85
bool determinate_level(const char *label, double g, int *nq, double *S, int *L, double *J)
{
  std::string m(label);
  if (m.size() > 20) m.resize(20);                       // ATOM_LABEL_WIDTH
  if (m.empty()) return false;
  size_t p = m.size() - 1;
  while (m[p] != 'E' && m[p] != 'O' && p > 0) p--;
  if (p == 0) return false;                              // "Cannot determine parity of atomic level"
  m.resize(p + 1);
  const std::vector<std::string> words = get_words(m);
  if (words.size() < 2) return false;                    // the reference indexes words[count-2] (UB)
  *nq = 0;
  sscanf(words[words.size()-2].c_str(), "%d", nq);
  const std::string &last = words.back();
  if (last.size() < 3) return false;
  int mult = 0; char orbit = 0;
  if (scan_mult_orbit(last.c_str() + last.size() - 3, &mult, &orbit) != 2) return false;
  const int Lq = get_orbital(orbit);
  if (Lq < 0) return false;
  *S = (mult - 1) / 2.0;
  *L = Lq;
  *J = (g - 1.0) / 2.0;
  if (*J > *L + *S) return false;                        // composite level
  return true;
}

}  // namespace

extern "C" double rhb200_lande(double S, int L, double J) { return lande(S, L, J); }

extern "C" int rhb200_rlk_determinate(const char *labeli, const char *labelj,
                                      double *Si, int *Li, double *Sj, int *Lj)
{
  if (!labeli || !labelj || !Si || !Li || !Sj || !Lj) { rhb200_set_error("bad arguments"); return RHB200_EINVAL; }
  if (!rlk_level(labeli, Si, Li)) return 0;
  if (!rlk_level(labelj, Sj, Lj)) return 0;
  return 1;
}

extern "C" int rhb200_rlk_zeeman(double gi, double gj, double Si, int Li, double Sj, int Lj,
                                 double gL_i, double gL_j, int LS_Lande, int cap,
                                 int *q, double *shift, double *strength)
{
  if (!q || !shift || !strength || cap < 0) { rhb200_set_error("bad arguments"); return RHB200_EINVAL; }
  const double Jl = (gi - 1.0) / 2.0, Ju = (gj - 1.0) / 2.0;
  double gLl, gLu;
  if (LS_Lande) {
    gLl = lande(Si, Li, Jl);
    gLu = lande(Sj, Lj, Ju);
  } else {                                               // -99 in the line list = "not given" (kurucz.c:886-897)
    gLl = (gL_i == -99*MILLI) ? lande(Si, Li, Jl) : gL_i;
    gLu = (gL_j == -99*MILLI) ? lande(Sj, Lj, Ju) : gL_j;
  }
  const int nc = anomalous(Jl, Ju, gLl, gLu, cap, q, shift, strength);
  if (nc < 0) { rhb200_set_error("Invalid dJ: %d", (int) (Ju - Jl)); return RHB200_EINVAL; }
  return nc;
}

extern "C" int rhb200_determinate(const char *label, double g, int *n, double *S, int *L, double *J)
{
  if (!label || !n || !S || !L || !J) { rhb200_set_error("bad arguments"); return RHB200_EINVAL; }
  return determinate_level(label, g, n, S, L, J) ? 1 : 0;
}

extern "C" int rhb200_zeeman(const char *label_i, double g_i, const char *label_j, double g_j,
                             double g_Lande_eff, int cap, int *q, double *shift, double *strength)
{
  if (!label_i || !label_j || !q || !shift || !strength || cap < 0) { rhb200_set_error("bad arguments"); return RHB200_EINVAL; }
  if (g_Lande_eff != 0.0) {                              // normal triplet, zeeman.c:219-231
    if (cap < 3) return 3;
    for (int n = 0; n < 3; n++) { q[n] = -1 + n; strength[n] = 1.0; shift[n] = q[n] * g_Lande_eff; }
    return 3;
  }
  int nq, Ll, Lu; double Sl, Su, Jl, Ju;
  if (!determinate_level(label_i, g_i, &nq, &Sl, &Ll, &Jl) ||
      !determinate_level(label_j, g_j, &nq, &Su, &Lu, &Ju)) {
    rhb200_set_error("cannot determine quantum numbers of the level labels"); return RHB200_EINVAL;
  }
  const int nc = anomalous(Jl, Ju, lande(Sl, Ll, Jl), lande(Su, Lu, Ju), cap, q, shift, strength);
  if (nc < 0) { rhb200_set_error("Invalid dJ: %d", (int) (Ju - Jl)); return RHB200_EINVAL; }
  return nc;
}

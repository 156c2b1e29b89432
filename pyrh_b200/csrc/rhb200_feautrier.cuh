// rhb200_feautrier.cuh -- second-order Feautrier solver (Rybicki-Hummer F/G elimination),
// one thread per ray.  Reference: Feautrier() rh/rhf1d/feautrier.c:56-202 with
// F_order = STANDARD (the only order Formal() requests, formal.c:299).
//
// The reference keeps nine scratch arrays of Ndep doubles per call; here only the two that the
// back-substitution really needs (F and z) are stored, in caller-provided per-ray scratch
// (the IO policy decides where: for the fused LTE path the unused K' slots of the ray's own
// ray-point records), everything else is recomputed from dtau on the fly with the reference's
// expressions, so the results are bit-identical.
#pragma once
#include "rhb200_delo.cuh"

namespace rhf {

// IO policy: chi(k), S(k), z(k) [height]; putF/getF, putZ/getZ scratch; storeP(k, v), storePsi(k, v), wantPsi()
template <class IO>
__device__ __forceinline__ double feautrier_ray(IO &io, const int ndep, const double muz,
                                                const int bc_top, const int bc_bottom,
                                                const double *__restrict__ T, const double lambda)
{
  using rhd::planck;
  const double zmu = 0.5 / muz;
  const int N = ndep;
  auto dtau_at = [&](int k) { return zmu * (io.chi(k) + io.chi(k+1)) * (io.z(k) - io.z(k+1)); };

  const double dtau0 = dtau_at(0), dtauN = dtau_at(N-2);
  double r0 = 0.0, h0 = 0.0, rN = 0.0, hN = 0.0;
  if (bc_top == RHB200_BC_THERMALIZED) {
    const double B0 = planck(T[0], lambda), B1 = planck(T[1], lambda);
    h0 = B0 - (B1 - B0) / dtau0;
  }
  if (bc_bottom == RHB200_BC_THERMALIZED) {
    const double B0 = planck(T[N-2], lambda), B1 = planck(T[N-1], lambda);
    hN = B1 - (B0 - B1) / dtauN;
  }
  const double f0 = (1.0 - r0) / (1.0 + r0), fN = (1.0 - rN) / (1.0 + rN);
  const double abc0 = 1.0 + 2.0*f0 / dtau0, C10 = 2.0 / (dtau0*dtau0);
  const double Stmp0 = io.S(0) + 2.0*h0 / ((1.0 + r0)*dtau0);
  const double abcN = 1.0 + 2.0*fN / dtauN, A1N = 2.0 / (dtauN*dtauN);
  const double StmpN = io.S(N-1) + 2.0*hN / ((1.0 + rN)*dtauN);

  // forward elimination, feautrier.c:160-165
  double F = abc0 / C10, zt = Stmp0 / (abc0 + C10);
  io.putF(0, F); io.putZ(0, zt);
  double dtau_km1 = dtau0;
  for (int k = 1; k < N-1; k++) {
    const double dtau_k = dtau_at(k);
    const double dtau_mid = 0.5*(dtau_k + dtau_km1);
    const double A1 = 1.0 / (dtau_mid * dtau_km1), C1 = 1.0 / (dtau_mid * dtau_k);
    const double Fk = (1.0 + A1*F/(1.0 + F)) / C1;
    zt = (io.S(k) + A1*zt) / (C1 * (1.0 + Fk));
    F = Fk;
    io.putF(k, F); io.putZ(k, zt);
    dtau_km1 = dtau_k;
  }
  // back-substitution, feautrier.c:168-171 (F = F[N-2], zt = ztmp[N-2] here)
  double P = (StmpN + A1N*zt) / (abcN + A1N*(F / (1.0 + F)));
  io.storeP(N-1, P);

  if (!io.wantPsi()) {
    for (int k = N-2; k >= 0; k--) { P = P / (1.0 + io.getF(k)) + io.getZ(k); io.storeP(k, P); }
  } else {
    // diagonal operator, feautrier.c:175-192: G sweep runs with the back-substitution
    io.storePsi(N-1, 1.0 / (abcN + A1N*F/(1.0 + F)));
    double G = abcN / A1N;                       // G[N-1]
    double dtau_k = dtauN;                       // dtau[k] for k = N-2
    for (int k = N-2; k >= 1; k--) {
      const double dkm1 = dtau_at(k-1);
      const double dtau_mid = 0.5*(dtau_k + dkm1);
      const double A1 = 1.0 / (dtau_mid * dkm1), C1 = 1.0 / (dtau_mid * dtau_k);
      const double Fk = io.getF(k), Fkm1 = io.getF(k-1);
      P = P / (1.0 + Fk) + io.getZ(k);
      io.storeP(k, P);
      io.storePsi(k, 1.0 / (1.0 + A1*Fkm1/(1.0 + Fkm1) + C1*G/(1.0 + G)));
      G = (1.0 + C1*G/(1.0 + G)) / A1;           // G[k]
      dtau_k = dkm1;
    }
    P = P / (1.0 + io.getF(0)) + io.getZ(0);
    io.storeP(0, P);
    io.storePsi(0, 1.0 / (abc0 + C10*G/(1.0 + G)));   // G = G[1]
  }
  return (1.0 + f0)*P - h0/(1.0 + r0);           // emergent intensity, feautrier.c:196
}

}  // namespace rhf

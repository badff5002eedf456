// Scalar core of the tridiagonal divide-and-conquer eigensolver (Cuppen 1981; Gu & Eisenstat 1995):
// deflation scan, secular-equation root finder ("middle way" rational interpolation, Li 1993, with
// bracketing), Loewner re-derivation of the rank-one vector.  __host__ __device__ so that the same code
// is validated on the CPU (tests/hostsim) and used by the kernels in eigh.cu.
//
// Replaces the host LAPACK call of the reference (np.linalg.eigh -> syevd, tdvp.py:61-64).
#pragma once
#include <cmath>
#include <cstdint>
#include "flow_core.cuh"  // VMC_HD

namespace vmc {

struct DcRot {  // Givens rotation applied to eigenvector rows (a,b): (a,b) <- (c*a + s*b, c*b - s*a)
  int a, b;
  double c, s;
};

// node [lo,hi) at `depth` levels below the root of the balanced split tree over [0,n)
VMC_HD void dc_node_range(int n, int depth, int index, int& lo, int& hi) {
  lo = 0; hi = n;
  for (int l = depth - 1; l >= 0; --l) {
    const int mid = lo + (hi - lo) / 2;
    if ((index >> l) & 1) lo = mid; else hi = mid;
  }
}
VMC_HD int dc_tree_depth(int n) {  // leaves of size 1 (or 0 for empty halves never happen: sizes >= 1)
  int depth = 0;
  while ((1 << depth) < n) ++depth;
  return depth;
}

// Deflation scan of one merge (LAPACK dlaed2 semantics restated).  Inputs: lam[0..nm) = eigenvalues of the two
// children (each ascending: [0,n1) and [n1,nm)), z[0..nm) = [last row of Q1 ; sign * first row of Q2].
// Outputs: k non-deflated (dl ascending, w, nd_idx = child eigenvector index), nm-k deflated
// (df_val, df_idx), rotation list, rho' and ||.||.  `order` is scratch of nm ints.
VMC_HD void dc_deflate(int nm, int n1, double rho_signed, const double* lam, double* z, double* dmod, int* order,
                       double* dl, double* w, int* nd_idx, double* df_val, int* df_idx, DcRot* rots,
                       int* k_out, int* nrot_out, double* rho_out) {
  const double eps = 1.1102230246251565e-16;
  // merge the two ascending halves
  {
    int i = 0, j = n1, p = 0;
    while (i < n1 && j < nm) order[p++] = (lam[j] < lam[i]) ? j++ : i++;
    while (i < n1) order[p++] = i++;
    while (j < nm) order[p++] = j++;
  }
  const double rho = fabs(2.0 * rho_signed);
  double zmax = 0.0, dmax = 0.0;
  for (int i = 0; i < nm; ++i) {
    z[i] *= 0.70710678118654752440;
    dmod[i] = lam[i];
    zmax = fmax(zmax, fabs(z[i]));
    dmax = fmax(dmax, fabs(lam[i]));
  }
  const double tol = 8.0 * eps * fmax(dmax, zmax);
  int k = 0, ndf = 0, nrot = 0;
  *rho_out = rho;
  if (rho * zmax <= tol) {  // everything deflates
    for (int p = 0; p < nm; ++p) { df_val[ndf] = dmod[order[p]]; df_idx[ndf++] = order[p]; }
    *k_out = 0; *nrot_out = 0;
    return;
  }
  int pj = -1;
  for (int p = 0; p < nm; ++p) {
    const int nj = order[p];
    if (rho * fabs(z[nj]) <= tol) {  // negligible component: deflate
      df_val[ndf] = dmod[nj]; df_idx[ndf++] = nj;
      continue;
    }
    if (pj < 0) { pj = nj; continue; }
    double s = z[pj], c = z[nj];
    const double tau = hypot(c, s);
    const double t = dmod[nj] - dmod[pj];
    c /= tau; s = -s / tau;
    if (fabs(t * c * s) <= tol) {  // close eigenvalues: rotate z[pj] away
      z[nj] = tau; z[pj] = 0.0;
      rots[nrot].a = pj; rots[nrot].b = nj; rots[nrot].c = c; rots[nrot].s = s; ++nrot;
      const double tt = dmod[pj] * c * c + dmod[nj] * s * s;
      dmod[nj] = dmod[pj] * s * s + dmod[nj] * c * c;
      dmod[pj] = tt;
      df_val[ndf] = dmod[pj]; df_idx[ndf++] = pj;
      pj = nj;
    } else {
      dl[k] = dmod[pj]; w[k] = z[pj]; nd_idx[k] = pj; ++k;
      pj = nj;
    }
  }
  if (pj >= 0) { dl[k] = dmod[pj]; w[k] = z[pj]; nd_idx[k] = pj; ++k; }
  *k_out = k; *nrot_out = nrot;
}

// ------------------------------------------------------------------------------------------------
// Secular equation  f(x) = 1/rho + sum_i w2_i / (dl_i - x) = 0, root j in (dl_j, dl_{j+1}) (last: beyond dl_{k-1}).
// The root is returned as (origin o, tau) with x = dl_o + tau so that dl_i - x = (dl_i - dl_o) - tau is accurate.
// `Sum` evaluates the partial sums; the serial version below is used on the host, eigh.cu supplies a warp version.
struct SecularSums {
  double psi, dpsi, phi, dphi, sabs;
};

struct SerialSums {
  const double* dl;
  const double* w2;
  int k;
  // split: indices <= jl go to psi, indices > jl go to phi
  VMC_HD SecularSums operator()(int o, double tau, int jl) const {
    SecularSums r{0, 0, 0, 0, 0};
    const double dlo = dl[o];
    for (int i = 0; i < k; ++i) {
      const double del = (dl[i] - dlo) - tau;
      const double t = w2[i] / del;
      if (i <= jl) { r.psi += t; r.dpsi += t / del; } else { r.phi += t; r.dphi += t / del; }
      r.sabs += fabs(t);
    }
    return r;
  }
};

template <class Sum>
VMC_HD void secular_root(int k, int j, const double* dl, const double* w2, double rho, const Sum& sums,
                         int* origin_out, double* tau_out) {
  const double eps = 1.1102230246251565e-16;
  const double rhoinv = 1.0 / rho;
  if (k == 1) { *origin_out = 0; *tau_out = rho * w2[0]; return; }
  int o, pa, pb;        // origin pole, anchor poles (pa < pb)
  double lb, ub, tau;   // bracket on tau (relative to origin)
  const bool last = (j == k - 1);
  int jl;               // psi takes indices <= jl
  if (!last) {
    pa = j; pb = j + 1; jl = j;
    const double del = dl[j + 1] - dl[j];
    // sign of f at the midpoint decides which pole is nearer the root
    SecularSums m = sums(j, 0.5 * del, jl);
    const double fm = rhoinv + m.psi + m.phi;
    if (fm == 0.0) { *origin_out = j; *tau_out = 0.5 * del; return; }
    if (fm > 0.0) { o = j; lb = 0.0; ub = 0.5 * del; tau = 0.5 * del; }
    else { o = j + 1; lb = -0.5 * del; ub = 0.0; tau = -0.5 * del; }
    // better start: keep the two anchor terms exact, freeze the rest at the midpoint (LAPACK's initial guess)
    {
      const double dA = (dl[pa] - dl[o]), dB = (dl[pb] - dl[o]);
      const double mid_rel = (o == j) ? 0.5 * del : -0.5 * del;
      const double c = fm - w2[pa] / (dA - mid_rel) - w2[pb] / (dB - mid_rel);
      // c + w2a/(dA - x) + w2b/(dB - x) = 0
      const double a = c * (dA + dB) + w2[pa] + w2[pb];
      const double b = c * dA * dB + w2[pa] * dB + w2[pb] * dA;
      double x;
      const double disc = sqrt(fabs(a * a - 4.0 * b * c));
      if (c == 0.0) x = b / a;
      else if (a <= 0.0) x = (a - disc) / (2.0 * c);
      else x = 2.0 * b / (a + disc);
      if (x > lb && x < ub) tau = x;
    }
  } else {
    o = k - 1; pa = k - 2; pb = k - 1; jl = k - 2;
    double wn = 0.0;
    for (int i = 0; i < k; ++i) wn += w2[i];
    lb = 0.0; ub = rho * wn;
    tau = 0.5 * ub;
    if (ub <= 0.0) { *origin_out = o; *tau_out = 0.0; return; }
    {
      const double midp = 0.5 * ub;
      SecularSums m = sums(o, midp, jl);
      const double fm = rhoinv + m.psi + m.phi;
      const double dA = dl[pa] - dl[o];
      const double c = fm - w2[pa] / (dA - midp) - w2[pb] / (-midp);
      // c + w2a/(dA - x) + w2b/(-x) = 0  ->  c x^2 - (c dA + w2a + w2b) x + w2b dA = 0
      const double a = c * dA + w2[pa] + w2[pb], b = w2[pb] * dA;
      const double disc = sqrt(fabs(a * a - 4.0 * b * c));
      const double q = 0.5 * (a + (a >= 0.0 ? disc : -disc));
      const double r1 = (c != 0.0) ? q / c : -1.0, r2 = (q != 0.0) ? b / q : -1.0;
      if (r1 > lb && r1 < ub) tau = r1;
      if (r2 > lb && r2 < ub && (!(r1 > lb && r1 < ub) || r2 < r1)) tau = r2;
    }
  }
  for (int it = 0; it < 80; ++it) {
    const SecularSums s = sums(o, tau, jl);
    const double f = rhoinv + s.psi + s.phi;
    const double erretm = 8.0 * s.sabs + 2.0 * rhoinv + fabs(tau) * (s.dpsi + s.dphi);
    if (fabs(f) <= eps * erretm) break;
    if (f < 0.0) lb = fmax(lb, tau); else ub = fmin(ub, tau);
    if (!(ub - lb > 4.0 * eps * fmax(fabs(lb), fabs(ub)))) break;
    // middle way: psi(x) ~ a1 + s1/(dA - x), phi(x) ~ a2 + s2/(dB - x), matched in value and slope at tau
    const double D1 = (dl[pa] - dl[o]) - tau, D2 = (dl[pb] - dl[o]) - tau;
    const double c = f - D1 * s.dpsi - D2 * s.dphi;
    const double a = (D1 + D2) * f - D1 * D2 * (s.dpsi + s.dphi);
    const double b = D1 * D2 * f;
    // roots of c*eta^2 - a*eta + b = 0 (stable forms); the bracket lies between / beyond the model's poles,
    // where the model is monotone, so at most one root falls inside it
    double eta = 0.0;
    bool ok = false;
    {
      const double disc = sqrt(fabs(a * a - 4.0 * b * c));
      const double q = 0.5 * (a + (a >= 0.0 ? disc : -disc));
      const double r1 = (c != 0.0) ? q / c : 0.0, r2 = (q != 0.0) ? b / q : 0.0;
      const bool v1 = (c != 0.0) && (tau + r1 > lb) && (tau + r1 < ub) && (f * r1 < 0.0);
      const bool v2 = (q != 0.0) && (tau + r2 > lb) && (tau + r2 < ub) && (f * r2 < 0.0);
      if (v1 && v2) { eta = fabs(r1) < fabs(r2) ? r1 : r2; ok = true; }
      else if (v1) { eta = r1; ok = true; }
      else if (v2) { eta = r2; ok = true; }
    }
    if (!ok) eta = -f / (s.dpsi + s.dphi);  // Newton step; bisection below if it leaves the bracket
    double tn = tau + eta;
    if (!(tn > lb && tn < ub)) tn = 0.5 * (lb + ub);                   // safeguard: bisect
    if (tn == tau) break;
    tau = tn;
  }
  *origin_out = o; *tau_out = tau;
}

// delta(i, j) = dl_i - lambda_j, accurate via the origin shift
VMC_HD double dc_delta(const double* dl, int i, int origin_j, double tau_j) { return (dl[i] - dl[origin_j]) - tau_j; }

// Gu-Eisenstat / Loewner: |w_i|^2 = -delta(i,i) * prod_{j != i} delta(i,j) / (dl_i - dl_j); sign from w_i
VMC_HD double dc_lowner_w(int k, int i, const double* dl, const int* origin, const double* tau, double w_i) {
  double p = dc_delta(dl, i, origin[i], tau[i]);
  for (int j = 0; j < k; ++j)
    if (j != i) p *= dc_delta(dl, i, origin[j], tau[j]) / (dl[i] - dl[j]);
  const double v = sqrt(fabs(p));
  return w_i < 0.0 ? -v : v;
}

}  // namespace vmc

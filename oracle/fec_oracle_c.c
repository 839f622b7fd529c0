/*
 * fec_oracle_c.c -- plain-C (OpenMP) restatement of the reference's CPU assembly path.
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product never does.
 *
 * Parity: pinned against oracle/fec_oracle.py (itself pinned to the reference's golden data, see
 * its header) by tests/test_oracle_c.py.  Absolute hex8 / neo-Hookean / J2 values are "parity
 * unpinned" in the reference itself.
 *
 * What is restated (reference path:line):
 *   element loop, one element per work item, static chunks over threads, atomic nodal scatter
 *       src/assemblers/Assemblers.jl:402-432, :72-87 ; src/Utils.jl:11-21,81-113
 *   K_el -> COO slots  storage[(e-1)*NDOF^2 + k] = K_el.data[k] (column-major)
 *       src/assemblers/Assemblers.jl:109-124
 *   K_el * v_el matrix action                       src/assemblers/MatrixAction.jl:208-238
 *   MappedH1OrL2Interpolants                        src/Physics.jl:86-92
 *   physics                                         test/poisson/TestPoissonCommon.jl, test/mechanics/ (physics files)
 *   SparseArrays.sparse!(I,J,V,m,n,+,klasttouch,csrrowptr,csrcolval,csrnzval,csccolptr,cscrowval,cscnzval)
 *       src/assemblers/SparsityPatterns.jl:301-308  (algorithm of the Julia stdlib restated: counting
 *       sort into CSR in COO order, duplicate combination per row, transpose to CSC)
 *   SparseMatrixCSR(csc)                            src/assemblers/SparsityPatterns.jl:326-329
 *
 * Note: the constitutive tangents are analytic here (the reference differentiates psi with
 * Tensors.jl AD) and K_q is formed with the structured loops of Formulations.jl:89-126 rather than
 * the dense G*A*G' products of the return-style physics; both make this baseline FASTER than the
 * reference's own CPU path, never slower.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { PHYS_POISSON = 1, PHYS_LINEAR = 2, PHYS_NEO = 3, PHYS_NEO_AS_WRITTEN = 4, PHYS_J2 = 5 };
#define MAXN 10 /* nodes per element */
#define MAXD 3

typedef struct {
  int nd, nnpe, nf, nq, phys;
  int64_t ne;
  const int64_t* conn; /* 1-based, element-major */
  const double *X, *N, *dN, *w, *props, *source_q, *state_old;
  double* state_new;
  int ns;
} Prob;

static inline double det3(const double F[3][3]) {
  return F[0][0] * (F[1][1] * F[2][2] - F[1][2] * F[2][1]) - F[0][1] * (F[1][0] * F[2][2] - F[1][2] * F[2][0]) +
         F[0][2] * (F[1][0] * F[2][1] - F[1][1] * F[2][0]);
}
static inline double inv_t3(const double F[3][3], double H[3][3]) { /* H = F^-T */
  double J = det3(F), iJ = 1.0 / J;
  H[0][0] = (F[1][1] * F[2][2] - F[1][2] * F[2][1]) * iJ;
  H[0][1] = (F[1][2] * F[2][0] - F[1][0] * F[2][2]) * iJ;
  H[0][2] = (F[1][0] * F[2][1] - F[1][1] * F[2][0]) * iJ;
  H[1][0] = (F[0][2] * F[2][1] - F[0][1] * F[2][2]) * iJ;
  H[1][1] = (F[0][0] * F[2][2] - F[0][2] * F[2][0]) * iJ;
  H[1][2] = (F[0][1] * F[2][0] - F[0][0] * F[2][1]) * iJ;
  H[2][0] = (F[0][1] * F[1][2] - F[0][2] * F[1][1]) * iJ;
  H[2][1] = (F[0][2] * F[1][0] - F[0][0] * F[1][2]) * iJ;
  H[2][2] = (F[0][0] * F[1][1] - F[0][1] * F[1][0]) * iJ;
  return J;
}

/* J2 return map shared by stress and tangent */
typedef struct { double tr, s[3][3], n[3][3], dg, q; int yld; } RM;
static inline void j2_return(const double g[3][3], const double* pr, const double* so, RM* r) {
  const double G = pr[2], sy = pr[3], Hh = pr[4];
  double ep[3][3];
  ep[0][0] = so[0]; ep[1][1] = so[1]; ep[2][2] = so[2];
  ep[1][2] = ep[2][1] = so[3]; ep[0][2] = ep[2][0] = so[4]; ep[0][1] = ep[1][0] = so[5];
  r->tr = g[0][0] + g[1][1] + g[2][2];
  double n2 = 0;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double e = 0.5 * (g[i][j] + g[j][i]) - (i == j ? r->tr / 3.0 : 0.0) - ep[i][j];
      r->s[i][j] = 2 * G * e;
      n2 += r->s[i][j] * r->s[i][j];
    }
  double nrm = sqrt(n2);
  r->q = sqrt(1.5) * nrm;
  double f = r->q - (sy + Hh * so[6]);
  r->yld = f > 0;
  r->dg = r->yld ? f / (3 * G + Hh) : 0.0;
  double inv = nrm > 0 ? 1.0 / nrm : 1.0;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r->n[i][j] = r->s[i][j] * inv;
}

/* P = dpsi/d(grad u) on 3x3 tensors; optionally A[i][j][k][l] = dP_ij/d(grad u)_kl */
static inline void constitutive(int phys, const double g[3][3], const double* pr, const double* so, double* sn,
                                double P[3][3], double (*A)[3][3][3]) {
  if (phys == PHYS_LINEAR) {
    const double K = pr[1], G = pr[2];
    double tr = g[0][0] + g[1][1] + g[2][2];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) P[i][j] = G * (g[i][j] + g[j][i]) + (i == j ? (K - 2 * G / 3) * tr : 0.0);
    if (A)
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
          for (int k = 0; k < 3; k++)
            for (int l = 0; l < 3; l++)
              A[i][j][k][l] = (K - 2 * G / 3) * (i == j) * (k == l) + G * ((i == k) * (j == l) + (i == l) * (j == k));
  } else if (phys == PHYS_NEO || phys == PHYS_NEO_AS_WRITTEN) {
    const double K = pr[1], G = pr[2];
    double F[3][3], H[3][3];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) F[i][j] = g[i][j] + (i == j);
    double J = inv_t3(F, H), I1 = 0;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) I1 += F[i][j] * F[i][j];
    double m = 1.0 / cbrt(J * J), c, cpJ;
    if (phys == PHYS_NEO) { c = 0.5 * K * (J * J - 1); cpJ = K * J * J; }
    else { c = 0.5 * K * (J * J - J - 1); cpJ = 0.5 * K * (2 * J - 1) * J; }
    double gm = G * m;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) P[i][j] = (c - gm * I1 / 3) * H[i][j] + gm * F[i][j];
    if (A)
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
          double dev = F[i][j] - I1 / 3 * H[i][j];
          for (int k = 0; k < 3; k++)
            for (int l = 0; l < 3; l++)
              A[i][j][k][l] = cpJ * H[i][j] * H[k][l] - c * H[i][l] * H[k][j] +
                              gm * (-(2.0 / 3) * dev * H[k][l] + ((i == k) && (j == l)) - (2.0 / 3) * H[i][j] * F[k][l] +
                                    I1 / 3 * H[i][l] * H[k][j]);
        }
  } else if (phys == PHYS_J2) {
    const double K = pr[1], G = pr[2], Hh = pr[4];
    RM r;
    j2_return(g, pr, so, &r);
    double fac = 2 * G * sqrt(1.5) * r.dg;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) P[i][j] = r.s[i][j] - fac * r.n[i][j] + (i == j ? K * r.tr : 0.0);
    if (sn) {
      double de = sqrt(1.5) * r.dg;
      sn[0] = so[0] + de * r.n[0][0]; sn[1] = so[1] + de * r.n[1][1]; sn[2] = so[2] + de * r.n[2][2];
      sn[3] = so[3] + de * r.n[1][2]; sn[4] = so[4] + de * r.n[0][2]; sn[5] = so[5] + de * r.n[0][1];
      sn[6] = so[6] + r.dg;
    }
    if (A) {
      double qs = r.q > 0 ? r.q : 1.0;
      double th = r.yld ? 1 - 3 * G * r.dg / qs : 1.0;
      double tb = r.yld ? 1.0 / (1 + Hh / (3 * G)) - (1 - th) : 0.0;
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
          for (int k = 0; k < 3; k++)
            for (int l = 0; l < 3; l++)
              A[i][j][k][l] = K * (i == j) * (k == l) +
                              2 * G * th * (0.5 * ((i == k) * (j == l) + (i == l) * (j == k)) - (i == j) * (k == l) / 3.0) -
                              2 * G * tb * r.n[i][j] * r.n[k][l];
    }
  }
}

/* geometry of one quadrature point: dN_X[a][k], JxW */
static inline double map_interpolants(const Prob* p, int q, const double x[MAXN][MAXD], double dNX[MAXN][MAXD]) {
  const int nd = p->nd, nn = p->nnpe;
  double J[3][3] = {{0}}, Ji[3][3];
  const double* dN = p->dN + (size_t)q * nn * nd;
  for (int a = 0; a < nn; a++)
    for (int i = 0; i < nd; i++)
      for (int j = 0; j < nd; j++) J[i][j] += x[a][i] * dN[a * nd + j];
  double det;
  if (nd == 2) {
    det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    Ji[0][0] = J[1][1] / det; Ji[0][1] = -J[0][1] / det; Ji[1][0] = -J[1][0] / det; Ji[1][1] = J[0][0] / det;
  } else {
    double H[3][3];
    det = inv_t3(J, H); /* H = J^-T  ->  Ji = H^T */
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) Ji[i][j] = H[j][i];
  }
  for (int a = 0; a < nn; a++)
    for (int k = 0; k < nd; k++) {
      double s = 0;
      for (int j = 0; j < nd; j++) s += dN[a * nd + j] * Ji[j][k];
      dNX[a][k] = s;
    }
  return det * p->w[q];
}

static inline void gather(const Prob* p, int64_t e, const double* U, const double* V, int conn[MAXN],
                          double x[MAXN][MAXD], double u[MAXN][MAXD], double v[MAXN][MAXD]) {
  for (int a = 0; a < p->nnpe; a++) {
    int n = (int)(p->conn[e * p->nnpe + a] - 1);
    conn[a] = n;
    for (int j = 0; j < p->nd; j++) x[a][j] = p->X[(size_t)n * p->nd + j];
    for (int d = 0; d < p->nf; d++) u[a][d] = U[(size_t)n * p->nf + d];
    if (V) for (int d = 0; d < p->nf; d++) v[a][d] = V[(size_t)n * p->nf + d];
  }
}

static inline void grad_u3(const Prob* p, const double u[MAXN][MAXD], const double dNX[MAXN][MAXD], double g[3][3]) {
  memset(g, 0, 9 * sizeof(double));
  for (int a = 0; a < p->nnpe; a++)
    for (int d = 0; d < p->nf; d++)
      for (int k = 0; k < p->nd; k++) g[d][k] += u[a][d] * dNX[a][k];
}

/* K_el (row-major Ke[r*ndof + c]) summed over quadrature points; kind 2 = stiffness, 3 = mass */
static void element_matrix(const Prob* p, int64_t e, int kind, const double x[MAXN][MAXD], const double u[MAXN][MAXD],
                           double* Ke) {
  const int nn = p->nnpe, nf = p->nf, nd = p->nd, ndof = nn * nf;
  memset(Ke, 0, sizeof(double) * ndof * ndof);
  for (int q = 0; q < p->nq; q++) {
    double dNX[MAXN][MAXD];
    double JxW = map_interpolants(p, q, x, dNX);
    if (kind == 3) {
      const double* N = p->N + (size_t)q * nn;
      double rho = (p->phys == PHYS_POISSON ? 1.0 : p->props[0]) * JxW;
      for (int a = 0; a < nn; a++)
        for (int b = 0; b < nn; b++)
          for (int d = 0; d < nf; d++) Ke[(a * nf + d) * ndof + b * nf + d] += rho * N[a] * N[b];
      continue;
    }
    if (p->phys == PHYS_POISSON) {
      for (int a = 0; a < nn; a++)
        for (int b = 0; b < nn; b++) {
          double s = 0;
          for (int j = 0; j < nd; j++) s += dNX[a][j] * dNX[b][j];
          Ke[a * ndof + b] += JxW * s;
        }
    } else {
      double g[3][3], P[3][3], A[3][3][3][3];
      grad_u3(p, u, dNX, g);
      const double* so = p->ns ? p->state_old + ((size_t)e * p->nq + q) * p->ns : NULL;
      constitutive(p->phys, g, p->props, so, NULL, P, A);
      /* scatter_with_gradients_and_gradients! (Formulations.jl:89-126) */
      for (int b = 0; b < nn; b++)
        for (int d2 = 0; d2 < nf; d2++) {
          double t[3][3]; /* t[d1][j1] = sum_j2 A[d1][j1][d2][j2] dNX[b][j2] */
          for (int d1 = 0; d1 < nf; d1++)
            for (int j1 = 0; j1 < nd; j1++) {
              double s = 0;
              for (int j2 = 0; j2 < nd; j2++) s += A[d1][j1][d2][j2] * dNX[b][j2];
              t[d1][j1] = s * JxW;
            }
          for (int a = 0; a < nn; a++)
            for (int d1 = 0; d1 < nf; d1++) {
              double s = 0;
              for (int j1 = 0; j1 < nd; j1++) s += dNX[a][j1] * t[d1][j1];
              Ke[(a * nf + d1) * ndof + b * nf + d2] += s;
            }
        }
    }
  }
}

static void element_residual(const Prob* p, int64_t e, const double x[MAXN][MAXD], const double u[MAXN][MAXD],
                             double* Re) {
  const int nn = p->nnpe, nf = p->nf, nd = p->nd;
  memset(Re, 0, sizeof(double) * nn * nf);
  for (int q = 0; q < p->nq; q++) {
    double dNX[MAXN][MAXD];
    double JxW = map_interpolants(p, q, x, dNX);
    double g[3][3], P[3][3];
    grad_u3(p, u, dNX, g);
    if (p->phys == PHYS_POISSON) {
      const double* N = p->N + (size_t)q * nn;
      double f = p->source_q ? p->source_q[(size_t)e * p->nq + q] : 0.0;
      for (int a = 0; a < nn; a++) {
        double s = 0;
        for (int j = 0; j < nd; j++) s += g[0][j] * dNX[a][j];
        Re[a] += JxW * (s - N[a] * f);
      }
    } else {
      const double* so = p->ns ? p->state_old + ((size_t)e * p->nq + q) * p->ns : NULL;
      double* sn = p->ns ? p->state_new + ((size_t)e * p->nq + q) * p->ns : NULL;
      constitutive(p->phys, g, p->props, so, sn, P, NULL);
      for (int a = 0; a < nn; a++)
        for (int d = 0; d < nf; d++) {
          double s = 0;
          for (int j = 0; j < nd; j++) s += dNX[a][j] * P[d][j];
          Re[a * nf + d] += JxW * s;
        }
    }
  }
}

static Prob make_prob(int nd, int nnpe, int nf, int nq, int phys, int64_t ne, const int64_t* conn, const double* X,
                      const double* N, const double* dN, const double* w, const double* props, const double* source_q,
                      const double* state_old, double* state_new) {
  Prob p;
  p.nd = nd; p.nnpe = nnpe; p.nf = nf; p.nq = nq; p.phys = phys; p.ne = ne; p.conn = conn; p.X = X; p.N = N; p.dN = dN;
  p.w = w; p.props = props; p.source_q = source_q; p.state_old = state_old; p.state_new = state_new;
  p.ns = (phys == PHYS_J2) ? 7 : 0;
  return p;
}

#define PROB_ARGS int nd, int nnpe, int nf, int nq, int phys, int64_t ne, const int64_t* conn, const double* X, \
                  const double* N, const double* dN, const double* w, const double* props, const double* source_q, \
                  const double* state_old, double* state_new
#define PROB_PASS nd, nnpe, nf, nq, phys, ne, conn, X, N, dN, w, props, source_q, state_old, state_new

/* assemble_vector!(storage, ..., residual, ...) : R must hold nf*nn doubles and is zeroed here.
 * state arrays use the reference layout [NS,NQ,NE]. */
int fec_oracle_assemble_vector(PROB_ARGS, const double* U, int64_t ndof, double* R, int nthreads) {
  Prob p = make_prob(PROB_PASS);
  memset(R, 0, sizeof(double) * ndof);
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int64_t e = 0; e < ne; e++) {
    int c[MAXN];
    double x[MAXN][MAXD], u[MAXN][MAXD], Re[MAXN * MAXD];
    gather(&p, e, U, NULL, c, x, u, NULL);
    element_residual(&p, e, x, u, Re);
    for (int d = 0; d < nf; d++)       /* loop order of _assemble_element! (Assemblers.jl:78-85) */
      for (int a = 0; a < nnpe; a++) {
#pragma omp atomic
        R[(size_t)c[a] * nf + d] += Re[a * nf + d];
      }
  }
  return 0;
}

/* assemble_matrix!: coo[e*NDOF^2 + k] = K_el.data[k], K_el column-major (k = r + NDOF*c) */
int fec_oracle_assemble_matrix_coo(PROB_ARGS, const double* U, int kind, double* coo, int nthreads) {
  Prob p = make_prob(PROB_PASS);
  const int ndof = nnpe * nf;
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int64_t e = 0; e < ne; e++) {
    int c[MAXN];
    double x[MAXN][MAXD], u[MAXN][MAXD], Ke[MAXN * MAXD * MAXN * MAXD];
    gather(&p, e, U, NULL, c, x, u, NULL);
    element_matrix(&p, e, kind, x, u, Ke);
    double* out = coo + (size_t)e * ndof * ndof;
    for (int cc = 0; cc < ndof; cc++)
      for (int r = 0; r < ndof; r++) out[r + ndof * cc] = Ke[r * ndof + cc];
  }
  return 0;
}

/* assemble_matrix_action!: Kv_el = K_el * v_el, atomic nodal scatter (MatrixAction.jl:226-236) */
int fec_oracle_assemble_action(PROB_ARGS, const double* U, const double* V, int kind, int64_t ndof_tot, double* out,
                               int nthreads) {
  Prob p = make_prob(PROB_PASS);
  const int ndof = nnpe * nf;
  memset(out, 0, sizeof(double) * ndof_tot);
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int64_t e = 0; e < ne; e++) {
    int c[MAXN];
    double x[MAXN][MAXD], u[MAXN][MAXD], v[MAXN][MAXD], Ke[MAXN * MAXD * MAXN * MAXD];
    gather(&p, e, U, V, c, x, u, v);
    element_matrix(&p, e, kind, x, u, Ke);
    for (int d = 0; d < nf; d++)
      for (int a = 0; a < nnpe; a++) {
        double s = 0;
        for (int b = 0; b < nnpe; b++)
          for (int d2 = 0; d2 < nf; d2++) s += Ke[(a * nf + d) * ndof + b * nf + d2] * v[b][d2];
#pragma omp atomic
        out[(size_t)c[a] * nf + d] += s;
      }
  }
  return 0;
}

/* Is/Js of SparseMatrixPattern(dof) (SparsityPatterns.jl:72-85): element-major, i outer, j inner, 1-based */
int fec_oracle_pattern(int nnpe, int nf, int64_t ne, const int64_t* conn, int64_t* Is, int64_t* Js) {
  const int ndof = nnpe * nf;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < ne; e++) {
    int64_t dc[MAXN * MAXD];
    for (int a = 0; a < nnpe; a++)
      for (int d = 0; d < nf; d++) dc[a * nf + d] = nf * (conn[e * nnpe + a] - 1) + d + 1;
    int64_t* I = Is + (size_t)e * ndof * ndof;
    int64_t* J = Js + (size_t)e * ndof * ndof;
    for (int i = 0; i < ndof; i++)
      for (int j = 0; j < ndof; j++) { I[i * ndof + j] = dc[i]; J[i * ndof + j] = dc[j]; }
  }
  return 0;
}

/* SparseArrays.sparse!(I, J, V, n, n, +, klasttouch, csrrowptr, csrcolval, csrnzval, csccolptr, cscrowval, cscnzval):
 * V[k] = coo[slots[k]-1] (the `coo_storage[pattern.unknown_dofs]` gather), all indices 1-based.
 * Work arrays are caller-provided like the reference's cached pattern arrays:
 *   klasttouch[n], csrrowptr[n+1], csrcolval[ncoo], csrnzval[ncoo]; outputs colptr[n+1], rowval[>=nnz], nzval[>=nnz].
 * Returns nnz. */
int64_t fec_oracle_sparse_csc(int64_t ncoo, const int64_t* Is, const int64_t* Js, const int64_t* slots,
                              const double* coo, int64_t n, int64_t* klasttouch, int64_t* csrrowptr,
                              int64_t* csrcolval, double* csrnzval, int64_t* colptr, int64_t* rowval, double* nzval) {
  /* 1. count entries per row: csrrowptr[r] (r = 1..n) = #entries of row r, then exclusive prefix sum so
   *    that csrrowptr[r] = 0-based offset of the first entry of row r */
  memset(csrrowptr, 0, sizeof(int64_t) * (n + 1));
  for (int64_t k = 0; k < ncoo; k++) csrrowptr[Is[k]]++;
  int64_t run = 0;
  for (int64_t r = 1; r <= n; r++) { int64_t c = csrrowptr[r]; csrrowptr[r] = run; run += c; }
  /* 2. scatter (col, val) into the CSR work arrays in COO order: inside a row, COO order is kept */
  int64_t* cur = (int64_t*)malloc(sizeof(int64_t) * (n + 1));
  for (int64_t r = 1; r <= n; r++) cur[r] = csrrowptr[r];
  for (int64_t k = 0; k < ncoo; k++) {
    int64_t pos = cur[Is[k]]++;
    csrcolval[pos] = Js[k];
    csrnzval[pos] = coo[slots[k] - 1];
  }
  /* 3. combine duplicates row by row: the first occurrence accumulates the later ones, in COO order.
   *    klasttouch[j] = 1-based compacted position where column j was last written. */
  memset(klasttouch, 0, sizeof(int64_t) * n);
  memset(colptr, 0, sizeof(int64_t) * (n + 1));
  int64_t w = 0;
  for (int64_t r = 1; r <= n; r++) {
    const int64_t start = csrrowptr[r], end = cur[r], newstart = w;
    for (int64_t kk = start; kk < end; kk++) {
      const int64_t j = csrcolval[kk];
      if (klasttouch[j - 1] > newstart) {
        csrnzval[klasttouch[j - 1] - 1] += csrnzval[kk];
      } else {
        csrcolval[w] = j;
        csrnzval[w] = csrnzval[kk];
        w++;
        klasttouch[j - 1] = w;
        colptr[j]++; /* entries of column j (1-based) */
      }
    }
    csrrowptr[r - 1] = newstart; /* compacted row pointer, 0-based rows */
  }
  csrrowptr[n] = w;
  free(cur);
  /* 4. transpose compacted CSR -> CSC; rows are visited ascending so rows end up sorted in every column */
  int64_t acc = 1;
  for (int64_t j = 1; j <= n; j++) { int64_t c = colptr[j]; colptr[j - 1] = acc; acc += c; }
  colptr[n] = acc;
  int64_t* ccur = (int64_t*)malloc(sizeof(int64_t) * n);
  for (int64_t j = 0; j < n; j++) ccur[j] = colptr[j] - 1;
  for (int64_t r = 0; r < n; r++)
    for (int64_t kk = csrrowptr[r]; kk < csrrowptr[r + 1]; kk++) {
      const int64_t pos = ccur[csrcolval[kk] - 1]++;
      rowval[pos] = r + 1;
      nzval[pos] = csrnzval[kk];
    }
  free(ccur);
  return w;
}

/* SparseMatrixCSR(csc): rowptr/colval/nzval, 1-based, columns ascending per row */
int fec_oracle_csc_to_csr(int64_t n, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                          int64_t* rowptr, int64_t* colval, double* nzr) {
  int64_t nnz = colptr[n] - 1;
  memset(rowptr, 0, sizeof(int64_t) * (n + 1));
  for (int64_t k = 0; k < nnz; k++) rowptr[rowval[k]]++;
  int64_t acc = 1;
  for (int64_t i = 1; i <= n; i++) { int64_t c = rowptr[i]; rowptr[i - 1] = acc; acc += c; }
  rowptr[n] = acc;
  int64_t* cur = (int64_t*)malloc(sizeof(int64_t) * n);
  for (int64_t i = 0; i < n; i++) cur[i] = rowptr[i] - 1;
  for (int64_t j = 0; j < n; j++)
    for (int64_t k = colptr[j] - 1; k < colptr[j + 1] - 1; k++) {
      int64_t pos = cur[rowval[k] - 1]++;
      colval[pos] = j + 1;
      nzr[pos] = nzval[k];
    }
  free(cur);
  return 0;
}

int fec_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

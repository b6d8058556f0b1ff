/* CPU oracle for the Opt hot path of PartitionedLS.jl -- plain C restatement.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs may load this; libpls_cuda.so never links or calls it.
 *
 * What is restated (paths relative to /root/reference):
 *   - src/PartitionedLS.jl:76-81    homogeneousCoords  (ones column, intercept group)
 *   - src/PartitionedLS.jl:108-123  regularizeProblem  (one sqrt(eta) row per group)
 *   - src/PartitionedLSOpt.jl:4-20  indextobeta        (LSB-first sign bits)
 *   - src/PartitionedLSOpt.jl:22-31 bmatrix            (fresh column-scaled copy per orthant)
 *   - src/PartitionedLSOpt.jl:85-96 the orthant loop, data-space objective, first-minimum argmin
 *   - the external nonneg_lsq(A, b, alg=:nnls) of Opt.jl:89 -- NonNegLeastSquares.jl (compat
 *     "0.4", Project.toml:20; source not under /root/reference).  Its :nnls algorithm is Lawson &
 *     Hanson's active-set NNLS ("Solving Least Squares Problems", 1974, ch. 23) working on the
 *     data matrix with Householder/Givens QR updates; pls_oracle_nnls restates that published
 *     algorithm, so the cost structure (O(N*M') per entering variable) matches the reference's.
 *
 * Parity pinning: see oracle/pls_oracle.py header -- the toy known answer is the only reference
 * fixture reproducible without Julia; everything else is cross-checked against scipy's
 * independent Lawson-Hanson in tests/test_oracle.py.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define A_(i, j) A[(size_t)(j) * (size_t)m + (size_t)(i)]

/* Build the Householder reflector that annihilates v[p+1..m-1].  v[p] is replaced by the new
 * pivot value; the returned scalar `up` together with v[p+1..] defines the reflector. */
static double reflector_build(double *v, int p, int m) {
  double big = 0.0;
  for (int i = p; i < m; ++i) { double a = fabs(v[i]); if (a > big) big = a; }
  if (big <= 0.0) return 0.0;
  double inv = 1.0 / big, s = 0.0;
  for (int i = p; i < m; ++i) { double t = v[i] * inv; s += t * t; }
  double nrm = big * sqrt(s);
  if (v[p] > 0.0) nrm = -nrm;
  double up = v[p] - nrm;
  v[p] = nrm;
  return up;
}

/* Apply the reflector (pivot value u[p], tail u[p+1..], scalar up) to vector c. */
static void reflector_apply(const double *u, double up, int p, int m, double *c) {
  double beta = up * u[p];
  if (beta >= 0.0) return;
  double s = c[p] * up;
  for (int i = p + 1; i < m; ++i) s += c[i] * u[i];
  if (s == 0.0) return;
  s /= beta;
  c[p] += s * up;
  for (int i = p + 1; i < m; ++i) c[i] += s * u[i];
}

static void tri_solve(const double *A, int m, const int *idx, int np, double *z) {
  for (int ip = np - 1; ip >= 0; --ip) {
    if (ip < np - 1) {
      int jn = idx[ip + 1];
      double zn = z[ip + 1];
      for (int i = 0; i <= ip; ++i) z[i] -= A_(i, jn) * zn;
    }
    z[ip] /= A_(ip, idx[ip]);
  }
}

/* Lawson-Hanson NNLS: min ||A x - b||_2 s.t. x >= 0.  A (m x n, column-major) and b are
 * overwritten.  work: w[n], z[m], idx[n].  Returns 0 ok, 3 iteration cap hit. */
int pls_oracle_nnls(double *A, int m, int n, double *b, double *x, double *rnorm,
                    double *w, double *z, int *idx) {
  int np = 0;         /* size of the passive set; idx[0..np) passive, idx[np..n) at bound */
  int iter = 0, itmax = 3 * n, status = 0;
  for (int j = 0; j < n; ++j) { x[j] = 0.0; idx[j] = j; }

  while (np < n && np < m) {
    for (int q = np; q < n; ++q) {
      int j = idx[q];
      double s = 0.0;
      for (int l = np; l < m; ++l) s += A_(l, j) * b[l];
      w[j] = s;
    }
    int qsel = -1, jsel = -1;
    double up = 0.0;
    for (;;) {
      double wmax = 0.0;
      qsel = -1;
      for (int q = np; q < n; ++q) if (w[idx[q]] > wmax) { wmax = w[idx[q]]; qsel = q; }
      if (qsel < 0) goto finish;
      jsel = idx[qsel];
      double keep = A_(np, jsel);
      up = reflector_build(&A_(0, jsel), np, m);
      double un = 0.0;
      for (int l = 0; l < np; ++l) un += A_(l, jsel) * A_(l, jsel);
      un = sqrt(un);
      volatile double probe = un + fabs(A_(np, jsel)) * 0.01;
      if (probe - un > 0.0) {
        memcpy(z, b, sizeof(double) * (size_t)m);
        reflector_apply(&A_(0, jsel), up, np, m, z);
        if (z[np] / A_(np, jsel) > 0.0) break;   /* column accepted */
      }
      A_(np, jsel) = keep;                        /* rejected: near-dependent or wrong sign */
      w[jsel] = 0.0;
    }
    memcpy(b, z, sizeof(double) * (size_t)m);
    idx[qsel] = idx[np];
    idx[np] = jsel;
    ++np;
    for (int q = np; q < n; ++q) reflector_apply(&A_(0, jsel), up, np - 1, m, &A_(0, idx[q]));
    for (int l = np; l < m; ++l) A_(l, jsel) = 0.0;
    w[jsel] = 0.0;
    tri_solve(A, m, idx, np, z);

    for (;;) {
      if (++iter > itmax) { status = 3; goto finish; }
      double alpha = 2.0;
      int qout = -1;
      for (int ip = 0; ip < np; ++ip) {
        if (z[ip] <= 0.0) {
          int l = idx[ip];
          double t = -x[l] / (z[ip] - x[l]);
          if (alpha > t) { alpha = t; qout = ip; }
        }
      }
      if (qout < 0) break;
      for (int ip = 0; ip < np; ++ip) { int l = idx[ip]; x[l] += alpha * (z[ip] - x[l]); }
      int iout = idx[qout];
      for (;;) {
        x[iout] = 0.0;
        /* delete column at position qout: shift the later passive columns one place down and
         * restore the triangle with plane rotations applied to all columns and to b */
        for (int jp = qout + 1; jp < np; ++jp) {
          int ii = idx[jp];
          idx[jp - 1] = ii;
          double a1 = A_(jp - 1, ii), a2 = A_(jp, ii), cc, ss, sig;
          if (fabs(a1) > fabs(a2)) {
            double xr = a2 / a1, yr = sqrt(1.0 + xr * xr);
            cc = copysign(1.0 / yr, a1); ss = cc * xr; sig = fabs(a1) * yr;
          } else if (a2 != 0.0) {
            double xr = a1 / a2, yr = sqrt(1.0 + xr * xr);
            ss = copysign(1.0 / yr, a2); cc = ss * xr; sig = fabs(a2) * yr;
          } else { cc = 0.0; ss = 1.0; sig = 0.0; }
          A_(jp - 1, ii) = sig;
          A_(jp, ii) = 0.0;
          for (int l = 0; l < n; ++l) {
            if (l == ii) continue;
            double t1 = A_(jp - 1, l), t2 = A_(jp, l);
            A_(jp - 1, l) = cc * t1 + ss * t2;
            A_(jp, l) = -ss * t1 + cc * t2;
          }
          double t1 = b[jp - 1], t2 = b[jp];
          b[jp - 1] = cc * t1 + ss * t2;
          b[jp] = -ss * t1 + cc * t2;
        }
        --np;
        idx[np] = iout;
        qout = -1;
        for (int ip = 0; ip < np; ++ip) if (x[idx[ip]] <= 0.0) { qout = ip; break; }
        if (qout < 0) break;
        iout = idx[qout];
      }
      memcpy(z, b, sizeof(double) * (size_t)m);
      tri_solve(A, m, idx, np, z);
    }
    for (int ip = 0; ip < np; ++ip) x[idx[ip]] = z[ip];
  }
finish:;
  double s = 0.0;
  for (int l = np; l < m; ++l) s += b[l] * b[l];
  if (rnorm) *rnorm = sqrt(s);
  return status;
}
#undef A_

/* ------------------------------------------------------------------------------------------
 * Opt loop.  X is N x M column-major (Julia layout), P is M x K column-major int64 0/1.
 * b_list == NULL: enumerate b = 0 .. 2^(K+1)-1 (nb ignored);  else solve the nb listed orthants.
 * obj_out[nb], alpha_out[nb*(M+1)] (nullable): per-orthant objective / raw alpha (Opt.jl:92).
 * b_best / obj_best / alpha_best[(M+1)] (nullable): first-minimum winner over the solved set.
 * nthreads: worker threads over orthants (the reference loop itself is serial; <=0 = all cores).
 * Returns 0, or -1 on allocation failure / bad args.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const double *Xa, *ya;
  const int64_t *P, *b_list;
  int64_t Na, M, K, nb;
  double *objs, *alpha_out;
  int64_t next;          /* shared work counter */
  int fail;
} opt_job;

static void *opt_worker(void *arg) {
  opt_job *J = (opt_job *)arg;
  const int64_t Na = J->Na, M = J->M, K = J->K, Mp = M + 1;
  double *Xb = (double *)malloc(sizeof(double) * (size_t)Na * (size_t)Mp);
  double *yb = (double *)malloc(sizeof(double) * (size_t)Na);
  double *alpha = (double *)malloc(sizeof(double) * (size_t)Mp);
  double *d = (double *)malloc(sizeof(double) * (size_t)Mp);
  double *w = (double *)malloc(sizeof(double) * (size_t)Mp);
  double *z = (double *)malloc(sizeof(double) * (size_t)Na);
  int *idx = (int *)malloc(sizeof(int) * (size_t)Mp);
  if (!Xb || !yb || !alpha || !d || !w || !z || !idx) {
    __atomic_store_n(&J->fail, 1, __ATOMIC_RELAXED);
  } else {
    for (;;) {
      int64_t q = __atomic_fetch_add(&J->next, 1, __ATOMIC_RELAXED);
      if (q >= J->nb) break;
      int64_t b = J->b_list ? J->b_list[q] : q;
      /* d = Po * beta, beta_k = 2*bit_k(b) - 1  (Opt.jl:4-20, 28-29) */
      for (int64_t j = 0; j < M; ++j) {
        double s = 0.0;
        for (int64_t k = 0; k < K; ++k)
          if (J->P[k * M + j]) s += (double)J->P[k * M + j] * (double)(2 * ((b >> k) & 1) - 1);
        d[j] = s;
      }
      d[M] = (double)(2 * ((b >> K) & 1) - 1);
      /* Xb = Xa .* d'  -- fresh materialised copy, as bmatrix does (Opt.jl:30) */
      for (int64_t j = 0; j < Mp; ++j) {
        const double dj = d[j];
        const double *src = J->Xa + j * Na;
        double *dst = Xb + j * Na;
        for (int64_t i = 0; i < Na; ++i) dst[i] = src[i] * dj;
      }
      memcpy(yb, J->ya, sizeof(double) * (size_t)Na);
      pls_oracle_nnls(Xb, (int)Na, (int)Mp, yb, alpha, NULL, w, z, idx);
      /* objective norm(Xa * (d .* alpha) - ya) in data space, eta rows included (Opt.jl:90) */
      for (int64_t i = 0; i < Na; ++i) z[i] = -J->ya[i];
      for (int64_t j = 0; j < Mp; ++j) {
        double wj = d[j] * alpha[j];
        if (wj == 0.0) continue;
        const double *src = J->Xa + j * Na;
        for (int64_t i = 0; i < Na; ++i) z[i] += src[i] * wj;
      }
      double s = 0.0;
      for (int64_t i = 0; i < Na; ++i) s += z[i] * z[i];
      J->objs[q] = sqrt(s);
      if (J->alpha_out) memcpy(J->alpha_out + q * Mp, alpha, sizeof(double) * (size_t)Mp);
    }
  }
  free(Xb); free(yb); free(alpha); free(d); free(w); free(z); free(idx);
  return NULL;
}

int pls_oracle_num_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

int pls_oracle_opt_fit(const double *X, int64_t N, int64_t M, const double *y, const int64_t *P,
                       int64_t K, double eta, const int64_t *b_list, int64_t nb, int nthreads,
                       double *obj_out, double *alpha_out, int64_t *b_best, double *obj_best,
                       double *alpha_best) {
  if (N <= 0 || M <= 0 || K <= 0 || K > 40) return -1;
  const int64_t Mp = M + 1, Kp = K + 1;
  const int64_t Na = N + (eta != 0.0 ? Kp : 0);
  if (!b_list) nb = (int64_t)1 << Kp;
  /* Xa: homogeneous (PartitionedLS.jl:76-81) + regularised (:108-123) data, column-major */
  double *Xa = (double *)calloc((size_t)Na * (size_t)Mp, sizeof(double));
  double *ya = (double *)calloc((size_t)Na, sizeof(double));
  double *objs = obj_out ? obj_out : (double *)malloc(sizeof(double) * (size_t)nb);
  if (!Xa || !ya || !objs) return -1;
  for (int64_t j = 0; j < M; ++j) memcpy(Xa + j * Na, X + j * N, sizeof(double) * (size_t)N);
  for (int64_t i = 0; i < N; ++i) Xa[M * Na + i] = 1.0;
  memcpy(ya, y, sizeof(double) * (size_t)N);
  if (eta != 0.0) {
    double se = sqrt(eta);
    for (int64_t k = 0; k < K; ++k)
      for (int64_t j = 0; j < M; ++j)
        if (P[k * M + j] == 1) Xa[j * Na + N + k] = se;
    Xa[M * Na + N + K] = se;
  }
  if (nthreads <= 0) nthreads = pls_oracle_num_threads();
  if (nthreads > nb) nthreads = (int)nb;
  if (nthreads > 256) nthreads = 256;
  opt_job J = {Xa, ya, P, b_list, Na, M, K, nb, objs, alpha_out, 0, 0};
  pthread_t tid[256];
  int started = 0;
  for (int t = 1; t < nthreads; ++t)
    if (pthread_create(&tid[started], NULL, opt_worker, &J) == 0) ++started;
  opt_worker(&J);
  for (int t = 0; t < started; ++t) pthread_join(tid[t], NULL);

  int rc = J.fail ? -1 : 0;
  if (!rc && (b_best || obj_best || alpha_best)) {
    int64_t best = 0;
    for (int64_t q = 0; q < nb; ++q) {
      if (isnan(objs[q])) { best = q; break; }        /* Julia argmin: NaN wins (Opt.jl:96) */
      if (objs[q] < objs[best]) best = q;
    }
    if (b_best) *b_best = b_list ? b_list[best] : best;
    if (obj_best) *obj_best = objs[best];
    if (alpha_best) {
      if (alpha_out) memcpy(alpha_best, alpha_out + best * Mp, sizeof(double) * (size_t)Mp);
      else {
        int64_t bb = b_list ? b_list[best] : best;    /* alphas not kept: re-solve the winner */
        double o1;
        rc = pls_oracle_opt_fit(X, N, M, y, P, K, eta, &bb, 1, 1, &o1, alpha_best, NULL, NULL, NULL);
      }
    }
  }
  free(Xa); free(ya);
  if (!obj_out) free(objs);
  return rc;
}

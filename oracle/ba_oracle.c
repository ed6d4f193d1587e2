/*
 * ORACLE (test infrastructure only -- never linked or called by the product path).
 *
 * Plain-C, single-thread, fp64 restatement of the reference's sliding-window bundle adjustment:
 *   optimize_map        /root/reference/src/stereo_visual_slam_main/optimization.cpp:103-288
 *   optimize_pose_only  /root/reference/src/stereo_visual_slam_main/optimization.cpp:290-436
 * with the reference's own vertex/edge callbacks
 *   VertexPose::oplusImpl      optimization.cpp:26-32   T <- exp([upsilon;omega]) * T   (Sophus, left-multiplicative)
 *   VertexXYZ::oplusImpl       optimization.cpp:34-39   p += delta
 *   EdgeProjection             optimization.cpp:41-73   e = z - pi(K (T p)), analytic 2x6 / 2x3 Jacobians (Zinv = 1/(Z+1e-18))
 *   PoseOnlyEdgeProjection     optimization.cpp:75-101  same residual, 2x6 Jacobian without the epsilon guard
 * and of the g2o machinery those calls run inside.  g2o is an un-vendored, unpinned dependency that is absent from
 * /root/reference and from this container (README.md:75, CMakeLists.txt:155,180), so its published algorithm is
 * restated here: OptimizationAlgorithmLevenberg::solve (tau = 1e-5, nu = 2, good-step scale
 * clamp [1/3, 2/3], at most 10 trials), BlockSolver<6,3>::buildSystem / solve (Schur complement over the marginalised
 * landmarks, 3x3 direct inverse), RobustKernelHuber (delta = 5.991 used as delta, first-order weight only),
 * SparseOptimizer::optimize (stops on Terminate), and the reference's adaptive chi2 relabel loop
 * (optimization.cpp:224-266).  The linear solver is a dense Cholesky; g2o's CSparse (AMD-ordered) / Eigen LDLT
 * differ from it only in rounding.
 *
 * PARITY UNPINNED BY THE REFERENCE: the reference has no tests or golden vectors and no g2o binary exists here.
 * Pinned instead by tests/test_oracle_ba.py: analytic Jacobians vs finite differences, Schur solve vs the full
 * dense normal equations, monotone chi2 on accepted steps, and agreement of the converged optimum with
 * scipy.optimize.least_squares on the same Huber-ised residuals.
 *
 * Conventions: poses are T_c_w as 12 doubles row-major [R | t]; observations are listed in graph-insertion order;
 * the landmark inlier flag is decided by the LAST observation of that landmark in this order (the reference iterates
 * a std::map keyed by edge pointer, optimization.cpp:156,254-266 -- monotone heap allocation yields insertion order).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    double huber_delta;   /* 5.991 (optimization.cpp:154,205) */
    double chi2_th;       /* 5.991 initial relabel threshold (optimization.cpp:154) */
    int num_iterations;   /* optimizer.optimize(num_ite) */
    int pose_only;        /* 0: optimize_map, 1: optimize_pose_only */
    int max_trials;       /* g2o maxTrialsAfterFailure = 10 */
    double tau;           /* g2o initial lambda factor 1e-5 */
} ba_options;

typedef struct {
    int iterations;       /* outer LM iterations executed */
    int trials;           /* total LM trials (linear solves) */
    int accepted;         /* accepted trials */
    double chi2_initial;  /* robustified */
    double chi2_final;    /* robustified, at the returned estimate */
    double lambda_final;
    double chi2_threshold; /* after the adaptive doubling loop */
    int n_inlier_obs, n_outlier_obs;
} ba_result;

static void mat3_mul(const double* A, const double* B, double* C) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}

/* Sophus::SE3d::exp: tangent = [upsilon; omega]; returns R (3x3) and t */
void ba_se3_exp(const double* xi, double* R, double* t) {
    const double* u = xi;
    const double* w = xi + 3;
    const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    const double th = sqrt(th2);
    const double O[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    double O2[9];
    mat3_mul(O, O, O2);
    /* SO3::exp through the unit quaternion, as Sophus does */
    double imag, real;
    if (th < 1e-10) {
        const double th4 = th2 * th2;
        imag = 0.5 - th2 / 48.0 + th4 / 3840.0;
        real = 1.0 - th2 / 8.0 + th4 / 384.0;
    } else {
        imag = sin(0.5 * th) / th;
        real = cos(0.5 * th);
    }
    const double qx = imag * w[0], qy = imag * w[1], qz = imag * w[2], qw = real;
    R[0] = 1 - 2 * (qy * qy + qz * qz); R[1] = 2 * (qx * qy - qz * qw);     R[2] = 2 * (qx * qz + qy * qw);
    R[3] = 2 * (qx * qy + qz * qw);     R[4] = 1 - 2 * (qx * qx + qz * qz); R[5] = 2 * (qy * qz - qx * qw);
    R[6] = 2 * (qx * qz - qy * qw);     R[7] = 2 * (qy * qz + qx * qw);     R[8] = 1 - 2 * (qx * qx + qy * qy);
    double V[9];
    if (th < 1e-10) {
        memcpy(V, R, sizeof(V));
    } else {
        const double a = (1 - cos(th)) / th2, b = (th - sin(th)) / (th2 * th);
        for (int i = 0; i < 9; ++i) V[i] = a * O[i] + b * O2[i];
        V[0] += 1; V[4] += 1; V[8] += 1;
    }
    for (int i = 0; i < 3; ++i) t[i] = V[i * 3] * u[0] + V[i * 3 + 1] * u[1] + V[i * 3 + 2] * u[2];
}

/* T <- exp(xi) * T   (optimization.cpp:31) */
static void pose_oplus(double* T, const double* xi) {
    double dR[9], dt[3], Rn[9], tn[3];
    ba_se3_exp(xi, dR, dt);
    double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
    double t[3] = {T[3], T[7], T[11]};
    mat3_mul(dR, R, Rn);
    for (int i = 0; i < 3; ++i) tn[i] = dR[i * 3] * t[0] + dR[i * 3 + 1] * t[1] + dR[i * 3 + 2] * t[2] + dt[i];
    T[0] = Rn[0]; T[1] = Rn[1]; T[2] = Rn[2]; T[3] = tn[0];
    T[4] = Rn[3]; T[5] = Rn[4]; T[6] = Rn[5]; T[7] = tn[1];
    T[8] = Rn[6]; T[9] = Rn[7]; T[10] = Rn[8]; T[11] = tn[2];
}

/* e = z - (K (T p)) / (.)_z   (optimization.cpp:41-50, 75-82); also returns the camera-frame point */
void ba_residual(const double* T, const double* p, const double* K, const double* z, double* e, double* pc) {
    for (int i = 0; i < 3; ++i) pc[i] = T[i * 4] * p[0] + T[i * 4 + 1] * p[1] + T[i * 4 + 2] * p[2] + T[i * 4 + 3];
    double q[3];
    for (int i = 0; i < 3; ++i) q[i] = K[i * 3] * pc[0] + K[i * 3 + 1] * pc[1] + K[i * 3 + 2] * pc[2];
    e[0] = z[0] - q[0] / q[2];
    e[1] = z[1] - q[1] / q[2];
}

/* A = d e / d pose (2x6), B = d e / d point (2x3)   (optimization.cpp:52-73 / 84-101) */
void ba_jacobians(const double* T, const double* pc, const double* K, int pose_only, double* A, double* B) {
    const double fx = K[0], fy = K[4];
    const double X = pc[0], Y = pc[1], Z = pc[2];
    if (!pose_only) {
        const double Zinv = 1.0 / (Z + 1e-18), Zinv2 = Zinv * Zinv;
        A[0] = -fx * Zinv; A[1] = 0; A[2] = fx * X * Zinv2; A[3] = fx * X * Y * Zinv2; A[4] = -fx - fx * X * X * Zinv2; A[5] = fx * Y * Zinv;
        A[6] = 0; A[7] = -fy * Zinv; A[8] = fy * Y * Zinv2; A[9] = fy + fy * Y * Y * Zinv2; A[10] = -fy * X * Y * Zinv2; A[11] = -fy * X * Zinv;
        for (int r = 0; r < 2; ++r)
            for (int c = 0; c < 3; ++c)
                B[r * 3 + c] = A[r * 6] * T[c] + A[r * 6 + 1] * T[4 + c] + A[r * 6 + 2] * T[8 + c];
    } else {
        const double Z2 = Z * Z;
        A[0] = -fx / Z; A[1] = 0; A[2] = fx * X / Z2; A[3] = fx * X * Y / Z2; A[4] = -fx - fx * X * X / Z2; A[5] = fx * Y / Z;
        A[6] = 0; A[7] = -fy / Z; A[8] = fy * Y / (Z * Z); A[9] = fy + fy * Y * Y / Z2; A[10] = -fy * X * Y / Z2; A[11] = -fy * X / Z;
    }
}

/* g2o RobustKernelHuber::robustify, rho[0], rho[1] */
static void huber(double e2, double delta, double* rho0, double* rho1) {
    const double dsqr = delta * delta;
    if (e2 <= dsqr) {
        *rho0 = e2;
        *rho1 = 1.0;
    } else {
        const double s = sqrt(e2);
        *rho0 = 2 * s * delta - dsqr;
        *rho1 = delta / s;
    }
}

/* dense Cholesky solve of the SPD system M x = b (M is n x n row-major, destroyed); returns 0 on failure */
static int chol_solve(double* M, const double* b, double* x, int n) {
    for (int j = 0; j < n; ++j) {
        double d = M[j * n + j];
        for (int k = 0; k < j; ++k) d -= M[j * n + k] * M[j * n + k];
        if (!(d > 0.0) || !isfinite(d)) return 0;
        d = sqrt(d);
        M[j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double s = M[i * n + j];
            for (int k = 0; k < j; ++k) s -= M[i * n + k] * M[j * n + k];
            M[i * n + j] = s / d;
        }
    }
    for (int i = 0; i < n; ++i) {
        double s = b[i];
        for (int k = 0; k < i; ++k) s -= M[i * n + k] * x[k];
        x[i] = s / M[i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = x[i];
        for (int k = i + 1; k < n; ++k) s -= M[k * n + i] * x[k];
        x[i] = s / M[i * n + i];
    }
    return 1;
}

static int inv3(const double* m, double* o) {
    const double c0 = m[4] * m[8] - m[5] * m[7], c1 = m[5] * m[6] - m[3] * m[8], c2 = m[3] * m[7] - m[4] * m[6];
    const double det = m[0] * c0 + m[1] * c1 + m[2] * c2;
    const double id = 1.0 / det;
    o[0] = c0 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    o[3] = c1 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    o[6] = c2 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
    return isfinite(id);
}

typedef struct {
    int K, L, n_obs, pose_only;
    const int32_t *op, *ol;
    const double *uv, *Kc;
    double delta;
    double *poses, *points; /* current estimate */
    double *err;            /* 2 per obs: the edges' _error */
    double *Hpp, *bp;       /* 6K x 6K (only diagonal blocks filled by buildSystem), 6K */
    double *Hll, *bl;       /* L x 9, L x 3 */
    double *Hpl;            /* n_obs x 18 (6x3 per observation) */
    double *x;              /* 6K + 3L */
    int *lm_start, *lm_obs; /* CSR landmark -> observations */
} ba_sys;

/* computeActiveErrors + activeRobustChi2 */
static double compute_errors(ba_sys* s) {
    double chi = 0;
    for (int i = 0; i < s->n_obs; ++i) {
        double pc[3];
        ba_residual(s->poses + 12 * s->op[i], s->points + 3 * s->ol[i], s->Kc, s->uv + 2 * i, s->err + 2 * i, pc);
        const double e2 = s->err[2 * i] * s->err[2 * i] + s->err[2 * i + 1] * s->err[2 * i + 1];
        double r0, r1;
        huber(e2, s->delta, &r0, &r1);
        chi += r0;
    }
    return chi;
}

/* BlockSolver::buildSystem: linearizeOplus + constructQuadraticForm per edge (errors from compute_errors) */
static void build_system(ba_sys* s) {
    const int n = 6 * s->K;
    memset(s->Hpp, 0, sizeof(double) * n * n);
    memset(s->bp, 0, sizeof(double) * n);
    if (!s->pose_only) {
        memset(s->Hll, 0, sizeof(double) * 9 * s->L);
        memset(s->bl, 0, sizeof(double) * 3 * s->L);
    }
    for (int i = 0; i < s->n_obs; ++i) {
        const int k = s->op[i], l = s->ol[i];
        const double* T = s->poses + 12 * k;
        const double* p = s->points + 3 * l;
        double pc[3], A[12], B[6];
        for (int r = 0; r < 3; ++r) pc[r] = T[r * 4] * p[0] + T[r * 4 + 1] * p[1] + T[r * 4 + 2] * p[2] + T[r * 4 + 3];
        ba_jacobians(T, pc, s->Kc, s->pose_only, A, B);
        const double* e = s->err + 2 * i;
        double r0, w;
        huber(e[0] * e[0] + e[1] * e[1], s->delta, &r0, &w);
        const double om0 = -w * e[0], om1 = -w * e[1]; /* omega_r = -rho1 * Omega * e */
        for (int a = 0; a < 6; ++a) {
            s->bp[6 * k + a] += A[a] * om0 + A[6 + a] * om1;
            for (int b = 0; b < 6; ++b) s->Hpp[(6 * k + a) * n + 6 * k + b] += w * (A[a] * A[b] + A[6 + a] * A[6 + b]);
        }
        if (!s->pose_only) {
            for (int a = 0; a < 3; ++a) {
                s->bl[3 * l + a] += B[a] * om0 + B[3 + a] * om1;
                for (int b = 0; b < 3; ++b) s->Hll[9 * l + a * 3 + b] += w * (B[a] * B[b] + B[3 + a] * B[3 + b]);
            }
            for (int a = 0; a < 6; ++a)
                for (int b = 0; b < 3; ++b) s->Hpl[18 * i + a * 3 + b] = w * (A[a] * B[b] + A[6 + a] * B[3 + b]);
        }
    }
}

/* BlockSolver::solve with lambda already conceptually added to every diagonal entry of Hpp and Hll */
static int solve_system(ba_sys* s, double lambda, double* S, double* bs) {
    const int n = 6 * s->K;
    memcpy(S, s->Hpp, sizeof(double) * n * n);
    for (int i = 0; i < n; ++i) S[i * n + i] += lambda;
    memcpy(bs, s->bp, sizeof(double) * n);
    double* Dinv = NULL;
    if (!s->pose_only) {
        Dinv = (double*)malloc(sizeof(double) * 9 * s->L);
        for (int l = 0; l < s->L; ++l) {
            if (s->lm_start[l + 1] == s->lm_start[l]) {
                memset(Dinv + 9 * l, 0, 72);
                continue;
            }
            double D[9];
            memcpy(D, s->Hll + 9 * l, 72);
            D[0] += lambda; D[4] += lambda; D[8] += lambda;
            inv3(D, Dinv + 9 * l);
            const double* di = Dinv + 9 * l;
            const double* b = s->bl + 3 * l;
            const double db[3] = {di[0] * b[0] + di[1] * b[1] + di[2] * b[2], di[3] * b[0] + di[4] * b[1] + di[5] * b[2],
                                  di[6] * b[0] + di[7] * b[1] + di[8] * b[2]};
            for (int oi = s->lm_start[l]; oi < s->lm_start[l + 1]; ++oi) {
                const int i = s->lm_obs[oi], ki = s->op[i];
                const double* Bi = s->Hpl + 18 * i;
                double BD[18];
                for (int a = 0; a < 6; ++a)
                    for (int c = 0; c < 3; ++c)
                        BD[a * 3 + c] = Bi[a * 3] * di[c] + Bi[a * 3 + 1] * di[3 + c] + Bi[a * 3 + 2] * di[6 + c];
                for (int a = 0; a < 6; ++a) bs[6 * ki + a] -= Bi[a * 3] * db[0] + Bi[a * 3 + 1] * db[1] + Bi[a * 3 + 2] * db[2];
                for (int oj = s->lm_start[l]; oj < s->lm_start[l + 1]; ++oj) {
                    const int j = s->lm_obs[oj], kj = s->op[j];
                    const double* Bj = s->Hpl + 18 * j;
                    for (int a = 0; a < 6; ++a)
                        for (int b = 0; b < 6; ++b)
                            S[(6 * ki + a) * n + 6 * kj + b] -= BD[a * 3] * Bj[b * 3] + BD[a * 3 + 1] * Bj[b * 3 + 1] + BD[a * 3 + 2] * Bj[b * 3 + 2];
                }
            }
        }
    }
    int ok = chol_solve(S, bs, s->x, n);
    if (ok && !s->pose_only) {
        for (int l = 0; l < s->L; ++l) {
            double cl[3] = {s->bl[3 * l], s->bl[3 * l + 1], s->bl[3 * l + 2]};
            for (int oi = s->lm_start[l]; oi < s->lm_start[l + 1]; ++oi) {
                const int i = s->lm_obs[oi], ki = s->op[i];
                const double* Bi = s->Hpl + 18 * i;
                for (int c = 0; c < 3; ++c)
                    for (int a = 0; a < 6; ++a) cl[c] -= Bi[a * 3 + c] * s->x[6 * ki + a];
            }
            const double* di = Dinv + 9 * l;
            for (int c = 0; c < 3; ++c) s->x[n + 3 * l + c] = di[c * 3] * cl[0] + di[c * 3 + 1] * cl[1] + di[c * 3 + 2] * cl[2];
        }
    }
    free(Dinv);
    return ok;
}

/* Exposed for tests: robustified chi2 at a given estimate. */
double ba_oracle_chi2(int K, const double* poses, int L, const double* points, int n_obs, const int32_t* op,
                      const int32_t* ol, const double* uv, const double* Kc, double delta) {
    (void)K; (void)L;
    double chi = 0;
    for (int i = 0; i < n_obs; ++i) {
        double e[2], pc[3], r0, r1;
        ba_residual(poses + 12 * op[i], points + 3 * ol[i], Kc, uv + 2 * i, e, pc);
        huber(e[0] * e[0] + e[1] * e[1], delta, &r0, &r1);
        chi += r0;
    }
    return chi;
}

/*
 * The whole of optimize_map / optimize_pose_only minus the container walking:
 * poses (K x 12) and points (L x 3) are updated in place (the caller decides about if_update_map /
 * if_update_landmark write-back, optimization.cpp:272-287, 429-435).
 * chi2_per_obs (n_obs, may be NULL) receives edge->chi2() as the relabel loop sees it; point_inlier (L, may be NULL)
 * receives the relabelled Landmark::is_inlier (points without observations keep the value passed in).
 * trace (may be NULL) receives, per LM trial, {lambda, chi2_trial, rho, accepted}; trace_cap rows of 4 doubles.
 */
int ba_oracle_optimize(int K, double* poses, int L, double* points, int n_obs, const int32_t* op, const int32_t* ol,
                       const double* uv, const double* Kc, const ba_options* opt, ba_result* res,
                       double* chi2_per_obs, uint8_t* point_inlier, double* trace, int trace_cap) {
    ba_sys s;
    memset(&s, 0, sizeof(s));
    s.K = K; s.L = L; s.n_obs = n_obs; s.pose_only = opt->pose_only;
    s.op = op; s.ol = ol; s.uv = uv; s.Kc = Kc; s.delta = opt->huber_delta;
    s.poses = poses; s.points = points;
    const int n = 6 * K, nx = n + (opt->pose_only ? 0 : 3 * L);
    s.err = (double*)calloc(2 * (size_t)n_obs + 2, sizeof(double));
    s.Hpp = (double*)calloc((size_t)n * n, sizeof(double));
    s.bp = (double*)calloc(n, sizeof(double));
    s.Hll = (double*)calloc(9 * (size_t)L + 9, sizeof(double));
    s.bl = (double*)calloc(3 * (size_t)L + 3, sizeof(double));
    s.Hpl = (double*)calloc(18 * (size_t)n_obs + 18, sizeof(double));
    s.x = (double*)calloc((size_t)n + 3 * L + 3, sizeof(double));
    s.lm_start = (int*)calloc(L + 2, sizeof(int));
    s.lm_obs = (int*)calloc(n_obs + 1, sizeof(int));
    for (int i = 0; i < n_obs; ++i) s.lm_start[ol[i] + 1]++;
    for (int l = 0; l < L; ++l) s.lm_start[l + 1] += s.lm_start[l];
    {
        int* fill = (int*)calloc(L + 1, sizeof(int));
        for (int i = 0; i < n_obs; ++i) s.lm_obs[s.lm_start[ol[i]] + fill[ol[i]]++] = i;
        free(fill);
    }
    double* S = (double*)calloc((size_t)n * n, sizeof(double));
    double* bs = (double*)calloc(n, sizeof(double));
    double* save_p = (double*)malloc(sizeof(double) * 12 * K);
    double* save_l = (double*)malloc(sizeof(double) * 3 * (L + 1));

    double lambda = 0, ni = 2;
    int trials = 0, accepted = 0, it = 0, ntrace = 0;
    double chi_first = 0, chi_last = 0;
    for (it = 0; it < opt->num_iterations; ++it) {
        double currentChi = compute_errors(&s);
        if (it == 0) chi_first = currentChi;
        double tempChi = currentChi;
        build_system(&s);
        if (it == 0) { /* computeLambdaInit: tau * max |diagonal| over the active vertices */
            double md = 0;
            for (int i = 0; i < n; ++i) md = fmax(md, fabs(s.Hpp[i * n + i]));
            if (!opt->pose_only)
                for (int l = 0; l < L; ++l)
                    for (int c = 0; c < 3; ++c) md = fmax(md, fabs(s.Hll[9 * l + c * 4]));
            lambda = opt->tau * md;
            ni = 2;
        }
        double rho = 0;
        int qmax = 0;
        do {
            memcpy(save_p, poses, sizeof(double) * 12 * K); /* push() */
            memcpy(save_l, points, sizeof(double) * 3 * L);
            const int ok2 = solve_system(&s, lambda, S, bs);
            if (ok2) { /* _optimizer->update(x) */
                for (int k = 0; k < K; ++k) pose_oplus(poses + 12 * k, s.x + 6 * k);
                if (!opt->pose_only)
                    for (int i = 0; i < 3 * L; ++i) points[i] += s.x[n + i];
            } else {
                memset(s.x, 0, sizeof(double) * nx);
            }
            tempChi = compute_errors(&s);
            if (!ok2) tempChi = DBL_MAX;
            rho = currentChi - tempChi;
            double scale = 0;
            for (int j = 0; j < n; ++j) scale += s.x[j] * (lambda * s.x[j] + s.bp[j]);
            if (!opt->pose_only)
                for (int j = 0; j < 3 * L; ++j) scale += s.x[n + j] * (lambda * s.x[n + j] + s.bl[j]);
            scale += 1e-3;
            rho /= scale;
            const int good = rho > 0 && isfinite(tempChi);
            if (trace && ntrace < trace_cap) {
                trace[4 * ntrace] = lambda; trace[4 * ntrace + 1] = tempChi; trace[4 * ntrace + 2] = rho; trace[4 * ntrace + 3] = good;
                ntrace++;
            }
            if (good) {
                double alpha = 1. - pow(2 * rho - 1, 3);
                alpha = fmin(alpha, 2. / 3.);
                lambda *= fmax(1. / 3., alpha);
                ni = 2;
                currentChi = tempChi;
                accepted++;
            } else {
                lambda *= ni;
                ni *= 2;
                memcpy(poses, save_p, sizeof(double) * 12 * K); /* pop(): estimates restored, edge errors are NOT */
                memcpy(points, save_l, sizeof(double) * 3 * L);
            }
            qmax++;
            trials++;
        } while (rho < 0 && qmax < opt->max_trials);
        chi_last = currentChi;
        if (qmax == opt->max_trials || rho == 0) { /* Terminate */
            it++;
            break;
        }
    }
    if (opt->num_iterations <= 0) { /* optimize(0): errors are never computed by g2o; define them at the input */
        chi_first = chi_last = compute_errors(&s);
    }

    /* adaptive chi2 relabel (optimization.cpp:224-266 / 382-424) on the edges' stored errors */
    double th = opt->chi2_th;
    int cin = 0, cout = 0;
    for (int r = 0; r < 5; ++r) {
        cin = cout = 0;
        for (int i = 0; i < n_obs; ++i) {
            const double c = s.err[2 * i] * s.err[2 * i] + s.err[2 * i + 1] * s.err[2 * i + 1];
            if (c > th) cout++; else cin++;
        }
        const double ratio = cin / (double)(cin + cout);
        if (ratio > 0.5) break;
        th *= 2;
    }
    if (n_obs > 0 && cin + cout == n_obs) { /* recount at the final threshold when the loop ended by exhaustion */
        cin = cout = 0;
        for (int i = 0; i < n_obs; ++i) {
            const double c = s.err[2 * i] * s.err[2 * i] + s.err[2 * i + 1] * s.err[2 * i + 1];
            if (c > th) cout++; else cin++;
        }
    }
    for (int i = 0; i < n_obs; ++i) {
        const double c = s.err[2 * i] * s.err[2 * i] + s.err[2 * i + 1] * s.err[2 * i + 1];
        if (chi2_per_obs) chi2_per_obs[i] = c;
        if (point_inlier) point_inlier[ol[i]] = c > th ? 0 : 1; /* last observation wins */
    }
    if (res) {
        res->iterations = it;
        res->trials = trials;
        res->accepted = accepted;
        res->chi2_initial = chi_first;
        res->chi2_final = chi_last;
        res->lambda_final = lambda;
        res->chi2_threshold = th;
        res->n_inlier_obs = cin;
        res->n_outlier_obs = cout;
    }
    free(s.err); free(s.Hpp); free(s.bp); free(s.Hll); free(s.bl); free(s.Hpl); free(s.x); free(s.lm_start); free(s.lm_obs);
    free(S); free(bs); free(save_p); free(save_l);
    return 0;
}

/* Full (un-Schur'd) dense normal equations at the current estimate, for the Schur cross-check in the tests:
 * H is (6K+3L)^2 row-major, b is 6K+3L. */
int ba_oracle_dense_system(int K, const double* poses, int L, const double* points, int n_obs, const int32_t* op,
                           const int32_t* ol, const double* uv, const double* Kc, double delta, double* H, double* b) {
    const int n = 6 * K, N = n + 3 * L;
    memset(H, 0, sizeof(double) * (size_t)N * N);
    memset(b, 0, sizeof(double) * N);
    for (int i = 0; i < n_obs; ++i) {
        const int k = op[i], l = ol[i];
        double e[2], pc[3], A[12], B[6], r0, w;
        ba_residual(poses + 12 * k, points + 3 * l, Kc, uv + 2 * i, e, pc);
        ba_jacobians(poses + 12 * k, pc, Kc, 0, A, B);
        huber(e[0] * e[0] + e[1] * e[1], delta, &r0, &w);
        double J[2][9];
        int idx[9];
        for (int a = 0; a < 6; ++a) { J[0][a] = A[a]; J[1][a] = A[6 + a]; idx[a] = 6 * k + a; }
        for (int a = 0; a < 3; ++a) { J[0][6 + a] = B[a]; J[1][6 + a] = B[3 + a]; idx[6 + a] = n + 3 * l + a; }
        for (int a = 0; a < 9; ++a) {
            b[idx[a]] += -w * (J[0][a] * e[0] + J[1][a] * e[1]);
            for (int c = 0; c < 9; ++c) H[(size_t)idx[a] * N + idx[c]] += w * (J[0][a] * J[0][c] + J[1][a] * J[1][c]);
        }
    }
    return 0;
}

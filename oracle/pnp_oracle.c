/*
 * ORACLE (test infrastructure only -- never linked or called by the product path).
 *
 * Plain-C, single-thread restatement of
 *     cv::solvePnPRansac(pts3d f32, pts2d f32, K, noDist, rvec, tvec, false, 100, 4.0, 0.99, inliers)
 * as VO::motion_estimation calls it (/root/reference/src/stereo_visual_slam_main/visual_odometry.cpp:277; inlier
 * list consumed at :295-311).  OpenCV is an un-vendored dependency of the reference (README.md:72 pins 3.2; the
 * parity target here is cv2 4.13.0 as installed), so its published algorithm is restated:
 *   calib3d/src/solvepnp.cpp   solvePnPRansac: 5-point minimal sample, EPnP kernel, refit of the inliers with
 *                              SOLVEPNP_ITERATIVE; PnPRansacCallback::computeError (float32 projections)
 *   calib3d/src/ptsetreg.cpp   RANSACPointSetRegistrator::run / getSubset / findInliers, RANSACUpdateNumIters
 *   calib3d/src/epnp.cpp       epnp::compute_pose and everything below it
 *   core/src/lapack.cpp        JacobiSVDImpl_<double> (matrices below the LAPACK hand-over size), SVBkSb
 *   core/include/.../core.hpp  cv::RNG (multiply-with-carry), seed (uint64)-1
 * PINNED against live cv2 4.13.0 by tests/test_oracle_pnp.py: cv2.SVDecomp (Jacobi path), cv2.solvePnP(EPNP) on
 * the very subsets the RANSAC draws, cv2.projectPoints (float32 errors) and cv2.solvePnPRansac (inlier lists
 * index for index, pose).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- cv::RNG -------------------------------------- */
typedef struct { uint64_t state; } cv_rng;
static unsigned rng_next(cv_rng* r) {
    r->state = (uint64_t)(unsigned)r->state * 4164903690U + (unsigned)(r->state >> 32);
    return (unsigned)r->state;
}
static int rng_uniform(cv_rng* r, int a, int b) { return a == b ? a : (int)(rng_next(r) % (unsigned)(b - a) + a); }

/* ------------------------------------------------- JacobiSVDImpl_<double> (lapack.cpp) ------------------------ */
/* lapack.cpp has its own hypot (not libm's): the rotation angles, hence every bit of the result, depend on it */
static double cv_hypot(double a, double b) {
    a = fabs(a); b = fabs(b);
    if (a > b) { b /= a; return a * sqrt(1 + b * b); }
    if (b > 0) { a /= b; return b * sqrt(1 + a * a); }
    return 0;
}

/* At: n rows of length m (the transposed input), overwritten by U^T rows; W: n singular values (descending);
 * Vt: n x n (rows = right singular vectors) or NULL; n1 = number of U^T rows to normalise. */
static void jacobi_svd(double* At, int astep, double* Wout, double* Vt, int vstep, int m, int n, int n1) {
    const double eps = DBL_EPSILON * 10, minval = DBL_MIN;
    double W[16];
    int i, j, k, iter, max_iter = m > 30 ? m : 30;
    double c, s, sd;
    for (i = 0; i < n; i++) {
        for (k = 0, sd = 0; k < m; k++) { double t = At[i * astep + k]; sd += t * t; }
        W[i] = sd;
        if (Vt) { for (k = 0; k < n; k++) Vt[i * vstep + k] = 0; Vt[i * vstep + i] = 1; }
    }
    for (iter = 0; iter < max_iter; iter++) {
        int changed = 0;
        for (i = 0; i < n - 1; i++)
            for (j = i + 1; j < n; j++) {
                double *Ai = At + i * astep, *Aj = At + j * astep;
                double a = W[i], p = 0, b = W[j];
                for (k = 0; k < m; k++) p += Ai[k] * Aj[k];
                if (fabs(p) <= eps * sqrt(a * b)) continue;
                p *= 2;
                double beta = a - b, gamma = cv_hypot(p, beta);
                if (beta < 0) {
                    double delta = (gamma - beta) * 0.5;
                    s = sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
                a = b = 0;
                for (k = 0; k < m; k++) {
                    double t0 = c * Ai[k] + s * Aj[k];
                    double t1 = -s * Ai[k] + c * Aj[k];
                    Ai[k] = t0; Aj[k] = t1;
                    a += t0 * t0; b += t1 * t1;
                }
                W[i] = a; W[j] = b;
                changed = 1;
                if (Vt) {
                    double *Vi = Vt + i * vstep, *Vj = Vt + j * vstep;
                    for (k = 0; k < n; k++) {
                        double t0 = c * Vi[k] + s * Vj[k];
                        double t1 = -s * Vi[k] + c * Vj[k];
                        Vi[k] = t0; Vj[k] = t1;
                    }
                }
            }
        if (!changed) break;
    }
    for (i = 0; i < n; i++) {
        for (k = 0, sd = 0; k < m; k++) { double t = At[i * astep + k]; sd += t * t; }
        W[i] = sqrt(sd);
    }
    for (i = 0; i < n - 1; i++) {
        j = i;
        for (k = i + 1; k < n; k++)
            if (W[j] < W[k]) j = k;
        if (i != j) {
            double t = W[i]; W[i] = W[j]; W[j] = t;
            if (Vt) {
                for (k = 0; k < m; k++) { t = At[i * astep + k]; At[i * astep + k] = At[j * astep + k]; At[j * astep + k] = t; }
                for (k = 0; k < n; k++) { t = Vt[i * vstep + k]; Vt[i * vstep + k] = Vt[j * vstep + k]; Vt[j * vstep + k] = t; }
            }
        }
    }
    for (i = 0; i < n; i++) Wout[i] = W[i];
    if (!Vt) return;
    cv_rng rng = {0x12345678};
    for (i = 0; i < n1; i++) {
        sd = i < n ? W[i] : 0;
        for (int ii = 0; ii < 100 && sd <= minval; ii++) {
            /* zero singular value: random vector, orthogonalised against the previous left vectors */
            const double val0 = 1. / m;
            for (k = 0; k < m; k++) {
                double val = (rng_next(&rng) & 256) != 0 ? val0 : -val0;
                At[i * astep + k] = val;
            }
            for (iter = 0; iter < 2; iter++) {
                for (j = 0; j < i; j++) {
                    sd = 0;
                    for (k = 0; k < m; k++) sd += At[i * astep + k] * At[j * astep + k];
                    double asum = 0;
                    for (k = 0; k < m; k++) {
                        double t = At[i * astep + k] - sd * At[j * astep + k];
                        At[i * astep + k] = t;
                        asum += fabs(t);
                    }
                    asum = asum > eps * 100 ? 1 / asum : 0;
                    for (k = 0; k < m; k++) At[i * astep + k] *= asum;
                }
            }
            sd = 0;
            for (k = 0; k < m; k++) { double t = At[i * astep + k]; sd += t * t; }
            sd = sqrt(sd);
        }
        s = sd > minval ? 1 / sd : 0.;
        for (k = 0; k < m; k++) At[i * astep + k] *= s;
    }
}

/* cv::SVD::compute(A (m x n, m >= n), w, u, vt) through the Jacobi path: returns w[n], Ut (n rows of length m = U^T)
 * and Vt (n x n).  Exposed for the tests (compared with cv2.SVDecomp). */
void pnp_oracle_svd(const double* A, int m, int n, double* w, double* Ut, double* Vt) {
    for (int i = 0; i < n; i++)
        for (int k = 0; k < m; k++) Ut[i * m + k] = A[k * n + i]; /* transpose(src, temp_a) */
    jacobi_svd(Ut, m, w, Vt, n, m, n, n);
}

/* SVBkSbImpl_ with nb == 1: x = V diag(1/w) U^T b, singular values <= sum(w)*2*DBL_EPSILON dropped */
static void svbksb(int m, int n, const double* w, const double* Ut, const double* Vt, const double* b, double* x) {
    double threshold = 0;
    int nm = m < n ? m : n;
    for (int i = 0; i < n; i++) x[i] = 0;
    for (int i = 0; i < nm; i++) threshold += w[i];
    threshold *= DBL_EPSILON * 2;
    for (int i = 0; i < nm; i++) {
        double wi = w[i];
        if (fabs(wi) <= threshold) continue;
        wi = 1 / wi;
        double s = 0;
        for (int j = 0; j < m; j++) s += Ut[i * m + j] * b[j];
        s *= wi;
        for (int j = 0; j < n; j++) x[j] = x[j] + s * Vt[i * n + j];
    }
}

/* cv::solve(A (m x n), b, x, DECOMP_SVD) */
static void solve_svd(const double* A, int m, int n, const double* b, double* x) {
    double Ut[6 * 6], Vt[6 * 6], w[6];
    pnp_oracle_svd(A, m, n, w, Ut, Vt);
    svbksb(m, n, w, Ut, Vt, b, x);
}

/* cv::invert(A 3x3, DECOMP_SVD): SVD::backSubst with an identity right-hand side (b == NULL branch: nb = m) */
static void invert3_svd(const double* A, double* Ainv) {
    double Ut[9], Vt[9], w[3];
    pnp_oracle_svd(A, 3, 3, w, Ut, Vt);
    double threshold = (w[0] + w[1] + w[2]) * DBL_EPSILON * 2;
    for (int i = 0; i < 9; i++) Ainv[i] = 0;
    for (int i = 0; i < 3; i++) {
        double wi = w[i];
        if (fabs(wi) <= threshold) continue;
        wi = 1 / wi;
        /* buffer[j] = u[j] * wi (b == NULL: column j of U^T b is u_i[j]); x[j][k] += v[j] * buffer[k] */
        for (int j = 0; j < 3; j++)
            for (int k = 0; k < 3; k++) Ainv[j * 3 + k] += Vt[i * 3 + j] * (Ut[i * 3 + k] * wi);
    }
}

/* MulTransposedR: dst = src^T src, src is rows x cols; upper triangle summed over the rows in order, then mirrored */
static void mul_transposed(const double* src, int rows, int cols, double* dst) {
    for (int i = 0; i < cols; i++)
        for (int j = i; j < cols; j++) {
            double s0 = 0;
            for (int k = 0; k < rows; k++) s0 += src[k * cols + i] * src[k * cols + j];
            dst[i * cols + j] = s0;
            dst[j * cols + i] = s0;
        }
}

/* ------------------------------------------------------- epnp.cpp ---------------------------------------------- */
#define EPNP_MAXN 16
typedef struct {
    double uc, vc, fu, fv;
    int n;
    double pws[3 * EPNP_MAXN], us[2 * EPNP_MAXN], alphas[4 * EPNP_MAXN], pcs[3 * EPNP_MAXN];
    double cws[4][3], ccs[4][3];
} epnp_t;

static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double dist2(const double* p1, const double* p2) {
    return (p1[0] - p2[0]) * (p1[0] - p2[0]) + (p1[1] - p2[1]) * (p1[1] - p2[1]) + (p1[2] - p2[2]) * (p1[2] - p2[2]);
}

static void choose_control_points(epnp_t* e) {
    const int n = e->n;
    e->cws[0][0] = e->cws[0][1] = e->cws[0][2] = 0;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < 3; j++) e->cws[0][j] += e->pws[3 * i + j];
    for (int j = 0; j < 3; j++) e->cws[0][j] /= n;
    double pw0[3 * EPNP_MAXN], pw0tpw0[9], dc[3], uct[9];
    for (int i = 0; i < n; i++)
        for (int j = 0; j < 3; j++) pw0[3 * i + j] = e->pws[3 * i + j] - e->cws[0][j];
    mul_transposed(pw0, n, 3, pw0tpw0);
    /* cvSVD(&PW0tPW0, &DC, &UCt, 0, CV_SVD_MODIFY_A | CV_SVD_U_T) */
    double At[9];
    for (int i = 0; i < 3; i++)
        for (int k = 0; k < 3; k++) At[i * 3 + k] = pw0tpw0[k * 3 + i];
    double Vt[9];
    jacobi_svd(At, 3, dc, Vt, 3, 3, 3, 3);
    memcpy(uct, At, sizeof(uct));
    for (int i = 1; i < 4; i++) {
        double k = sqrt(dc[i - 1] / n);
        for (int j = 0; j < 3; j++) e->cws[i][j] = e->cws[0][j] + k * uct[3 * (i - 1) + j];
    }
}

static void compute_barycentric_coordinates(epnp_t* e) {
    double cc[9], ci[9];
    for (int i = 0; i < 3; i++)
        for (int j = 1; j < 4; j++) cc[3 * i + j - 1] = e->cws[j][i] - e->cws[0][i];
    invert3_svd(cc, ci);
    for (int i = 0; i < e->n; i++) {
        const double* pi = e->pws + 3 * i;
        double* a = e->alphas + 4 * i;
        for (int j = 0; j < 3; j++)
            a[1 + j] = ci[3 * j] * (pi[0] - e->cws[0][0]) + ci[3 * j + 1] * (pi[1] - e->cws[0][1]) +
                       ci[3 * j + 2] * (pi[2] - e->cws[0][2]);
        a[0] = 1.0f - a[1] - a[2] - a[3];
    }
}

static void fill_M(const epnp_t* e, double* M, int row, const double* as, double u, double v) {
    double* M1 = M + row * 12;
    double* M2 = M1 + 12;
    for (int i = 0; i < 4; i++) {
        M1[3 * i] = as[i] * e->fu; M1[3 * i + 1] = 0.0; M1[3 * i + 2] = as[i] * (e->uc - u);
        M2[3 * i] = 0.0; M2[3 * i + 1] = as[i] * e->fv; M2[3 * i + 2] = as[i] * (e->vc - v);
    }
}

static void compute_ccs(epnp_t* e, const double* betas, const double* ut) {
    for (int i = 0; i < 4; i++) e->ccs[i][0] = e->ccs[i][1] = e->ccs[i][2] = 0.0;
    for (int i = 0; i < 4; i++) {
        const double* v = ut + 12 * (11 - i);
        for (int j = 0; j < 4; j++)
            for (int k = 0; k < 3; k++) e->ccs[j][k] += betas[i] * v[3 * j + k];
    }
}

static void compute_pcs(epnp_t* e) {
    for (int i = 0; i < e->n; i++) {
        const double* a = e->alphas + 4 * i;
        double* pc = e->pcs + 3 * i;
        for (int j = 0; j < 3; j++)
            pc[j] = a[0] * e->ccs[0][j] + a[1] * e->ccs[1][j] + a[2] * e->ccs[2][j] + a[3] * e->ccs[3][j];
    }
}

static void solve_for_sign(epnp_t* e) {
    if (e->pcs[2] < 0.0) {
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 3; j++) e->ccs[i][j] = -e->ccs[i][j];
        for (int i = 0; i < 3 * e->n; i++) e->pcs[i] = -e->pcs[i];
    }
}

static void estimate_R_and_t(const epnp_t* e, double R[3][3], double t[3]) {
    const int n = e->n;
    double pc0[3] = {0, 0, 0}, pw0[3] = {0, 0, 0};
    for (int i = 0; i < n; i++)
        for (int j = 0; j < 3; j++) { pc0[j] += e->pcs[3 * i + j]; pw0[j] += e->pws[3 * i + j]; }
    for (int j = 0; j < 3; j++) { pc0[j] /= n; pw0[j] /= n; }
    double abt[9] = {0};
    for (int i = 0; i < n; i++) {
        const double* pc = e->pcs + 3 * i;
        const double* pw = e->pws + 3 * i;
        for (int j = 0; j < 3; j++) {
            abt[3 * j] += (pc[j] - pc0[j]) * (pw[0] - pw0[0]);
            abt[3 * j + 1] += (pc[j] - pc0[j]) * (pw[1] - pw0[1]);
            abt[3 * j + 2] += (pc[j] - pc0[j]) * (pw[2] - pw0[2]);
        }
    }
    /* cvSVD(&ABt, &ABt_D, &ABt_U, &ABt_V, CV_SVD_MODIFY_A): U and V (not transposed) */
    double w[3], Ut[9], Vt[9];
    pnp_oracle_svd(abt, 3, 3, w, Ut, Vt);
    /* R[i][j] = dot(row i of U, row j of V) */
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            R[i][j] = Ut[0 * 3 + i] * Vt[0 * 3 + j] + Ut[1 * 3 + i] * Vt[1 * 3 + j] + Ut[2 * 3 + i] * Vt[2 * 3 + j];
    const double det = R[0][0] * R[1][1] * R[2][2] + R[0][1] * R[1][2] * R[2][0] + R[0][2] * R[1][0] * R[2][1] -
                       R[0][2] * R[1][1] * R[2][0] - R[0][1] * R[1][0] * R[2][2] - R[0][0] * R[1][2] * R[2][1];
    if (det < 0) { R[2][0] = -R[2][0]; R[2][1] = -R[2][1]; R[2][2] = -R[2][2]; }
    t[0] = pc0[0] - dot3(R[0], pw0);
    t[1] = pc0[1] - dot3(R[1], pw0);
    t[2] = pc0[2] - dot3(R[2], pw0);
}

static double reprojection_error(const epnp_t* e, double R[3][3], const double t[3]) {
    double sum2 = 0.0;
    for (int i = 0; i < e->n; i++) {
        const double* pw = e->pws + 3 * i;
        double Xc = dot3(R[0], pw) + t[0];
        double Yc = dot3(R[1], pw) + t[1];
        double inv_Zc = 1.0 / (dot3(R[2], pw) + t[2]);
        double ue = e->uc + e->fu * Xc * inv_Zc;
        double ve = e->vc + e->fv * Yc * inv_Zc;
        double u = e->us[2 * i], v = e->us[2 * i + 1];
        sum2 += sqrt((u - ue) * (u - ue) + (v - ve) * (v - ve));
    }
    return sum2 / e->n;
}

static double compute_R_and_t(epnp_t* e, const double* ut, const double* betas, double R[3][3], double t[3]) {
    compute_ccs(e, betas, ut);
    compute_pcs(e);
    solve_for_sign(e);
    estimate_R_and_t(e, R, t);
    return reprojection_error(e, R, t);
}

static void compute_L_6x10(const double* ut, double* l_6x10) {
    const double* v[4] = {ut + 12 * 11, ut + 12 * 10, ut + 12 * 9, ut + 12 * 8};
    double dv[4][6][3];
    for (int i = 0; i < 4; i++) {
        int a = 0, b = 1;
        for (int j = 0; j < 6; j++) {
            dv[i][j][0] = v[i][3 * a] - v[i][3 * b];
            dv[i][j][1] = v[i][3 * a + 1] - v[i][3 * b + 1];
            dv[i][j][2] = v[i][3 * a + 2] - v[i][3 * b + 2];
            b++;
            if (b > 3) { a++; b = a + 1; }
        }
    }
    for (int i = 0; i < 6; i++) {
        double* row = l_6x10 + 10 * i;
        row[0] = dot3(dv[0][i], dv[0][i]);
        row[1] = 2.0f * dot3(dv[0][i], dv[1][i]);
        row[2] = dot3(dv[1][i], dv[1][i]);
        row[3] = 2.0f * dot3(dv[0][i], dv[2][i]);
        row[4] = 2.0f * dot3(dv[1][i], dv[2][i]);
        row[5] = dot3(dv[2][i], dv[2][i]);
        row[6] = 2.0f * dot3(dv[0][i], dv[3][i]);
        row[7] = 2.0f * dot3(dv[1][i], dv[3][i]);
        row[8] = 2.0f * dot3(dv[2][i], dv[3][i]);
        row[9] = dot3(dv[3][i], dv[3][i]);
    }
}

static void compute_rho(const epnp_t* e, double* rho) {
    rho[0] = dist2(e->cws[0], e->cws[1]); rho[1] = dist2(e->cws[0], e->cws[2]); rho[2] = dist2(e->cws[0], e->cws[3]);
    rho[3] = dist2(e->cws[1], e->cws[2]); rho[4] = dist2(e->cws[1], e->cws[3]); rho[5] = dist2(e->cws[2], e->cws[3]);
}

/* betas10 = [B11 B12 B22 B13 B23 B33 B14 B24 B34 B44]; approx_1 solves for [B11 B12 B13 B14] */
static void find_betas_approx_1(const double* L, const double* rho, double* betas) {
    double l[6 * 4], b4[4];
    for (int i = 0; i < 6; i++) {
        l[i * 4] = L[i * 10]; l[i * 4 + 1] = L[i * 10 + 1]; l[i * 4 + 2] = L[i * 10 + 3]; l[i * 4 + 3] = L[i * 10 + 6];
    }
    solve_svd(l, 6, 4, rho, b4);
    if (b4[0] < 0) {
        betas[0] = sqrt(-b4[0]); betas[1] = -b4[1] / betas[0]; betas[2] = -b4[2] / betas[0]; betas[3] = -b4[3] / betas[0];
    } else {
        betas[0] = sqrt(b4[0]); betas[1] = b4[1] / betas[0]; betas[2] = b4[2] / betas[0]; betas[3] = b4[3] / betas[0];
    }
}
/* approx_2 solves for [B11 B12 B22] */
static void find_betas_approx_2(const double* L, const double* rho, double* betas) {
    double l[6 * 3], b3[3];
    for (int i = 0; i < 6; i++) { l[i * 3] = L[i * 10]; l[i * 3 + 1] = L[i * 10 + 1]; l[i * 3 + 2] = L[i * 10 + 2]; }
    solve_svd(l, 6, 3, rho, b3);
    if (b3[0] < 0) {
        betas[0] = sqrt(-b3[0]); betas[1] = (b3[2] < 0) ? sqrt(-b3[2]) : 0.0;
    } else {
        betas[0] = sqrt(b3[0]); betas[1] = (b3[2] > 0) ? sqrt(b3[2]) : 0.0;
    }
    if (b3[1] < 0) betas[0] = -betas[0];
    betas[2] = 0.0; betas[3] = 0.0;
}
/* approx_3 solves for [B11 B12 B22 B13 B23] */
static void find_betas_approx_3(const double* L, const double* rho, double* betas) {
    double l[6 * 5], b5[5];
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 5; j++) l[i * 5 + j] = L[i * 10 + j];
    solve_svd(l, 6, 5, rho, b5);
    if (b5[0] < 0) {
        betas[0] = sqrt(-b5[0]); betas[1] = (b5[2] < 0) ? sqrt(-b5[2]) : 0.0;
    } else {
        betas[0] = sqrt(b5[0]); betas[1] = (b5[2] > 0) ? sqrt(b5[2]) : 0.0;
    }
    if (b5[1] < 0) betas[0] = -betas[0];
    betas[2] = b5[3] / betas[0];
    betas[3] = 0.0;
}

static void compute_A_and_b_gauss_newton(const double* l_6x10, const double* rho, const double betas[4], double* A, double* b) {
    for (int i = 0; i < 6; i++) {
        const double* rowL = l_6x10 + i * 10;
        double* rowA = A + i * 4;
        rowA[0] = 2 * rowL[0] * betas[0] + rowL[1] * betas[1] + rowL[3] * betas[2] + rowL[6] * betas[3];
        rowA[1] = rowL[1] * betas[0] + 2 * rowL[2] * betas[1] + rowL[4] * betas[2] + rowL[7] * betas[3];
        rowA[2] = rowL[3] * betas[0] + rowL[4] * betas[1] + 2 * rowL[5] * betas[2] + rowL[8] * betas[3];
        rowA[3] = rowL[6] * betas[0] + rowL[7] * betas[1] + rowL[8] * betas[2] + 2 * rowL[9] * betas[3];
        b[i] = rho[i] - (rowL[0] * betas[0] * betas[0] + rowL[1] * betas[0] * betas[1] + rowL[2] * betas[1] * betas[1] +
                         rowL[3] * betas[0] * betas[2] + rowL[4] * betas[1] * betas[2] + rowL[5] * betas[2] * betas[2] +
                         rowL[6] * betas[0] * betas[3] + rowL[7] * betas[1] * betas[3] + rowL[8] * betas[2] * betas[3] +
                         rowL[9] * betas[3] * betas[3]);
    }
}

/* epnp::qr_solve: Householder QR of the 6x4 system (A is overwritten) */
static void qr_solve(double* A, double* b, double* X) {
    const int nr = 6, nc = 4;
    double A1[4], A2[4];
    double *pA = A, *ppAkk = pA;
    for (int k = 0; k < nc; k++) {
        double *ppAik1 = ppAkk, eta = fabs(*ppAik1);
        for (int i = k + 1; i < nr; i++) {
            double elt = fabs(*ppAik1);
            if (eta < elt) eta = elt;
            ppAik1 += nc;
        }
        if (eta == 0) {
            A1[k] = A2[k] = 0.0;
            return; /* "God damnit, A is singular, this shouldn't happen." */
        } else {
            double *ppAik2 = ppAkk, sum2 = 0.0, inv_eta = 1. / eta;
            for (int i = k; i < nr; i++) {
                *ppAik2 *= inv_eta;
                sum2 += *ppAik2 * *ppAik2;
                ppAik2 += nc;
            }
            double sigma = sqrt(sum2);
            if (*ppAkk < 0) sigma = -sigma;
            *ppAkk += sigma;
            A1[k] = sigma * *ppAkk;
            A2[k] = -eta * sigma;
            for (int j = k + 1; j < nc; j++) {
                double *ppAik = ppAkk, sum = 0;
                for (int i = k; i < nr; i++) {
                    sum += *ppAik * ppAik[j - k];
                    ppAik += nc;
                }
                double tau = sum / A1[k];
                ppAik = ppAkk;
                for (int i = k; i < nr; i++) {
                    ppAik[j - k] -= tau * *ppAik;
                    ppAik += nc;
                }
            }
        }
        ppAkk += nc + 1;
    }
    /* b <- Qt b */
    double *ppAjj = pA, *pb = b;
    for (int j = 0; j < nc; j++) {
        double *ppAij = ppAjj, tau = 0;
        for (int i = j; i < nr; i++) {
            tau += *ppAij * pb[i];
            ppAij += nc;
        }
        tau /= A1[j];
        ppAij = ppAjj;
        for (int i = j; i < nr; i++) {
            pb[i] -= tau * *ppAij;
            ppAij += nc;
        }
        ppAjj += nc + 1;
    }
    /* X = R-1 b */
    double* pX = X;
    pX[nc - 1] = pb[nc - 1] / A2[nc - 1];
    for (int i = nc - 2; i >= 0; i--) {
        double *ppAij = pA + i * nc + (i + 1), sum = 0;
        for (int j = i + 1; j < nc; j++) {
            sum += *ppAij * pX[j];
            ppAij++;
        }
        pX[i] = (pb[i] - sum) / A2[i];
    }
}

static void gauss_newton(const double* L, const double* rho, double betas[4]) {
    double a[6 * 4], b[6], x[4] = {0, 0, 0, 0};
    for (int k = 0; k < 5; k++) {
        compute_A_and_b_gauss_newton(L, rho, betas, a, b);
        qr_solve(a, b, x);
        for (int i = 0; i < 4; i++) betas[i] += x[i];
    }
}

/* epnp::compute_pose for n correspondences already loaded into e (pws, us); dbg (optional, 12*12 + 12 doubles)
 * receives Ut and the singular values */
static void epnp_compute_pose(epnp_t* e, double Rout[9], double tout[3], double* dbg) {
    choose_control_points(e);
    compute_barycentric_coordinates(e);
    double M[2 * EPNP_MAXN * 12];
    for (int i = 0; i < e->n; i++) fill_M(e, M, 2 * i, e->alphas + 4 * i, e->us[2 * i], e->us[2 * i + 1]);
    double mtm[144], d[12], ut[144], vt[144];
    mul_transposed(M, 2 * e->n, 12, mtm);
    for (int i = 0; i < 12; i++)
        for (int k = 0; k < 12; k++) ut[i * 12 + k] = mtm[k * 12 + i];
    jacobi_svd(ut, 12, d, vt, 12, 12, 12, 12);
    if (dbg) { memcpy(dbg, ut, sizeof(ut)); memcpy(dbg + 144, d, sizeof(d)); }
    double l_6x10[60], rho[6];
    compute_L_6x10(ut, l_6x10);
    compute_rho(e, rho);
    double Betas[4][4], rep_errors[4], Rs[4][3][3], ts[4][3];
    find_betas_approx_1(l_6x10, rho, Betas[1]);
    gauss_newton(l_6x10, rho, Betas[1]);
    rep_errors[1] = compute_R_and_t(e, ut, Betas[1], Rs[1], ts[1]);
    find_betas_approx_2(l_6x10, rho, Betas[2]);
    gauss_newton(l_6x10, rho, Betas[2]);
    rep_errors[2] = compute_R_and_t(e, ut, Betas[2], Rs[2], ts[2]);
    find_betas_approx_3(l_6x10, rho, Betas[3]);
    gauss_newton(l_6x10, rho, Betas[3]);
    rep_errors[3] = compute_R_and_t(e, ut, Betas[3], Rs[3], ts[3]);
    int N = 1;
    if (rep_errors[2] < rep_errors[1]) N = 2;
    if (rep_errors[3] < rep_errors[N]) N = 3;
    memcpy(Rout, Rs[N], 9 * sizeof(double));
    memcpy(tout, ts[N], 3 * sizeof(double));
}

/* ------------------------------------------------ cv::Rodrigues ------------------------------------------------ */
/* matrix -> vector, as cv::Rodrigues (calib3d/src/calibration.cpp): R is first replaced by its nearest rotation
 * U V^T (Jacobi SVD), then theta = acos((trace-1)/2) and the axis from the skew part */
void pnp_oracle_rodrigues_to_vec(const double* Rin, double* r) {
    double w[3], Ut[9], Vt[9], R[9];
    pnp_oracle_svd(Rin, 3, 3, w, Ut, Vt);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R[i * 3 + j] = Ut[0 * 3 + i] * Vt[0 * 3 + j] + Ut[1 * 3 + i] * Vt[1 * 3 + j] + Ut[2 * 3 + i] * Vt[2 * 3 + j];
    double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (R[0] + R[4] + R[8] - 1) * 0.5;
    c = c > 1. ? 1. : c < -1. ? -1. : c;
    double theta = acos(c);
    if (s < 1e-5) {
        double t;
        if (c > 0) { r[0] = r[1] = r[2] = 0; }
        else {
            t = (R[0] + 1) * 0.5; rx = sqrt(t > 0. ? t : 0.);
            t = (R[4] + 1) * 0.5; ry = sqrt(t > 0. ? t : 0.) * (R[1] < 0 ? -1. : 1.);
            t = (R[8] + 1) * 0.5; rz = sqrt(t > 0. ? t : 0.) * (R[2] < 0 ? -1. : 1.);
            if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
            theta /= sqrt(rx * rx + ry * ry + rz * rz);
            r[0] = rx * theta; r[1] = ry * theta; r[2] = rz * theta;
        }
    } else {
        double vth = 1 / (2 * s);
        vth *= theta;
        r[0] = rx * vth; r[1] = ry * vth; r[2] = rz * vth;
    }
}

/* vector -> matrix */
void pnp_oracle_rodrigues_to_mat(const double* r, double* R) {
    double theta = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (theta < DBL_EPSILON) {
        for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1 : 0;
        return;
    }
    double c = cos(theta), s = sin(theta), c1 = 1. - c, itheta = theta ? 1. / theta : 0.;
    double rx = r[0] * itheta, ry = r[1] * itheta, rz = r[2] * itheta;
    double rrt[9] = {rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz};
    double r_x[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};
    for (int k = 0; k < 9; k++) R[k] = c * ((k % 4 == 0) ? 1 : 0) + c1 * rrt[k] + s * r_x[k];
}

/* ---------------------------- solvePnP(SOLVEPNP_EPNP) on a (sub)set: the RANSAC kernel ------------------------ */
/* obj: n x 3 float32, img: n x 2 float32 pixels, K row-major 3x3.  undistortPoints (no distortion) stores the
 * normalised coordinates as float32; epnp's constructor maps them back with the camera matrix in double. */
void pnp_oracle_epnp(const float* obj, const float* img, int n, const double* K, double* rvec, double* tvec, double* Rout,
                     double* dbg) {
    epnp_t e;
    e.fu = K[0]; e.fv = K[4]; e.uc = K[2]; e.vc = K[5];
    e.n = n;
    const double ifx = 1. / K[0], ify = 1. / K[4];
    for (int i = 0; i < n; i++) {
        e.pws[3 * i] = obj[3 * i]; e.pws[3 * i + 1] = obj[3 * i + 1]; e.pws[3 * i + 2] = obj[3 * i + 2];
        const float xn = (float)(((double)img[2 * i] - K[2]) * ifx), yn = (float)(((double)img[2 * i + 1] - K[5]) * ify);
        e.us[2 * i] = xn * e.fu + e.uc;
        e.us[2 * i + 1] = yn * e.fv + e.vc;
    }
    double R[9], t[3];
    epnp_compute_pose(&e, R, t, dbg);
    if (Rout) memcpy(Rout, R, sizeof(R));
    pnp_oracle_rodrigues_to_vec(R, rvec);
    memcpy(tvec, t, sizeof(t));
}

/* PnPRansacCallback::computeError: cv::projectPoints into float32, squared float32 distance */
void pnp_oracle_errors(const float* obj, const float* img, int n, const double* K, const double* rvec, const double* tvec,
                       float* err) {
    double R[9];
    pnp_oracle_rodrigues_to_mat(rvec, R);
    const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    for (int i = 0; i < n; i++) {
        const double X = obj[3 * i], Y = obj[3 * i + 1], Z = obj[3 * i + 2];
        double x = R[0] * X + R[1] * Y + R[2] * Z + tvec[0];
        double y = R[3] * X + R[4] * Y + R[5] * Z + tvec[1];
        double z = R[6] * X + R[7] * Y + R[8] * Z + tvec[2];
        z = z ? 1. / z : 1;
        x *= z; y *= z;
        const float pu = (float)(x * fx + cx), pv = (float)(y * fy + cy);
        const float dx = img[2 * i] - pu, dy = img[2 * i + 1] - pv;
        err[i] = dx * dx + dy * dy;
    }
}

static int ransac_update_num_iters(double p, double ep, int modelPoints, int maxIters) {
    p = p > 0. ? p : 0.; p = p < 1. ? p : 1.;
    ep = ep > 0. ? ep : 0.; ep = ep < 1. ? ep : 1.;
    double num = 1. - p > DBL_MIN ? 1. - p : DBL_MIN;
    double denom = 1. - pow(1. - ep, modelPoints);
    if (denom < DBL_MIN) return 0;
    num = log(num);
    denom = log(denom);
    return denom >= 0 || -num >= maxIters * (-denom) ? maxIters : (int)lrint(num / denom);
}

/* RANSACPointSetRegistrator::run with the PnP callback.  Outputs: mask[n] of the best model, its (rvec, tvec),
 * trace (optional, cap rows of 8): per executed iteration {5 sample indices, goodCount, niters after, accepted}.
 * Returns the number of executed iterations (0 when no model gathered more than 4 inliers -> result false). */
int pnp_oracle_ransac(const float* obj, const float* img, int n, const double* K, int max_iters, double threshold,
                      double confidence, uint8_t* best_mask, double* best_rvec, double* best_tvec, int* n_good,
                      int* trace, int trace_cap) {
    const int modelPoints = 5;
    int niters = max_iters > 1 ? max_iters : 1, maxGoodCount = 0, iter;
    cv_rng rng = {(uint64_t)-1};
    float* err = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    uint8_t* mask = (uint8_t*)malloc((size_t)(n > 0 ? n : 1));
    const float thresh = (float)(threshold * threshold);
    *n_good = 0;
    if (n < modelPoints) { free(err); free(mask); return 0; }
    for (iter = 0; iter < niters; iter++) {
        int idx[5];
        float so[15], si[10];
        for (int i = 0; i < modelPoints; ++i) { /* getSubset (checkSubset of the PnP callback accepts everything) */
            int idx_i;
            for (;;) {
                idx_i = rng_uniform(&rng, 0, n);
                int dup = 0;
                for (int q = 0; q < i; ++q) dup |= idx[q] == idx_i;
                if (!dup) break;
            }
            idx[i] = idx_i;
            memcpy(so + 3 * i, obj + 3 * idx_i, 12);
            memcpy(si + 2 * i, img + 2 * idx_i, 8);
        }
        double rvec[3], tvec[3];
        pnp_oracle_epnp(so, si, modelPoints, K, rvec, tvec, NULL, NULL);
        pnp_oracle_errors(obj, img, n, K, rvec, tvec, err);
        int goodCount = 0;
        for (int i = 0; i < n; i++) { mask[i] = err[i] <= thresh; goodCount += mask[i]; }
        int accepted = 0;
        if (goodCount > (maxGoodCount > modelPoints - 1 ? maxGoodCount : modelPoints - 1)) {
            memcpy(best_mask, mask, (size_t)n);
            memcpy(best_rvec, rvec, sizeof(rvec));
            memcpy(best_tvec, tvec, sizeof(tvec));
            maxGoodCount = goodCount;
            niters = ransac_update_num_iters(confidence, (double)(n - goodCount) / n, modelPoints, niters);
            accepted = 1;
        }
        if (trace && iter < trace_cap) {
            int* t = trace + 8 * iter;
            for (int q = 0; q < 5; ++q) t[q] = idx[q];
            t[5] = goodCount; t[6] = niters; t[7] = accepted;
        }
    }
    free(err); free(mask);
    *n_good = maxGoodCount;
    return iter;
}

/* ---------------------- refit of the inliers: solvePnP(SOLVEPNP_ITERATIVE, no extrinsic guess) ----------------- */
/* OpenCV initialises with a DLT (non-planar) or a homography (planar) and runs Levenberg-Marquardt on (rvec, tvec)
 * to max 20 iterations / FLT_EPSILON; the result is the least-squares optimum of the reprojection error over the
 * inliers (5e-15 relative from a perturbed start).  Restated as Gauss-Newton with a left SE3 update
 * from the RANSAC model, iterated to convergence: same optimum, compared with cv2 at 1e-6. */
static void se3_exp_left(double* T, const double* xi) {
    const double w0 = xi[3], w1 = xi[4], w2 = xi[5];
    const double th2 = w0 * w0 + w1 * w1 + w2 * w2, th = sqrt(th2);
    double A, B, Cc;
    if (th < 1e-10) { A = 1.0 - th2 / 6.0; B = 0.5 - th2 / 24.0; Cc = 1.0 / 6.0 - th2 / 120.0; }
    else { A = sin(th) / th; B = (1 - cos(th)) / th2; Cc = (th - sin(th)) / (th2 * th); }
    const double W[9] = {0, -w2, w1, w2, 0, -w0, -w1, w0, 0};
    double W2[9], dR[9], V[9], Tn[12];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) W2[i * 3 + j] = W[i * 3] * W[j] + W[i * 3 + 1] * W[3 + j] + W[i * 3 + 2] * W[6 + j];
    for (int i = 0; i < 9; ++i) { dR[i] = A * W[i] + B * W2[i]; V[i] = B * W[i] + Cc * W2[i]; }
    dR[0] += 1; dR[4] += 1; dR[8] += 1; V[0] += 1; V[4] += 1; V[8] += 1;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) Tn[r * 4 + c] = dR[r * 3] * T[c] + dR[r * 3 + 1] * T[4 + c] + dR[r * 3 + 2] * T[8 + c];
        Tn[r * 4 + 3] = dR[r * 3] * T[3] + dR[r * 3 + 1] * T[7] + dR[r * 3 + 2] * T[11] + V[r * 3] * xi[0] + V[r * 3 + 1] * xi[1] + V[r * 3 + 2] * xi[2];
    }
    memcpy(T, Tn, sizeof(Tn));
}

static int chol_solve6(double* H, double* g) {
    for (int j = 0; j < 6; ++j) {
        double d = H[j * 6 + j];
        for (int k = 0; k < j; ++k) d -= H[j * 6 + k] * H[j * 6 + k];
        if (!(d > 0)) return 0;
        d = sqrt(d);
        H[j * 6 + j] = d;
        for (int i = j + 1; i < 6; ++i) {
            double s = H[i * 6 + j];
            for (int k = 0; k < j; ++k) s -= H[i * 6 + k] * H[j * 6 + k];
            H[i * 6 + j] = s / d;
        }
    }
    for (int i = 0; i < 6; ++i) { double s = g[i]; for (int k = 0; k < i; ++k) s -= H[i * 6 + k] * g[k]; g[i] = s / H[i * 6 + i]; }
    for (int i = 5; i >= 0; --i) { double s = g[i]; for (int k = i + 1; k < 6; ++k) s -= H[k * 6 + i] * g[k]; g[i] = s / H[i * 6 + i]; }
    return 1;
}

void pnp_oracle_refit(const float* obj, const float* img, int n, const uint8_t* mask, const double* K, double* rvec,
                      double* tvec, double* T_out) {
    double R[9], T[12];
    pnp_oracle_rodrigues_to_mat(rvec, R);
    for (int r = 0; r < 3; ++r) { T[r * 4] = R[r * 3]; T[r * 4 + 1] = R[r * 3 + 1]; T[r * 4 + 2] = R[r * 3 + 2]; T[r * 4 + 3] = tvec[r]; }
    const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    for (int it = 0; it < 50; ++it) {
        double H[36] = {0}, g[6] = {0};
        for (int i = 0; i < n; ++i) {
            if (!mask[i]) continue;
            const double X = obj[3 * i], Y = obj[3 * i + 1], Z = obj[3 * i + 2];
            const double px = T[0] * X + T[1] * Y + T[2] * Z + T[3], py = T[4] * X + T[5] * Y + T[6] * Z + T[7],
                         pz = T[8] * X + T[9] * Y + T[10] * Z + T[11];
            const double iz = 1.0 / pz, iz2 = iz * iz;
            const double e0 = (double)img[2 * i] - (fx * px * iz + cx), e1 = (double)img[2 * i + 1] - (fy * py * iz + cy);
            const double J0[6] = {-fx * iz, 0, fx * px * iz2, fx * px * py * iz2, -fx - fx * px * px * iz2, fx * py * iz};
            const double J1[6] = {0, -fy * iz, fy * py * iz2, fy + fy * py * py * iz2, -fy * px * py * iz2, -fy * px * iz};
            for (int a = 0; a < 6; ++a) {
                g[a] -= J0[a] * e0 + J1[a] * e1;
                for (int b = 0; b < 6; ++b) H[a * 6 + b] += J0[a] * J0[b] + J1[a] * J1[b];
            }
        }
        if (!chol_solve6(H, g)) break;
        se3_exp_left(T, g);
        double nx = 0;
        for (int a = 0; a < 6; ++a) nx += g[a] * g[a];
        if (nx < 1e-28) break;
    }
    for (int r = 0; r < 3; ++r) { R[r * 3] = T[r * 4]; R[r * 3 + 1] = T[r * 4 + 1]; R[r * 3 + 2] = T[r * 4 + 2]; tvec[r] = T[r * 4 + 3]; }
    pnp_oracle_rodrigues_to_vec(R, rvec);
    if (T_out) memcpy(T_out, T, sizeof(T));
}

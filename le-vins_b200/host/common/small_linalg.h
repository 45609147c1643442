// Two dense kernels the host mirror needs without Eigen: the 5x3 column-pivoting Householder QR solve that
// Eigen::ColPivHouseholderQR performs at tracking.cc:584-589 / pointcloud.cc:89 (Eigen >= 3.3 algorithm with
// LAPACK-style norm down-dating), and a symmetric 3x3 eigen-decomposition (cyclic Jacobi) standing in for the
// Eigen::EigenSolver call at pointcloud.cc:132-134 (the matrix there is a covariance, i.e. symmetric).
#ifndef LEVINS_HOST_SMALL_LINALG_H
#define LEVINS_HOST_SMALL_LINALG_H

#include <cfloat>
#include <cmath>

namespace levins_host {

// A: row-major 5x3, b: 5, x: 3
inline void colpiv_qr_solve_5x3(const double A[15], const double b[5], double x[3]) {
    enum { ROWS = 5, COLS = 3 };
    double q[COLS][ROWS];
    for (int r = 0; r < ROWS; r++)
        for (int c = 0; c < COLS; c++) q[c][r] = A[r * COLS + c];
    double hcoef[COLS], nu[COLS], nd[COLS];
    int perm[COLS] = {0, 1, 2};
    for (int k = 0; k < COLS; k++) {
        double s = 0;
        for (int r = 0; r < ROWS; r++) s += q[k][r] * q[k][r];
        nd[k] = nu[k] = std::sqrt(s);
    }
    double maxn = nu[0];
    for (int k = 1; k < COLS; k++)
        if (nu[k] > maxn) maxn = nu[k];
    const double th = maxn * DBL_EPSILON / (double) ROWS;
    const double threshold_helper = th * th, downdate = std::sqrt(DBL_EPSILON);
    int nonzero = COLS;
    for (int k = 0; k < COLS; k++) {
        int big = k;
        for (int j = k + 1; j < COLS; j++)
            if (nu[j] > nu[big]) big = j;
        if (nonzero == COLS && nu[big] * nu[big] < threshold_helper * (double) (ROWS - k)) nonzero = k;
        if (big != k) {
            for (int r = 0; r < ROWS; r++) std::swap(q[k][r], q[big][r]);
            std::swap(nu[k], nu[big]);
            std::swap(nd[k], nd[big]);
            std::swap(perm[k], perm[big]);
        }
        const int len = ROWS - k;
        double *col = &q[k][k];
        double tail_sq = 0;
        for (int r = 1; r < len; r++) tail_sq += col[r] * col[r];
        const double c0 = col[0];
        double beta, tau;
        if (tail_sq <= DBL_MIN) {
            tau = 0;
            beta = c0;
            for (int r = 1; r < len; r++) col[r] = 0;
        } else {
            beta = std::sqrt(c0 * c0 + tail_sq);
            if (c0 >= 0) beta = -beta;
            for (int r = 1; r < len; r++) col[r] = col[r] / (c0 - beta);
            tau = (beta - c0) / beta;
        }
        hcoef[k] = tau;
        col[0] = beta;
        if (tau != 0)
            for (int j = k + 1; j < COLS; j++) {
                double *cj = &q[j][k];
                double tmp = 0;
                for (int r = 1; r < len; r++) tmp += col[r] * cj[r];
                tmp += cj[0];
                cj[0] -= tau * tmp;
                for (int r = 1; r < len; r++) cj[r] -= (tau * col[r]) * tmp;
            }
        for (int j = k + 1; j < COLS; j++)
            if (nu[j] != 0) {
                double temp = std::fabs(q[j][k]) / nu[j];
                temp = (1.0 + temp) * (1.0 - temp);
                temp = temp < 0 ? 0 : temp;
                const double ratio = nu[j] / nd[j];
                if (temp * (ratio * ratio) <= downdate) {
                    double s = 0;
                    for (int r = k + 1; r < ROWS; r++) s += q[j][r] * q[j][r];
                    nd[j] = nu[j] = std::sqrt(s);
                } else {
                    nu[j] *= std::sqrt(temp);
                }
            }
    }
    double c[ROWS];
    for (int r = 0; r < ROWS; r++) c[r] = b[r];
    for (int k = 0; k < nonzero; k++) {
        const double tau = hcoef[k];
        const int len = ROWS - k;
        const double *ess = &q[k][k];
        if (tau != 0) {
            double tmp = 0;
            for (int r = 1; r < len; r++) tmp += ess[r] * c[k + r];
            tmp += c[k];
            c[k] -= tau * tmp;
            for (int r = 1; r < len; r++) c[k + r] -= (tau * ess[r]) * tmp;
        }
    }
    for (int i = nonzero - 1; i >= 0; i--) {
        c[i] /= q[i][i];
        for (int r = 0; r < i; r++) c[r] -= c[i] * q[i][r];
    }
    x[0] = x[1] = x[2] = 0;
    for (int i = 0; i < nonzero; i++) x[perm[i]] = c[i];
}

// symmetric 3x3 (row-major a[9]) -> eigenvalues w[3], eigenvectors as columns of v (row-major v[9]); cyclic Jacobi
inline void sym_eigen_3x3(const double a_in[9], double w[3], double v[9]) {
    double a[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) a[i][j] = a_in[3 * i + j];
    for (int sweep = 0; sweep < 60; sweep++) {
        const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        if (off < 1e-300) break;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                if (a[p][q] == 0.0) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double cs = 1.0 / std::sqrt(t * t + 1.0), sn = t * cs;
                for (int k = 0; k < 3; k++) {
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = cs * akp - sn * akq;
                    a[k][q] = sn * akp + cs * akq;
                }
                for (int k = 0; k < 3; k++) {
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = cs * apk - sn * aqk;
                    a[q][k] = sn * apk + cs * aqk;
                }
                for (int k = 0; k < 3; k++) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = cs * vkp - sn * vkq;
                    V[k][q] = sn * vkp + cs * vkq;
                }
            }
    }
    for (int i = 0; i < 3; i++) {
        w[i] = a[i][i];
        for (int j = 0; j < 3; j++) v[3 * i + j] = V[i][j];
    }
}

} // namespace levins_host
#endif

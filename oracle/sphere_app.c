/*
 * sphere_app.c -- CPU restatement of the shallow-water-on-the-sphere application helpers
 * (TEST INFRASTRUCTURE, part of the oracle): apps/shallow-sphere/mapc2p.f:2-76,
 * setaux.f:2-218, qinit.f:3-107, src2.f:2-147 under /root/reference.
 * Arrays are Fortran ordered, aux(16, 1-mbc:mx+mbc, 1-mbc:my+mbc), q(4, ...).
 */
#include <math.h>
#include <stdlib.h>

static double dmax(double a, double b) { return a > b ? a : b; }
static double dsign1(double x) { return signbit(x) ? -1.0 : 1.0; }

/* mapc2p.f:2-76 */
void oracle_sphere_mapc2p(double x1, double y1, double *xp, double *yp, double *zp, double Rsphere)
{
    double r1 = Rsphere, xc = x1, yc = y1, sgnz;
    if (xc >= 1.0) xc = xc - 4.0;
    if (xc < -3.0) xc = xc + 4.0;
    if (yc >= 1.0) { yc = 2.0 - yc; xc = -2.0 - xc; }
    if (yc < -1.0) { yc = -2.0 - yc; xc = -2.0 - xc; }
    if (xc < -1.0) { xc = -2.0 - xc; sgnz = -1.0; } else sgnz = 1.0;
    double sgnxc = dsign1(xc), sgnyc = dsign1(yc);
    double xc1 = fabs(xc), yc1 = fabs(yc);
    double d = dmax(dmax(xc1, yc1), 1.e-10);
    double DD = r1 * d * (2.0 - d) / sqrt(2.0);
    double R = r1;
    double center = DD - sqrt(dmax(R * R - DD * DD, 0.0));
    double x = DD / d * xc1, y = DD / d * yc1;
    if (yc1 > xc1) y = center + sqrt(dmax(R * R - x * x, 0.0));
    else x = center + sqrt(dmax(R * R - y * y, 0.0));
    double z = sqrt(dmax(r1 * r1 - (x * x + y * y), 0.0));
    *xp = x * sgnxc; *yp = y * sgnyc; *zp = z * sgnz;
}

#define AUX(ma, i, j) aux[((ma)-1) + 16 * (((i) + mbc - 1) + (size_t)nx * ((j) + mbc - 1))]

/* setaux.f:2-218 */
void oracle_sphere_setaux(int mbc, int mx, int my, double xlower, double ylower, double dxc, double dyc,
                          double *aux, double Rsphere)
{
    const double pi = 4.0 * atan(1.0);
    int nx = mx + 2 * mbc;
    int n1 = mx + 2 * mbc + 1, n2 = my + 2 * mbc + 1;
    double *xp = (double *)malloc(sizeof(double) * n1 * n2), *yp = (double *)malloc(sizeof(double) * n1 * n2);
    double *zp = (double *)malloc(sizeof(double) * n1 * n2), *th = (double *)calloc(n1 * n2, sizeof(double));
    double *ph = (double *)malloc(sizeof(double) * n1 * n2);
#define C2(arr, i, j) arr[((i) + mbc - 1) + (size_t)n1 * ((j) + mbc - 1)]
    for (int j = 1 - mbc; j <= my + mbc + 1; j++)
        for (int i = 1 - mbc; i <= mx + mbc + 1; i++) {
            double xc = xlower + (i - 1.0) * dxc, yc = ylower + (j - 1.0) * dyc;
            oracle_sphere_mapc2p(xc, yc, &C2(xp, i, j), &C2(yp, i, j), &C2(zp, i, j), Rsphere);
            double r = sqrt(C2(xp, i, j) * C2(xp, i, j) + C2(yp, i, j) * C2(yp, i, j));
            if (r > 1.e-4) C2(th, i, j) = acos(C2(xp, i, j) / r);
            else if (C2(yp, i, j) > 0.0) C2(th, i, j) = 0.0;
            if (C2(yp, i, j) < 0.0) C2(th, i, j) = -C2(th, i, j);
            if (C2(zp, i, j) > 0.0) C2(ph, i, j) = pi / 2.0 - acos(r / Rsphere);
            else C2(ph, i, j) = pi / 2.0 + acos(r / Rsphere);
        }
    for (int j = 1 - mbc; j <= my + mbc; j++)
        for (int i = 1 - mbc; i <= mx + mbc; i++) {
            double etx = C2(xp, i, j + 1) - C2(xp, i, j), ety = C2(yp, i, j + 1) - C2(yp, i, j),
                   etz = C2(zp, i, j + 1) - C2(zp, i, j);
            AUX(5, i, j) = etx; AUX(6, i, j) = ety; AUX(7, i, j) = etz;
            double erx = 0.5 * (C2(xp, i, j) + C2(xp, i, j + 1)), ery = 0.5 * (C2(yp, i, j) + C2(yp, i, j + 1)),
                   erz = 0.5 * (C2(zp, i, j) + C2(zp, i, j + 1));
            double enx = ety * erz - etz * ery, eny = etz * erx - etx * erz, enz = etx * ery - ety * erx;
            double ennorm = sqrt(enx * enx + eny * eny + enz * enz);
            AUX(2, i, j) = enx / ennorm; AUX(3, i, j) = eny / ennorm; AUX(4, i, j) = enz / ennorm;
            etx = C2(xp, i + 1, j) - C2(xp, i, j); ety = C2(yp, i + 1, j) - C2(yp, i, j);
            etz = C2(zp, i + 1, j) - C2(zp, i, j);
            AUX(11, i, j) = etx; AUX(12, i, j) = ety; AUX(13, i, j) = etz;
            erx = 0.5 * (C2(xp, i, j) + C2(xp, i + 1, j)); ery = 0.5 * (C2(yp, i, j) + C2(yp, i + 1, j));
            erz = 0.5 * (C2(zp, i, j) + C2(zp, i + 1, j));
            enx = ery * etz - erz * ety; eny = erz * etx - erx * etz; enz = erx * ety - ery * etx;
            ennorm = sqrt(enx * enx + eny * eny + enz * enz);
            AUX(8, i, j) = enx / ennorm; AUX(9, i, j) = eny / ennorm; AUX(10, i, j) = enz / ennorm;
            /* "(i-0.5)" is a REAL(4) literal in setaux.f:161-162, exactly representable */
            double xcm = xlower + (i - 0.5) * dxc, ycm = ylower + (j - 0.5) * dyc, xpm, ypm, zpm;
            oracle_sphere_mapc2p(xcm, ycm, &xpm, &ypm, &zpm, Rsphere);
            AUX(14, i, j) = xpm; AUX(15, i, j) = ypm; AUX(16, i, j) = zpm;
#define BETA(pa, ta, pb, tb) (sin(pa) * sin(pb) * cos((ta) - (tb)) + cos(pa) * cos(pb))
            double beta12 = BETA(C2(ph, i, j), C2(th, i, j), C2(ph, i + 1, j), C2(th, i + 1, j));
            double beta23 = BETA(C2(ph, i, j + 1), C2(th, i, j + 1), C2(ph, i + 1, j), C2(th, i + 1, j));
            double beta13 = BETA(C2(ph, i, j + 1), C2(th, i, j + 1), C2(ph, i, j), C2(th, i, j));
            double beta24 = BETA(C2(ph, i + 1, j + 1), C2(th, i + 1, j + 1), C2(ph, i + 1, j), C2(th, i + 1, j));
            double beta34 = BETA(C2(ph, i + 1, j + 1), C2(th, i + 1, j + 1), C2(ph, i, j + 1), C2(th, i, j + 1));
            double d12 = Rsphere * acos(beta12), d23 = Rsphere * acos(beta23), d13 = Rsphere * acos(beta13);
            double d24 = Rsphere * acos(beta24), d34 = Rsphere * acos(beta34);
            double s123 = 0.5 * (d12 + d23 + d13), s234 = 0.5 * (d23 + d34 + d24);
            double t123 = tan(s123 / 2.0) * tan((s123 - d12) / 2.0) * tan((s123 - d23) / 2.0) * tan((s123 - d13) / 2.0);
            t123 = dmax(t123, 0.0);
            double E123 = 4.0 * atan(sqrt(t123));
            double t234 = tan(s234 / 2.0) * tan((s234 - d23) / 2.0) * tan((s234 - d34) / 2.0) * tan((s234 - d24) / 2.0);
            t234 = dmax(t234, 0.0);
            double E234 = 4.0 * atan(sqrt(t234));
            double area = (E123 + E234);
            AUX(1, i, j) = area / (dxc * dyc);
        }
    free(xp); free(yp); free(zp); free(th); free(ph);
}

#define QQ(m, i, j) q[((m)-1) + 4 * (((i) + mbc - 1) + (size_t)nx * ((j) + mbc - 1))]

/* qinit.f:3-107 (4-Rossby-Haurwitz wave) */
void oracle_sphere_qinit(int mbc, int mx, int my, double xlower, double ylower, double dx, double dy,
                         double *q, double Rsphere)
{
    const double pi = 4.0 * atan(1.0);
    const double a = 6.37122e6, K = 7.848e-6, Omega = 7.292e-5, G = 9.80616, t0 = 86400.0, h0 = 8.e3, R = 4.0;
    int nx = mx + 2 * mbc;
    for (int i = 1; i <= mx; i++) {
        double xc = xlower + (i - 0.5) * dx;
        for (int j = 1; j <= my; j++) {
            double yc = ylower + (j - 0.5) * dy, xp, yp, zp, theta = 0.0, phi;
            oracle_sphere_mapc2p(xc, yc, &xp, &yp, &zp, Rsphere);
            double rad = dmax(sqrt(xp * xp + yp * yp), 1.e-6);
            if (xp > 0.0 && yp > 0.0) theta = asin(yp / rad);
            else if (xp < 0.0 && yp > 0.0) theta = pi - asin(yp / rad);
            else if (xp < 0.0 && yp < 0.0) theta = -pi + asin(-yp / rad);
            else if (xp > 0.0 && yp < 0.0) theta = -asin(-yp / rad);
            if (zp > 0.0) phi = asin(zp / Rsphere);
            else phi = -asin(-zp / Rsphere);
            xp = theta; yp = phi;
            double cy = cos(yp);
            double bigA = 0.5 * K * (2.0 * Omega + K) * pow(cy, 2.0) +
                          0.25 * K * K * pow(cy, 2.0 * R) *
                              ((1.0 * R + 1.0) * pow(cy, 2.0) + (2.0 * R * R - 1.0 * R - 2.0) -
                               2.0 * R * R * pow(cy, -2.0));
            double bigB = (2.0 * (Omega + K) * K) / ((1.0 * R + 1.0) * (1.0 * R + 2.0)) * pow(cy, R) *
                          ((1.0 * R * R + 2.0 * R + 2.0) - (1.0 * R + 1.0) * (1.0 * R + 1.0) * (cy * cy));
            double bigC = 0.25 * K * K * pow(cy, 2 * R) * ((1.0 * R + 1.0) * (cy * cy) - (1.0 * R + 2.0));
            double Uin1 = (K * cy + K * pow(cy, R - 1.) * (R * pow(sin(yp), 2.) - pow(cy, 2.)) * cos(R * xp)) * t0;
            double Uin2 = (-K * R * pow(cy, R - 1.) * sin(yp) * sin(R * xp)) * t0;
            double Uout1 = (-sin(xp) * Uin1 - sin(yp) * cos(xp) * Uin2);
            double Uout2 = (cos(xp) * Uin1 - sin(yp) * sin(xp) * Uin2);
            double Uout3 = cos(yp) * Uin2;
            QQ(1, i, j) = h0 / a + (a / G) * (bigA + bigB * cos(R * xp) + bigC * cos(2.0 * R * xp));
            QQ(2, i, j) = QQ(1, i, j) * Uout1;
            QQ(3, i, j) = QQ(1, i, j) * Uout2;
            QQ(4, i, j) = QQ(1, i, j) * Uout3;
        }
    }
}

/* src2.f:2-147 on interior arrays q(4, mx, my), aux(16, mx, my) */
void oracle_sphere_src2(int mx, int my, double xlower, double ylower, double dx, double dy, double *q,
                        const double *aux, double dt, double Rsphere)
{
    const double df = (double)12.600576e0f; /* "12.600576e0" is a REAL(4) literal (src2.f:38) */
#define QI(m, i, j) q[((m)-1) + 4 * (((i)-1) + (size_t)mx * ((j)-1))]
#define AI(ma, i, j) aux[((ma)-1) + 16 * (((i)-1) + (size_t)mx * ((j)-1))]
    for (int i = 1; i <= mx; i++)
        for (int j = 1; j <= my; j++) {
            double erx = AI(14, i, j), ery = AI(15, i, j), erz = AI(16, i, j);
            double qn = erx * QI(2, i, j) + ery * QI(3, i, j) + erz * QI(4, i, j);
            QI(2, i, j) = QI(2, i, j) - qn * erx;
            QI(3, i, j) = QI(3, i, j) - qn * ery;
            QI(4, i, j) = QI(4, i, j) - qn * erz;
        }
    for (int i = 1; i <= mx; i++) {
        double xc = xlower + (i - 0.5) * dx;
        for (int j = 1; j <= my; j++) {
            double yc = ylower + (j - 0.5) * dy, erx, ery, erz;
            oracle_sphere_mapc2p(xc, yc, &erx, &ery, &erz, Rsphere);
            double fcor = df * erz;
            double RK[4][3];
            double hu = QI(2, i, j), hv = QI(3, i, j), hw = QI(4, i, j);
            for (int st = 0; st < 4; st++) {
                if (st > 0) {
                    hu = QI(2, i, j) + 0.5 * RK[st - 1][0];
                    hv = QI(3, i, j) + 0.5 * RK[st - 1][1];
                    hw = QI(4, i, j) + 0.5 * RK[st - 1][2];
                }
                RK[st][0] = fcor * dt * (erz * hv - ery * hw);
                RK[st][1] = dt * fcor * (erx * hw - erz * hu);
                RK[st][2] = dt * fcor * (ery * hu - erx * hv);
            }
            for (int m = 2; m <= 4; m++)
                QI(m, i, j) = QI(m, i, j) +
                              (RK[0][m - 2] + 2.0 * RK[1][m - 2] + 2.0 * RK[2][m - 2] + RK[3][m - 2]) / 6.0;
        }
    }
    for (int i = 1; i <= mx; i++)
        for (int j = 1; j <= my; j++) {
            double erx = AI(14, i, j), ery = AI(15, i, j), erz = AI(16, i, j);
            double qn = erx * QI(2, i, j) + ery * QI(3, i, j) + erz * QI(4, i, j);
            QI(2, i, j) = QI(2, i, j) - qn * erx;
            QI(3, i, j) = QI(3, i, j) - qn * ery;
            QI(4, i, j) = QI(4, i, j) - qn * erz;
        }
}

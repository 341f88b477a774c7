// fused.cuh -- single-pass unsplit step for the light Riemann solvers (sm_100a).
//
// step2.f:84-238 makes two passes over the grid (x-sweeps over every row, then y-sweeps over
// every column), each of which reads qold and read-modify-writes qnew; the two-kernel engines of
// classic.cuh mirror that (184 B of DRAM traffic per cell-step for acoustics against the 48 B
// of "read q once, write it once").  This kernel does both passes in ONE walk: a CTA owns a
// strip of columns (thread t <-> column i0-2+t) and walks down the rows once.  At row r
//   * the x-interfaces of row r are solved between neighbouring threads (shared memory), limited,
//     split transversely; their contributions complete the x-part of row r-1 and open row r's;
//   * the y-interface between rows r-1 and r is solved inside the thread; interface r-1 can now
//     be limited (it needs the waves of r-2, r-1, r) and split transversely, which completes
//     cell row r-2: its x-part (finished one iteration ago, kept in registers) receives the
//     y-contributions in the reference's order:
//         q = X + dtdx G2'(i-1);  q = q + (qadd' - dtdy dF' - dtdx (G2' - G1'));  q = q - dtdx G1'(i+1)
// q is read once (cp.async, double buffered) and written once; every floating-point operation
// and its order are those of the two-pass engines, so results are bit-identical to them and to
// the oracle.  Built for solvers without aux data and without a capacity function whose rolling
// windows fit in registers (acoustics, advection, shallow water); the Euler solver's windows
// (x-engine 254 registers + y-engine 254 registers and 103 doubles of shared memory per thread)
// do not fit one kernel: it keeps the two-pass engines.
#pragma once
#include "classic.cuh"

template <class RPX, class RPY, int NT>
__global__ void __launch_bounds__(NT, rp_f_minb<RPX>::value) fused_step2_kernel(const SweepArgs A)
{
    constexpr int MEQN = RPX::MEQN, MW = RPX::MWAVES, NROE = RPX::NROE;
    constexpr int NC = NT - 4;
    constexpr int QS = NT + 1;
    extern __shared__ double sm[];
    double *qs0 = sm;                   // [2][MEQN][NT+1] staged rows of qold, index k <-> column i0-3+k
    double *ws = qs0 + 2 * MEQN * QS;   // [MEQN*MW][NT]  unlimited x-waves of this row
    double *xs = ws + MEQN * MW * NT;   // [4*MEQN][NT]   amdq, F, bm(A-), bp(A-) of each x-interface
    double *gs = xs + 4 * MEQN * NT;    // [2][2*MEQN][NT] G1', G2' of the y-sweep (double buffered)

    const int t = threadIdx.x;
    const int mbc = A.mbc;
    const int i0 = A.ilo + blockIdx.x * NC;
    const int ic = i0 - 2 + t;                       // this thread's column (Fortran index)
    const int imax = A.mx + mbc;
    const int icl = min(max(ic, 1 - mbc), imax) + mbc - 1;
    const int cload = min(max(i0 - 3 + t, 1 - mbc), imax) + mbc - 1;
    const int cload2 = min(i0 - 3 + NT, imax) + mbc - 1;
    const int j0 = A.jlo + blockIdx.y * A.rows_per_cta;
    const int j1 = min(j0 + A.rows_per_cta, A.jhi + 1);
    const bool col_out = (t >= 2) && (t <= NT - 3) && (ic <= A.ihi);
    const bool xiface_ok = (ic >= 1) && (ic <= A.mx + 1) && (t >= 1) && (t <= NT - 2);
    const bool ycol_cfl = (ic >= 0) && (ic <= A.mx + 1);
    const bool order2 = (A.order != 1);
    const bool trans2 = order2 && (A.trans == 2);
    double dtdx, dtdy;
    load_dt(A, dtdx, dtdy);
    const double hdtdx = 0.5 * dtdx, hdtdy = 0.5 * dtdy;
    const AuxCell nocell{nullptr, 0};

    unsigned long long smaxx = 0ULL, smaxy = 0ULL;   // bits of max|s| per direction
    // x-part window
    double accPrev[MEQN], pendA[MEQN], xd1[MEQN], xd2[MEQN];
    // y-part window (interface r-1 and r-2 of this column)
    double qm1[MEQN], qm2[MEQN], wl[MEQN][MW], sm1[MW], norm1[MW], dot1[MW];
    double am1[MEQN], ap1[MEQN], ap2[MEQN], f2[MEQN], roe1[NROE], bmp2[MEQN], bpp2[MEQN];
#pragma unroll
    for (int m = 0; m < MEQN; m++) {
        accPrev[m] = pendA[m] = xd1[m] = xd2[m] = 0.0;
        qm1[m] = qm2[m] = 1.0;
        am1[m] = ap1[m] = ap2[m] = f2[m] = bmp2[m] = bpp2[m] = 0.0;
#pragma unroll
        for (int mw = 0; mw < MW; mw++) wl[m][mw] = 0.0;
    }
#pragma unroll
    for (int mw = 0; mw < MW; mw++) { sm1[mw] = 0.0; norm1[mw] = 0.0; dot1[mw] = 0.0; }
#pragma unroll
    for (int n = 0; n < NROE; n++) roe1[n] = 1.0;

    // CLAW_Y_PEEL (classic.cuh): row j0-2 goes straight into the y-window, the walk starts at j0-1
    const int rbeg = CLAW_Y_PEEL ? j0 - 1 : j0 - 2, rend = j1 + 1;
    if (CLAW_Y_PEEL) {
#pragma unroll
        for (int m = 0; m < MEQN; m++)
            qm1[m] = __ldg(&A.qin[m * A.mstride + (long long)A.pitch * (j0 - 2 + mbc - 1) + icl]);
    }
    {
        const long long ro = (long long)A.pitch * (rbeg + mbc - 1);
#pragma unroll
        for (int m = 0; m < MEQN; m++) {
            cp_async8(&qs0[m * QS + t], &A.qin[m * A.mstride + ro + cload]);
            if (t == 0) cp_async8(&qs0[m * QS + NT], &A.qin[m * A.mstride + ro + cload2]);
        }
        cp_async_commit();
    }
    int qb = 0, buf = 0;
    constexpr int kUnroll = CLAW_F_UNROLL;
#pragma unroll kUnroll
    for (int r = rbeg; r <= rend; r++) {
        const long long rowoff = (long long)A.pitch * (r + mbc - 1);
        cp_async_wait_all();
        __syncthreads();
        double *qs = qs0 + qb * (MEQN * QS);
        if (r < rend) {
            double *qsn = qs0 + (qb ^ 1) * (MEQN * QS);
            const long long ro = rowoff + A.pitch;
#pragma unroll
            for (int m = 0; m < MEQN; m++) {
                cp_async8(&qsn[m * QS + t], &A.qin[m * A.mstride + ro + cload]);
                if (t == 0) cp_async8(&qsn[m * QS + NT], &A.qin[m * A.mstride + ro + cload2]);
            }
            cp_async_commit();
        }
        qb ^= 1;
        double l[MEQN], qk[MEQN];
#pragma unroll
        for (int m = 0; m < MEQN; m++) { l[m] = qs[m * QS + t]; qk[m] = qs[m * QS + t + 1]; }

        // ================= x-sweep of row r (flux2.f, ixy = 1), slices j0-1 .. j1 =================
        const bool xrow = (CLAW_Y_PEEL || r >= j0 - 1) && (r <= j1);   // block-uniform
        if (xrow) {
            double wave[MEQN][MW], s[MW], amdq[MEQN], apdq[MEQN], roe[NROE];
            with_arith([&](auto &ar) { RPX::solve(ar, A.rp, l, qk, nocell, nocell, wave, s, amdq, apdq, roe); });
            if (xiface_ok) {
#pragma unroll
                for (int mw = 0; mw < MW; mw++) smax_update(smaxx, s[mw]);
            }
            if (order2) {
#pragma unroll
                for (int m = 0; m < MEQN; m++)
#pragma unroll
                    for (int mw = 0; mw < MW; mw++)
                        if (RPX::nz(m, mw)) ws[(m * MW + mw) * NT + t] = wave[m][mw];
            }
            __syncthreads();
            double cqxx[MEQN], bmp[MEQN], bpp[MEQN], bmm[MEQN], bpm[MEQN];
#pragma unroll
            for (int m = 0; m < MEQN; m++) cqxx[m] = 0.0;
            double wnorm2[MW], dotu[MW];
            const bool lim = order2 && t >= 1 && t <= NT - 2;
            if (lim) {
#pragma unroll
                for (int mw = 0; mw < MW; mw++) {
                    const int nb = (s[mw] > 0.0) ? t - 1 : t + 1;
                    double n2 = 0.0, du = 0.0;
#pragma unroll
                    for (int m = 0; m < MEQN; m++) {
                        if (RPX::nz(m, mw)) {
                            double w = wave[m][mw];
                            n2 = n2 + w * w;
                            du = du + ws[(m * MW + mw) * NT + nb] * w;
                        }
                    }
                    wnorm2[mw] = n2; dotu[mw] = du;
                }
            }
            with_arith([&](auto &ar) {
                if (!ar.FAST && lim) {
#pragma unroll
                    for (int m = 0; m < MEQN; m++)
#pragma unroll
                        for (int mw = 0; mw < MW; mw++)
                            if (RPX::nz(m, mw)) wave[m][mw] = ws[(m * MW + mw) * NT + t];
                }
                if (lim) {
                    limit_waves<RPX>(ar, wave, s, wnorm2, dotu, dotu, A.mthlim);
                    double dtdxave = 0.5 * (dtdx + dtdx);
                    second_order<RPX>(wave, s, dtdxave, cqxx);
                }
                if (A.trans > 0) {
                    double asdq[MEQN];
#pragma unroll
                    for (int m = 0; m < MEQN; m++) asdq[m] = trans2 ? (amdq[m] + cqxx[m]) : amdq[m];
                    RPX::transverse(ar, A.rp, roe, l, nocell, nocell, nocell, asdq, bmm, bpm);
#pragma unroll
                    for (int m = 0; m < MEQN; m++) asdq[m] = trans2 ? (apdq[m] - cqxx[m]) : apdq[m];
                    RPX::transverse(ar, A.rp, roe, qk, nocell, nocell, nocell, asdq, bmp, bpp);
                } else {
#pragma unroll
                    for (int m = 0; m < MEQN; m++) bmm[m] = bpm[m] = bmp[m] = bpp[m] = 0.0;
                }
            });
            double F[MEQN];
#pragma unroll
            for (int m = 0; m < MEQN; m++) {
                F[m] = 0.5 * cqxx[m];
                xs[m * NT + t] = amdq[m];
                xs[(MEQN + m) * NT + t] = F[m];
                xs[(2 * MEQN + m) * NT + t] = bmm[m];
                xs[(3 * MEQN + m) * NT + t] = bpm[m];
            }
            __syncthreads();
            if (t <= NT - 2) {
                const int tn = t + 1;
#pragma unroll
                for (int m = 0; m < MEQN; m++) {
                    double qaddv = (0.0 - dtdx * apdq[m]) - dtdx * xs[m * NT + tn];
                    double dF = xs[(MEQN + m) * NT + tn] - F[m];
                    double G1 = (0.0 - hdtdx * xs[(2 * MEQN + m) * NT + tn]) - hdtdx * bmp[m];
                    double G2 = (0.0 - hdtdx * xs[(3 * MEQN + m) * NT + tn]) - hdtdx * bpp[m];
                    xd2[m] = xd1[m];
                    xd1[m] = accPrev[m] - dtdy * G1;   // row r-1: last x-sweep contribution
                    double acc = qk[m] + pendA[m];
                    acc = acc + qaddv - dtdx * dF - dtdy * (G2 - G1);
                    pendA[m] = dtdy * G2;
                    accPrev[m] = acc;
                }
            }
        } else {
#pragma unroll
            for (int m = 0; m < MEQN; m++) { xd2[m] = xd1[m]; xd1[m] = accPrev[m]; }
        }

        // ================= y-sweep: interface r (rows r-1 | r) of this column ======================
        double wave[MEQN][MW], s[MW], amdq[MEQN], apdq[MEQN], roe[NROE];
        if (CLAW_Y_PEEL || r >= j0 - 1) {
            with_arith([&](auto &ar) { RPY::solve(ar, A.rp, qm1, qk, nocell, nocell, wave, s, amdq, apdq, roe); });
            if (ycol_cfl && r >= 1 && r <= A.my + 1) {
#pragma unroll
                for (int mw = 0; mw < MW; mw++) smax_update(smaxy, s[mw]);
            }
        } else {
#pragma unroll
            for (int m = 0; m < MEQN; m++) {
                amdq[m] = apdq[m] = 0.0;
#pragma unroll
                for (int mw = 0; mw < MW; mw++) wave[m][mw] = 0.0;
            }
#pragma unroll
            for (int mw = 0; mw < MW; mw++) s[mw] = 0.0;
#pragma unroll
            for (int n = 0; n < NROE; n++) roe[n] = 1.0;
        }
        double normk[MW], dotk[MW];
#pragma unroll
        for (int mw = 0; mw < MW; mw++) {
            double n2 = 0.0, d = 0.0;
#pragma unroll
            for (int m = 0; m < MEQN; m++) {
                if (RPY::nz(m, mw)) {
                    n2 = n2 + wave[m][mw] * wave[m][mw];
                    d = d + wl[m][mw] * wave[m][mw];
                }
            }
            normk[mw] = n2; dotk[mw] = d;
        }
        // limit interface r-1, form its correction flux and split it transversely
        double cqxx[MEQN], F1[MEQN], bmm[MEQN], bpm[MEQN], bmp1[MEQN], bpp1[MEQN];
        {
            const bool lim = order2 && r >= j0 + 1;
            with_arith([&](auto &ar) {
                double wlim[MEQN][MW];
#pragma unroll
                for (int m = 0; m < MEQN; m++) {
                    cqxx[m] = 0.0;
#pragma unroll
                    for (int mw = 0; mw < MW; mw++) wlim[m][mw] = wl[m][mw];
                }
                if (lim) {
                    limit_waves<RPY>(ar, wlim, sm1, norm1, dot1, dotk, A.mthlim);
                    double dtdxave = 0.5 * (dtdy + dtdy);
                    second_order<RPY>(wlim, sm1, dtdxave, cqxx);
                }
                if (A.trans > 0) {
                    double asdq[MEQN];
#pragma unroll
                    for (int m = 0; m < MEQN; m++) asdq[m] = trans2 ? (am1[m] + cqxx[m]) : am1[m];
                    RPY::transverse(ar, A.rp, roe1, qm2, nocell, nocell, nocell, asdq, bmm, bpm);
#pragma unroll
                    for (int m = 0; m < MEQN; m++) asdq[m] = trans2 ? (ap1[m] - cqxx[m]) : ap1[m];
                    RPY::transverse(ar, A.rp, roe1, qm1, nocell, nocell, nocell, asdq, bmp1, bpp1);
                } else {
#pragma unroll
                    for (int m = 0; m < MEQN; m++) bmm[m] = bpm[m] = bmp1[m] = bpp1[m] = 0.0;
                }
            });
        }
#pragma unroll
        for (int m = 0; m < MEQN; m++) F1[m] = 0.5 * cqxx[m];

        // complete cell row r-2
        const int jc = r - 2;
        const bool row_out = (jc >= j0) && (jc < j1);
        double mainE[MEQN];
        if (row_out) {
#pragma unroll
            for (int m = 0; m < MEQN; m++) {
                double qaddv = (0.0 - dtdy * ap2[m]) - dtdy * am1[m];
                double dF = F1[m] - f2[m];
                double G1 = (0.0 - hdtdy * bmm[m]) - hdtdy * bmp2[m];
                double G2 = (0.0 - hdtdy * bpm[m]) - hdtdy * bpp2[m];
                mainE[m] = (qaddv - dtdy * dF - dtdx * (G2 - G1));
                gs[(buf * 2 * MEQN + m) * NT + t] = G1;
                gs[(buf * 2 * MEQN + MEQN + m) * NT + t] = G2;
            }
        }
        __syncthreads();
        if (row_out && col_out) {
            const long long oidx = (long long)A.pitch * (jc + mbc - 1) + icl;
#pragma unroll
            for (int m = 0; m < MEQN; m++) {
                double G2l = gs[(buf * 2 * MEQN + MEQN + m) * NT + t - 1];
                double G1r = gs[(buf * 2 * MEQN + m) * NT + t + 1];
                double q = xd2[m];
                q = q + dtdx * G2l;
                q = q + mainE[m];
                q = q - dtdx * G1r;
                A.qout[m * A.mstride + oidx] = q;
            }
        }
        buf ^= 1;

        // shift the y-window
#pragma unroll
        for (int m = 0; m < MEQN; m++) {
            qm2[m] = qm1[m]; qm1[m] = qk[m];
            ap2[m] = ap1[m]; f2[m] = F1[m];
            bmp2[m] = bmp1[m]; bpp2[m] = bpp1[m];
            am1[m] = amdq[m]; ap1[m] = apdq[m];
#pragma unroll
            for (int mw = 0; mw < MW; mw++) wl[m][mw] = RPY::nz(m, mw) ? wave[m][mw] : 0.0;
        }
#pragma unroll
        for (int n = 0; n < NROE; n++) roe1[n] = roe[n];
#pragma unroll
        for (int mw = 0; mw < MW; mw++) { sm1[mw] = s[mw]; norm1[mw] = normk[mw]; dot1[mw] = dotk[mw]; }
    }
    // the two sweeps' Courant numbers (flux2.f:109-117), maximum of the two (step2.f:121,206)
    double cflx = dtdx * __longlong_as_double((long long)smaxx);
    double cfly = dtdy * __longlong_as_double((long long)smaxy);
    cfl_commit(dmax2(cflx, cfly), A.cfl_bits);
}

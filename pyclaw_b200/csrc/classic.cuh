// classic.cuh -- fused wave-propagation sweeps (classic Clawpack) for sm_100a.
//
// Replaces, per sweep, the reference's step1 / step2 / step2ds + flux2 + limiter +
// philim + rpn2 + rpt2 call tree
//   src/fortran/1d/classic/step1.f:4-142, limiter.f:4-60, philim.f:4-58
//   src/fortran/2d/classic/step2.f:2-241, step2ds.f:2-248, flux2.f:5-193
// with kernels that do the Riemann solve, wave limiting, second-order correction,
// transverse solves and the flux-difference update in one pass over q.
//
// Data layout (HBM): structure-of-arrays, q[m][j][i] with i fastest; `pitch` is the
// row stride and `mstride` the component stride, both in doubles.  Ghost cells are
// stored in place (mbc on every side), Fortran cell (i,j) lives at
// (i+mbc-1) + pitch*(j+mbc-1).
//
// Two engines, both with one thread per column i so that every global access is a
// unit-stride, coalesced float64 stream:
//   * x-engine: a CTA owns a strip of columns and walks down the rows.  Riemann
//     problems are between neighbouring threads; q rows, waves and the "goes to the
//     left cell" parts are exchanged through shared memory.  Transverse increments go
//     to rows j-1 / j+1, i.e. stay in the thread (rolling accumulators).
//   * y-engine: a thread walks down its own column keeping a rolling window of the
//     last interfaces in registers; only the transverse increments (which go to the
//     columns i-1 / i+1) are exchanged through shared memory.
// The arithmetic order of every cell update is the reference's.
#pragma once
// tuning switches (scratch/build_variant.py builds variants with other values)
#ifndef CLAW_Y_UNROLL
#define CLAW_Y_UNROLL 1
#endif
#ifndef CLAW_F_UNROLL
#define CLAW_F_UNROLL 1
#endif
#ifndef CLAW_Y_PEEL
#define CLAW_Y_PEEL 1
#endif
#include "rp.cuh"

#define CLAW_MAXWAVES 8

struct SweepArgs {
    const double *qin;   // Riemann data (qold, ghost cells filled)
    const double *qbase; // values being updated (== qin for the first sweep)
    double *qout;
    long long mstride;
    int pitch;
    int mx, my, mbc;
    double dtdx, dtdy; // dt/dx, dt/dy (used when dt_dev is null)
    double dt;         // the time step itself (step1 with a capacity function)
    double dx, dy;
    const double *dt_dev; // if set, the time step is read from device memory: the launch
                          // sequence of a step is then independent of dt (CUDA-graph replay)
    int order;         // method(2)
    int trans;         // method(3): -1 dim-split, 0 none, 1 increment, 2 increment+correction
    int mthlim[CLAW_MAXWAVES];
    RpParams rp;
    unsigned long long *cfl_bits; // running max of the Courant number, as ordered bits (cfl >= 0)
    int rows_per_cta;
    int jlo, jhi; // first / last output row (Fortran index)
    int ilo, ihi; // first / last output column
    const double *aux;      // aux[ma][j][i], same pitch as q (or null)
    long long amstride;     // component stride of aux
    int mcapa;              // method(6): 0 = none, else 1-based aux component of the capacity
};

// aux of Fortran cell (i, j), clamped into the padded array (threads outside the strip
// compute throw-away values from valid memory)
__device__ __forceinline__ AuxCell aux_cell(const SweepArgs &A, int i, int j)
{
    int ic = min(max(i, 1 - A.mbc), A.mx + A.mbc) + A.mbc - 1;
    int jc = min(max(j, 1 - A.mbc), A.my + A.mbc) + A.mbc - 1;
    return AuxCell{A.aux + (long long)A.pitch * jc + ic, A.amstride};
}

// dt/dx, dt/dy: from the host-computed values or, for graph replay, from the device scalar
// (one IEEE division, the same bits as the host's `dt / dx`)
__device__ __forceinline__ void load_dt(const SweepArgs &A, double &dtdx, double &dtdy)
{
    dtdx = A.dtdx;
    dtdy = A.dtdy;
    if (A.dt_dev) {
        const double dt = __ldg(A.dt_dev);
        dtdx = dt / A.dx;
        dtdy = dt / A.dy;
    }
}

// a / b with one correctly rounded division, fast path first (arith.cuh)
__device__ __forceinline__ double div1(double a, double b)
{
    FastArith fa;
    double v = fa.div(a, b);
    if (fa.bad()) v = a / b;
    return v;
}

// philim.f:4-58
template <class AR>
__device__ __forceinline__ double philim(AR &ar, double a, double b, int meth)
{
    double r = ar.div(b, a);
    switch (meth) {
    case 2: return dmax2(dmax2(0.0, dmin2(1.0, 2.0 * r)), dmin2(2.0, r));
    case 3: return ar.div(r + fabs(r), 1.0 + fabs(r));
    case 4: {
        double c = (1.0 + r) / 2.0;
        return dmax2(0.0, dmin2(dmin2(c, 2.0), 2.0 * r));
    }
    case 5: return r;
    // philim.f:19: a computed GO TO with an index outside 1..5 falls through to label 10
    default: return dmax2(0.0, dmin2(1.0, r));
    }
}

// limiter.f:29-55 for one interface, given the (unlimited) dot products with the
// neighbouring interfaces.  Entries with RP::nz == false are structurally zero and
// skipped: adding +0 to a running sum that started at +0 never changes it.
template <class RP, class AR>
__device__ __forceinline__ void limit_waves(AR &ar, double (&wave)[RP::MEQN][RP::MWAVES],
                                            const double (&s)[RP::MWAVES],
                                            const double (&wnorm2)[RP::MWAVES],
                                            const double (&dotl)[RP::MWAVES],
                                            const double (&dotr)[RP::MWAVES], const int *mthlim)
{
#pragma unroll
    for (int mw = 0; mw < RP::MWAVES; mw++) {
        if (mthlim[mw] == 0) continue;
        if (wnorm2[mw] == 0.0) continue;
        double wlimitr = philim(ar, wnorm2[mw], (s[mw] > 0.0) ? dotl[mw] : dotr[mw], mthlim[mw]);
#pragma unroll
        for (int m = 0; m < RP::MEQN; m++)
            if (RP::nz(m, mw)) wave[m][mw] = wlimitr * wave[m][mw];
    }
}

// flux2.f:127-145: cqxx(m) = sum_mw |s| (1 - |s| dtdxave) wave(m,mw);
// flux2fw.f:151-152 for f-wave solvers: sign(s) in place of the leading |s|
template <class RP>
__device__ __forceinline__ double lead_factor(double s)
{
    return rp_is_fwave<RP>::value ? copysign(1.0, s) : fabs(s);
}
template <class RP>
__device__ __forceinline__ void second_order(const double (&wave)[RP::MEQN][RP::MWAVES],
                                             const double (&s)[RP::MWAVES], double dtdxave,
                                             double (&cqxx)[RP::MEQN])
{
#pragma unroll
    for (int m = 0; m < RP::MEQN; m++) {
        double c = 0.0;
#pragma unroll
        for (int mw = 0; mw < RP::MWAVES; mw++)
            if (RP::nz(m, mw)) c = c + lead_factor<RP>(s[mw]) * (1.0 - fabs(s[mw]) * dtdxave) * wave[m][mw];
        cqxx[m] = c;
    }
}

__device__ __forceinline__ void cfl_commit(double cfl, unsigned long long *cfl_bits)
{
    // warp-shuffle max, then one atomicMax per warp on the ordered bit pattern
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double other = __shfl_xor_sync(0xffffffffu, cfl, o);
        cfl = dmax2(cfl, other);
    }
    if ((threadIdx.x & 31) == 0 && cfl > 0.0)
        atomicMax(cfl_bits, (unsigned long long)__double_as_longlong(cfl));
}

// Without a capacity function the Courant number of a sweep is dtdx * max|s|: rounding is
// monotone, so max_i fl(dtdx |s_i|) == fl(dtdx max_i |s_i|) bit for bit.  The running maximum
// of |s| is kept as an integer maximum of the bit patterns (non-negative doubles order like
// unsigned integers) -- four FP64-pipe instructions per wave and interface less, on the pipe
// that bounds these kernels.
__device__ __forceinline__ void smax_update(unsigned long long &smax, double s)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(s) & 0x7fffffffffffffffULL;
    smax = (b > smax) ? b : smax;
}

// ---------------------------------------------------------------------------
// x-engine.  Thread t <-> interface/cell ii = i0-1+t ; cells i0 .. i0+NT-4 are output.
// TRANS=false: dimensional splitting (step2ds.f ids=1), every row independent.
// TRANS=true : unsplit (step2.f x-sweeps): slices j0-1 .. j1 are processed for output
//              rows j0 .. j1-1, contributions applied in the order of SURVEY.md A.3.
// ---------------------------------------------------------------------------
template <class RP, bool TRANS, bool CAPA, int NT>
__global__ void __launch_bounds__(NT, RP::X_MINB) xsweep_kernel(const SweepArgs A)
{
    constexpr bool AUXRP = (RP::MAUX > 0);
    constexpr int MEQN = RP::MEQN, MW = RP::MWAVES, NROE = RP::NROE;
    constexpr int NC = NT - 3;
    constexpr int QS = NT + 1;
    extern __shared__ double sm[];
    double *qs0 = sm;                // [2][MEQN][NT+1] staged q rows (double buffered)
    double *qs = qs0;
    double *ws = qs0 + 2 * MEQN * QS; // [MEQN*MW][NT]  unlimited waves
    double *xs = ws + MEQN * MW * NT; // [4*MEQN][NT]   amdq, F, bm(A-), bp(A-) of each interface
    // Solvers that read many aux components per interface (the sphere: ~70 loads, two thirds of them
    // L1 misses with 8 warps per SM to hide them): the aux rows r-1, r, r+1 live in a ring of four
    // shared-memory rows, row r+2 arrives by cp.async during iteration r.  Column c <-> cell i0-2+c.
    constexpr bool AUXS = rp_x_aux_smem<RP>::value && TRANS;
    constexpr int AQS = NT + 2;
    double *as = xs + 4 * MEQN * NT;  // [4][MAUX][NT+2] (AUXS only)

    const int t = threadIdx.x;
    const int i0 = A.ilo + blockIdx.x * NC;
    const int ii = i0 - 1 + t;
    const int j0 = A.jlo + blockIdx.y * A.rows_per_cta;
    const int j1 = min(j0 + A.rows_per_cta, A.jhi + 1);
    const int mbc = A.mbc;
    const int imax = A.mx + mbc;
    const int cload = min(i0 - 2 + t, imax) + mbc - 1;      // array column this thread stages
    const int cload2 = min(i0 - 2 + NT, imax) + mbc - 1;    // extra column staged by thread 0
    const bool cell_ok = (t >= 1) && (t <= NC) && (ii <= A.ihi);
    auto stage_aux_row = [&](int row) { // aux row `row` (clamped like aux_cell) into its ring slot
        if constexpr (AUXS) {
            const int jc = min(max(row, 1 - mbc), A.my + mbc) + mbc - 1;
            double *dst = as + (row & 3) * (RP::MAUX * AQS);
            const double *src = A.aux + (long long)A.pitch * jc;
            const int c0 = min(max(i0 - 2 + t, 1 - mbc), imax) + mbc - 1;
#pragma unroll
            for (int ma = 0; ma < RP::MAUX; ma++) cp_async8(&dst[ma * AQS + t], &src[ma * A.amstride + c0]);
            if (t < 2) {
                const int c1 = min(i0 - 2 + NT + t, imax) + mbc - 1;
#pragma unroll
                for (int ma = 0; ma < RP::MAUX; ma++) cp_async8(&dst[ma * AQS + NT + t], &src[ma * A.amstride + c1]);
            }
        }
    };
    // the aux cell di columns right of cell ii (= column t+1+di of the ring) in row `row`
    auto ax = [&](int di, int row) {
        if constexpr (AUXS) return AuxCellS{as + (row & 3) * (RP::MAUX * AQS) + (t + 1 + di), AQS};
        else return (AUXRP || CAPA) ? aux_cell(A, ii + di, row) : AuxCell{nullptr, 0};
    };
    const bool iface_ok = (ii >= 1) && (ii <= A.mx + 1) && (t >= 1) && (t <= NT - 2);
    const bool order2 = (A.order != 1);
    const bool trans2 = order2 && (A.trans == 2);
    double dtdx, dtdy;
    load_dt(A, dtdx, dtdy);
    const AuxCell nocell{nullptr, 0};

    double cfl = 0.0;
    unsigned long long smax = 0ULL; // bits of max|s| (CAPA == false)
    double accPrev[MEQN], pendA[MEQN];
#pragma unroll
    for (int m = 0; m < MEQN; m++) { accPrev[m] = 0.0; pendA[m] = 0.0; }

    const int rbeg = TRANS ? j0 - 1 : j0;
    const int rend = TRANS ? j1 : j1 - 1;
    // Asynchronous staging (cp.async): the row for iteration r+1 is requested before the
    // arithmetic of row r starts, straight into the other half of a double-buffered
    // shared-memory row, so its HBM latency hides behind a full row of Riemann solves.
    {
        const long long ro = (long long)A.pitch * (rbeg + mbc - 1);
#pragma unroll
        for (int m = 0; m < MEQN; m++) {
            cp_async8(&qs[m * QS + t], &A.qin[m * A.mstride + ro + cload]);
            if (t == 0) cp_async8(&qs[m * QS + NT], &A.qin[m * A.mstride + ro + cload2]);
        }
        stage_aux_row(rbeg - 1);
        stage_aux_row(rbeg);
        stage_aux_row(rbeg + 1);
        cp_async_commit();
    }
    int qb = 0;
    for (int r = rbeg; r <= rend; r++) {
        const long long rowoff = (long long)A.pitch * (r + mbc - 1);
        if (RP::MAUX >= 8 && !AUXS) { // aux of row r+2 (first needed as the "row above" of iteration r+1)
            const int jp = min(r + 2, A.my + mbc) + mbc - 1;
            const double *ap = A.aux + (long long)A.pitch * jp + (min(max(ii, 1 - mbc), imax) + mbc - 1);
#pragma unroll
            for (int ma = 0; ma < RP::MAUX; ma++) prefetch_l1(ap + ma * A.amstride);
        }
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
            stage_aux_row(r + 2); // replaces row r-2, last read in iteration r-1
            cp_async_commit();
        }
        qb ^= 1;
        double l[MEQN], rr[MEQN];
#pragma unroll
        for (int m = 0; m < MEQN; m++) { l[m] = qs[m * QS + t]; rr[m] = qs[m * QS + t + 1]; }

        // capacity function: dtdx1d(i) = dtdx / capa(i,j)  (step2.f:91-95)
        double dtdx_c = dtdx, dtdx_l = dtdx, capa_c = 1.0, capa_m = 1.0, capa_p = 1.0;
        if (CAPA) {
            capa_c = ax(0, r)(A.mcapa - 1);
            dtdx_c = div1(dtdx, capa_c);
            dtdx_l = div1(dtdx, ax(-1, r)(A.mcapa - 1));
            if (TRANS) {
                capa_m = ax(0, r - 1)(A.mcapa - 1);
                capa_p = ax(0, r + 1)(A.mcapa - 1);
            }
        }
        const double hdtdx = 0.5 * dtdx_c;
        const auto axl = ax(-1, r);
        const auto axr = ax(0, r);

        double wave[MEQN][MW], s[MW], amdq[MEQN], apdq[MEQN], roe[NROE];
        with_arith([&](auto &ar) { RP::solve(ar, A.rp, l, rr, axl, axr, wave, s, amdq, apdq, roe); });
        if (iface_ok) {
#pragma unroll
            for (int mw = 0; mw < MW; mw++) {
                if (CAPA) cfl = dmax2(dmax2(cfl, dtdx_c * s[mw]), -dtdx_l * s[mw]);
                else smax_update(smax, s[mw]);
            }
        }
        if (order2) {
#pragma unroll
            for (int m = 0; m < MEQN; m++)
#pragma unroll
                for (int mw = 0; mw < MW; mw++)
                    if (RP::nz(m, mw)) ws[(m * MW + mw) * NT + t] = wave[m][mw];
        }
        __syncthreads();

        double cqxx[MEQN];
#pragma unroll
        for (int m = 0; m < MEQN; m++) cqxx[m] = 0.0;
        double bmp[MEQN], bpp[MEQN], bmm[MEQN], bpm[MEQN];
        // limiter.f:35-43 takes the dot product with the UPWIND neighbour's wave: only that one
        // is formed (a * b == b * a bit for bit, so the product order of dotl / dotr is immaterial)
        double wnorm2[MW], dotu[MW];
        if (order2 && t >= 1 && t <= NT - 2) {
#pragma unroll
            for (int mw = 0; mw < MW; mw++) {
                const int nb = (s[mw] > 0.0) ? t - 1 : t + 1;
                double n2 = 0.0, du = 0.0;
#pragma unroll
                for (int m = 0; m < MEQN; m++) {
                    if (RP::nz(m, mw)) {
                        double w = wave[m][mw];
                        n2 = n2 + w * w;
                        du = du + ws[(m * MW + mw) * NT + nb] * w;
                    }
                }
                wnorm2[mw] = n2; dotu[mw] = du;
            }
        }
        // limiter, second-order correction and the two transverse solves: one block of
        // branch-free arithmetic (the exact re-run restores the unlimited waves first)
        {
            const bool lim = order2 && t >= 1 && t <= NT - 2;
            with_arith([&](auto &ar) {
                if (!ar.FAST && lim) { // re-run: the unlimited waves are still in shared memory
#pragma unroll
                    for (int m = 0; m < MEQN; m++)
#pragma unroll
                        for (int mw = 0; mw < MW; mw++)
                            if (RP::nz(m, mw)) wave[m][mw] = ws[(m * MW + mw) * NT + t];
                }
                if (lim) {
                    limit_waves<RP>(ar, wave, s, wnorm2, dotu, dotu, A.mthlim);
                    double dtdxave = 0.5 * (dtdx_l + dtdx_c);
                    second_order<RP>(wave, s, dtdxave, cqxx);
                }
                if (TRANS) {
                    double asdq[MEQN];
                    if (A.trans > 0) {
                        // rpt2 sees the cell the fluctuation moves into, in the rows below /
                        // at / above this one (aux1, aux2, aux3 of step2.f:99-107)
#pragma unroll
                        for (int m = 0; m < MEQN; m++) asdq[m] = trans2 ? (amdq[m] + cqxx[m]) : amdq[m];
                        RP::transverse(ar, A.rp, roe, l, ax(-1, r - 1), axl, ax(-1, r + 1), asdq, bmm, bpm);
                        // hand the "goes into the left cell" parts to thread t-1 right away: they
                        // are dead in this thread, and the second solve needs the registers
#pragma unroll
                        for (int m = 0; m < MEQN; m++) {
                            xs[(2 * MEQN + m) * NT + t] = bmm[m];
                            xs[(3 * MEQN + m) * NT + t] = bpm[m];
                        }
#pragma unroll
                        for (int m = 0; m < MEQN; m++) asdq[m] = trans2 ? (apdq[m] - cqxx[m]) : apdq[m];
                        RP::transverse(ar, A.rp, roe, rr, ax(0, r - 1), axr, ax(0, r + 1), asdq, bmp, bpp);
                    } else { // flux2.f:151 -- gadd stays zero
#pragma unroll
                        for (int m = 0; m < MEQN; m++) bmm[m] = bpm[m] = bmp[m] = bpp[m] = 0.0;
                    }
                }
            });
        }
        double F[MEQN];
#pragma unroll
        for (int m = 0; m < MEQN; m++) F[m] = 0.5 * cqxx[m];

        if (TRANS && !(A.trans > 0)) { // (otherwise stored right after the first transverse solve)
#pragma unroll
            for (int m = 0; m < MEQN; m++) {
                xs[(2 * MEQN + m) * NT + t] = bmm[m];
                xs[(3 * MEQN + m) * NT + t] = bpm[m];
            }
        }
#pragma unroll
        for (int m = 0; m < MEQN; m++) {
            xs[m * NT + t] = amdq[m];
            xs[(MEQN + m) * NT + t] = F[m];
        }
        __syncthreads();

        if (t <= NT - 2) {
            const int tn = t + 1;
            const long long oidx = rowoff + (ii + mbc - 1);
            if (!TRANS) {
                // step2ds.f leaves the ghost columns of qnew equal to qold; the y-sweep (and
                // its CFL number) reads them, so carry them across.
                if (t < mbc && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) {
                    const int gc = (blockIdx.x == 0) ? t : (A.mx + mbc + t);
#pragma unroll
                    for (int m = 0; m < MEQN; m++) {
                        if (blockIdx.x == 0)
                            A.qout[m * A.mstride + rowoff + t] = A.qin[m * A.mstride + rowoff + t];
                        if (blockIdx.x == gridDim.x - 1)
                            A.qout[m * A.mstride + rowoff + A.mx + mbc + t] =
                                A.qin[m * A.mstride + rowoff + A.mx + mbc + t];
                    }
                    (void)gc;
                }
                if (cell_ok) {
#pragma unroll
                    for (int m = 0; m < MEQN; m++) {
                        double qaddv = (0.0 - dtdx_c * apdq[m]) - dtdx_c * xs[m * NT + tn];
                        double dF = xs[(MEQN + m) * NT + tn] - F[m];
                        if (!CAPA)
                            A.qout[m * A.mstride + oidx] = (rr[m] + qaddv) - dtdx * dF;
                        else // step2ds.f:152-156: the division binds to the flux difference only
                            A.qout[m * A.mstride + oidx] = (rr[m] + qaddv) - div1(dtdx * dF, capa_c);
                    }
                }
            } else {
                double qcv[MEQN];
#pragma unroll
                for (int m = 0; m < MEQN; m++) qcv[m] = 0.0;
                if constexpr (RP::QCOR) { // apps/shallow-sphere/step2qcor.f:147
                    with_arith([&](auto &ar) { RP::qcor(ar, A.rp, rr, axr, ax(1, r), qcv); });
                }
#pragma unroll
                for (int m = 0; m < MEQN; m++) {
                    double qaddv = (0.0 - dtdx_c * apdq[m]) - dtdx_c * xs[m * NT + tn];
                    double dF = xs[(MEQN + m) * NT + tn] - F[m];
                    double G1 = (0.0 - hdtdx * xs[(2 * MEQN + m) * NT + tn]) - hdtdx * bmp[m];
                    double G2 = (0.0 - hdtdx * xs[(3 * MEQN + m) * NT + tn]) - hdtdx * bpp[m];
                    // row r-1 receives its last x-sweep contribution and is complete
                    double done = CAPA ? accPrev[m] - div1(dtdy * G1, capa_m) : accPrev[m] - dtdy * G1;
                    if (cell_ok && r - 1 >= j0 && r - 1 < j1)
                        A.qout[m * A.mstride + oidx - A.pitch] = done;
                    double acc = rr[m] + pendA[m];
                    if (!CAPA) {
                        acc = acc + qaddv - dtdx * dF - dtdy * (G2 - G1);
                        pendA[m] = dtdy * G2;
                    } else { // step2.f:145-152
                        acc = acc + qaddv - div1(dtdx * dF + dtdy * (G2 - G1), capa_c);
                        if (RP::QCOR) acc = acc - div1(dtdx * qcv[m], capa_c);
                        pendA[m] = div1(dtdy * G2, capa_p);
                    }
                    accPrev[m] = acc;
                }
            }
        }
    }
    if (!CAPA) cfl = dtdx * __longlong_as_double((long long)smax);
    cfl_commit(cfl, A.cfl_bits);
}

// ---------------------------------------------------------------------------
// y-engine.  Thread <-> column; walks rows k = j0-2 .. j1+1 and completes cell k-2
// at step k.  TRANS=false: step2ds.f ids=2.  TRANS=true: step2.f y-sweeps; qout is
// updated in place (it already holds the x-sweep result), transverse increments to
// the neighbouring columns go through shared memory.
//
// The rolling window (waves, Roe averages and fluctuations of interface k-1, the
// "goes to the upper cell" parts of interface k-2) lives in thread-private shared-memory
// slots laid out [slot][thread], not in registers: the Riemann solve needs ~100 live
// registers on its own, and keeping ~50 more doubles of window across it either spills
// or halves the occupancy.
// ---------------------------------------------------------------------------
template <class RP, bool TRANS>
struct YSlots {
    static constexpr int MEQN = RP::MEQN, MW = RP::MWAVES, NROE = RP::NROE;
    static constexpr int W = 0;                       // wave(m,mw) of interface k-1 (unlimited)
    static constexpr int AM1 = W + MEQN * MW;         // amdq of interface k-1
    static constexpr int AP1 = AM1 + MEQN;            // apdq of interface k-1
    static constexpr int AP2 = AP1 + MEQN;            // apdq of interface k-2
    static constexpr int F2 = AP2 + MEQN;             // correction flux of interface k-2
    static constexpr int ROE = F2 + MEQN;             // Roe data of interface k-1   (TRANS)
    static constexpr int BMP2 = ROE + NROE;           // B^- A^+ dq of interface k-2 (TRANS)
    static constexpr int BPP2 = BMP2 + MEQN;          // B^+ A^+ dq of interface k-2 (TRANS)
    static constexpr int STATE_END = TRANS ? BPP2 + MEQN : ROE;
    static constexpr int QN = STATE_END;              // [2][MEQN] cp.async landing slots: next row of qin
    static constexpr int QX = QN + 2 * MEQN;          // [2][MEQN] cp.async landing slots: x-sweep result (TRANS)
    static constexpr int COUNT = TRANS ? QX + 2 * MEQN : QX;
};

template <class RP, bool TRANS, bool CAPA, int NT>
__global__ void __launch_bounds__(NT, RP::Y_MINB) ysweep_kernel(const SweepArgs A)
{
    constexpr bool AUXRP = (RP::MAUX > 0);
    constexpr int MEQN = RP::MEQN, MW = RP::MWAVES, NROE = RP::NROE;
    constexpr int NC = TRANS ? NT - 2 : NT;
    using SL = YSlots<RP, TRANS>;
    extern __shared__ double sm[];
    double *gs = sm;                                   // [2][2*MEQN][NT] G1', G2' exchange (TRANS)
    double *ys = sm + (TRANS ? 4 * MEQN * NT : 0);     // [SL::COUNT][NT] rolling window

    const int t = threadIdx.x;
#define YQ(slot) ys[(slot) * NT + t]
    // Light solvers (RP::Y_REGS) keep the rolling window in registers: they have the registers to
    // spare, and the window traffic (~50 LDS / STS per row for acoustics) competes with the
    // arithmetic for issue slots.  Every window index is a compile-time constant after unrolling.
    constexpr bool YREGS = rp_y_regs<RP>::value;
    double yr[YREGS ? SL::STATE_END : 1];
#define YS(slot) (*(YREGS ? &yr[YREGS ? (slot) : 0] : &ys[(slot) * NT + t]))
    const int mbc = A.mbc;
    const int i0 = A.ilo + blockIdx.x * NC;
    const int ic = TRANS ? i0 - 1 + t : i0 + t;      // this thread's column (Fortran index)
    const int icl = min(ic, A.mx + mbc) + mbc - 1;   // clamped array column
    const bool col_out = TRANS ? (t >= 1 && t <= NC && ic <= A.ihi) : (ic <= A.ihi);
    const bool col_cfl = TRANS ? (ic >= 0 && ic <= A.mx + 1) : (ic <= A.ihi);
    const int j0 = A.jlo + blockIdx.y * A.rows_per_cta;
    const int j1 = min(j0 + A.rows_per_cta, A.jhi + 1);
    const bool order2 = (A.order != 1);
    const bool trans2 = order2 && (A.trans == 2);
    double dtdx, dtdy;
    load_dt(A, dtdx, dtdy);
    const AuxCell nocell{nullptr, 0};
    // capacity function: dtdy1d(j) = dtdy / capa(i,j), rolling along the column
    double dy_k = dtdy, dy_1 = dtdy, dy_2 = dtdy, cap_1 = 1.0, cap_2 = 1.0;

    double cfl = 0.0;
    // integer running maximum of |s| (see smax_update)
    constexpr bool ICFL = !CAPA;
    unsigned long long smax = 0ULL;
    double qm1[MEQN], qm2[MEQN], sm1[MW], norm1[MW], dot1[MW];
#pragma unroll
    for (int m = 0; m < MEQN; m++) { qm1[m] = 1.0; qm2[m] = 1.0; }
#pragma unroll
    for (int mw = 0; mw < MW; mw++) { sm1[mw] = 0.0; norm1[mw] = 0.0; dot1[mw] = 0.0; }
#pragma unroll
    for (int sl = 0; sl < SL::STATE_END; sl++) YS(sl) = (TRANS && sl >= SL::ROE && sl < SL::ROE + NROE) ? 1.0 : 0.0;

    int buf = 0;
    // Row j0-2 only enters as the left state of interface j0-1: with CLAW_Y_PEEL it is loaded
    // straight into the window and the walk starts at k = j0-1, so that every iteration solves an
    // interface (the start-up iteration with its zero-filled alternative to the solve -- a register
    // move per solver output where the two paths merge -- is gone).
    const int kbeg = CLAW_Y_PEEL ? j0 - 1 : j0 - 2;
    if (CLAW_Y_PEEL) {
#pragma unroll
        for (int m = 0; m < MEQN; m++)
            qm1[m] = __ldg(&A.qin[m * A.mstride + (long long)A.pitch * (j0 - 2 + mbc - 1) + icl]);
        if (CAPA) {
            cap_1 = aux_cell(A, ic, j0 - 2)(A.mcapa - 1);
            dy_1 = div1(dtdy, cap_1);
        }
    }
    // cp.async staging of row k+1 (and, TRANS, of the x-sweep result of row k-1) into
    // thread-private slots, double buffered on the parity of k
#pragma unroll
    for (int m = 0; m < MEQN; m++)
        cp_async8(&YQ(SL::QN + m), &A.qin[m * A.mstride + (long long)A.pitch * (kbeg + mbc - 1) + icl]);
    cp_async_commit();
    int par = 0;
    constexpr int kUnroll = CLAW_Y_UNROLL;
#pragma unroll kUnroll
    for (int k = kbeg; k <= j1 + 1; k++) {
        const long long rowoff = (long long)A.pitch * (k + mbc - 1);
        if (RP::MAUX >= 8) { // aux of row k+2, a full iteration before its first use
            const int jp = min(k + 2, A.my + mbc) + mbc - 1;
            const double *ap = A.aux + (long long)A.pitch * jp + icl;
#pragma unroll
            for (int ma = 0; ma < RP::MAUX; ma++) prefetch_l1(ap + ma * A.amstride);
        }
        cp_async_wait_all();
        double qk[MEQN];
#pragma unroll
        for (int m = 0; m < MEQN; m++) qk[m] = YQ(SL::QN + par * MEQN + m);
        const int qxslot = SL::QX + par * MEQN; // x-sweep result of row k-2, requested last iteration
        if (k <= j1) {
#pragma unroll
            for (int m = 0; m < MEQN; m++)
                cp_async8(&YQ(SL::QN + (par ^ 1) * MEQN + m), &A.qin[m * A.mstride + rowoff + A.pitch + icl]);
        }
        if (TRANS) {
            const bool need = (k - 1 >= j0) && (k - 1 < j1) && col_out;
            if (need) {
#pragma unroll
                for (int m = 0; m < MEQN; m++)
                    cp_async8(&YQ(SL::QX + (par ^ 1) * MEQN + m),
                              &A.qout[m * A.mstride + (long long)A.pitch * (k - 1 + mbc - 1) + icl]);
            }
        }
        cp_async_commit();
        par ^= 1;
        double wave[MEQN][MW], s[MW], amdq[MEQN], apdq[MEQN], roe[NROE];
        double normk[MW], dotk[MW];
        double wl[MEQN][MW]; // unlimited waves of interface k-1
        double cap_k = 1.0;
        if (CAPA) {
            cap_k = aux_cell(A, ic, k)(A.mcapa - 1);
            dy_k = div1(dtdy, cap_k);
        }
        if (CLAW_Y_PEEL || k >= j0 - 1) {
            with_arith([&](auto &ar) {
                RP::solve(ar, A.rp, qm1, qk, AUXRP ? aux_cell(A, ic, k - 1) : nocell,
                          AUXRP ? aux_cell(A, ic, k) : nocell, wave, s, amdq, apdq, roe);
            });
            if (col_cfl && k >= 1 && k <= A.my + 1) {
#pragma unroll
                for (int mw = 0; mw < MW; mw++) {
                    if (!ICFL) cfl = dmax2(dmax2(cfl, dy_k * s[mw]), -dy_1 * s[mw]);
                    else smax_update(smax, s[mw]);
                }
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
#pragma unroll
        for (int mw = 0; mw < MW; mw++) {
            double n2 = 0.0, d = 0.0;
#pragma unroll
            for (int m = 0; m < MEQN; m++) {
                if (RP::nz(m, mw)) {
                    wl[m][mw] = YS(SL::W + m * MW + mw);
                    n2 = n2 + wave[m][mw] * wave[m][mw];
                    d = d + wl[m][mw] * wave[m][mw];
                } else {
                    wl[m][mw] = 0.0;
                }
            }
            normk[mw] = n2; dotk[mw] = d;
        }
        // Interface k-1's waves, fluctuations and Roe data are in registers now: put
        // interface k's into the slots at once instead of at the end of the iteration, so that
        // they are not live (32 doubles for Euler) across the limiter and the transverse solves.
#pragma unroll
        for (int m = 0; m < MEQN; m++)
#pragma unroll
            for (int mw = 0; mw < MW; mw++)
                if (RP::nz(m, mw)) YS(SL::W + m * MW + mw) = wave[m][mw];

        // limit interface k-1, form its correction flux and split it transversely
        double cqxx[MEQN], F1[MEQN], amdq1[MEQN], apdq1[MEQN];
        double bmm[MEQN], bpm[MEQN], bmp1[MEQN], bpp1[MEQN];
#pragma unroll
        for (int m = 0; m < MEQN; m++) {
            amdq1[m] = YS(SL::AM1 + m); apdq1[m] = YS(SL::AP1 + m);
            YS(SL::AM1 + m) = amdq[m]; YS(SL::AP1 + m) = apdq[m];
        }
        {
            const bool lim = order2 && k >= j0 + 1;
            double roe1[NROE];
            if (TRANS) {
#pragma unroll
                for (int n = 0; n < NROE; n++) { roe1[n] = YS(SL::ROE + n); YS(SL::ROE + n) = roe[n]; }
            }
            with_arith([&](auto &ar) {
                double wlim[MEQN][MW];
#pragma unroll
                for (int m = 0; m < MEQN; m++) {
                    cqxx[m] = 0.0;
#pragma unroll
                    for (int mw = 0; mw < MW; mw++) wlim[m][mw] = wl[m][mw];
                }
                if (lim) {
                    limit_waves<RP>(ar, wlim, sm1, norm1, dot1, dotk, A.mthlim);
                    double dtdxave = 0.5 * (dy_2 + dy_1);
                    second_order<RP>(wlim, sm1, dtdxave, cqxx);
                }
                if (TRANS) {
                    if (A.trans > 0) {
                        // interface k-1: A- dq moves into cell k-2, A+ dq into cell k-1; the
                        // transverse slices are the columns i-1, i, i+1 (step2.f:178-186)
                        double asdq[MEQN];
#pragma unroll
                        for (int m = 0; m < MEQN; m++) asdq[m] = trans2 ? (amdq1[m] + cqxx[m]) : amdq1[m];
                        RP::transverse(ar, A.rp, roe1, qm2, AUXRP ? aux_cell(A, ic - 1, k - 2) : nocell,
                                       AUXRP ? aux_cell(A, ic, k - 2) : nocell,
                                       AUXRP ? aux_cell(A, ic + 1, k - 2) : nocell, asdq, bmm, bpm);
#pragma unroll
                        for (int m = 0; m < MEQN; m++) asdq[m] = trans2 ? (apdq1[m] - cqxx[m]) : apdq1[m];
                        RP::transverse(ar, A.rp, roe1, qm1, AUXRP ? aux_cell(A, ic - 1, k - 1) : nocell,
                                       AUXRP ? aux_cell(A, ic, k - 1) : nocell,
                                       AUXRP ? aux_cell(A, ic + 1, k - 1) : nocell, asdq, bmp1, bpp1);
                    } else {
#pragma unroll
                        for (int m = 0; m < MEQN; m++) bmm[m] = bpm[m] = bmp1[m] = bpp1[m] = 0.0;
                    }
                }
            });
        }
#pragma unroll
        for (int m = 0; m < MEQN; m++) F1[m] = 0.5 * cqxx[m];

        // complete cell k-2
        const int jc = k - 2;
        const bool row_out = (jc >= j0) && (jc < j1);
        const long long oidx = (long long)A.pitch * (jc + mbc - 1) + icl;
        const double hdtdy = 0.5 * dy_2;
        if (!TRANS) {
            if (row_out && col_out) {
#pragma unroll
                for (int m = 0; m < MEQN; m++) {
                    double qaddv = (0.0 - dy_2 * YS(SL::AP2 + m)) - dy_2 * amdq1[m];
                    double dF = F1[m] - YS(SL::F2 + m);
                    if (!CAPA)
                        A.qout[m * A.mstride + oidx] = (qm2[m] + qaddv) - dtdy * dF;
                    else
                        A.qout[m * A.mstride + oidx] = (qm2[m] + qaddv) - div1(dtdy * dF, cap_2);
                }
            }
        } else {
            double mainE[MEQN], qaddk[MEQN], qcv[MEQN];
#pragma unroll
            for (int m = 0; m < MEQN; m++) qcv[m] = 0.0;
            if (row_out) {
                if constexpr (RP::QCOR) { // apps/shallow-sphere/step2qcor.f:233
                    with_arith([&](auto &ar) {
                        RP::qcor(ar, A.rp, qm2, aux_cell(A, ic, k - 2), aux_cell(A, ic, k - 1), qcv);
                    });
                }
#pragma unroll
                for (int m = 0; m < MEQN; m++) {
                    double qaddv = (0.0 - dy_2 * YS(SL::AP2 + m)) - dy_2 * amdq1[m];
                    double dF = F1[m] - YS(SL::F2 + m);
                    double G1 = (0.0 - hdtdy * bmm[m]) - hdtdy * YS(SL::BMP2 + m);
                    double G2 = (0.0 - hdtdy * bpm[m]) - hdtdy * YS(SL::BPP2 + m);
                    if (!CAPA) mainE[m] = (qaddv - dtdy * dF - dtdx * (G2 - G1));
                    else mainE[m] = div1(dtdy * dF + dtdx * (G2 - G1), cap_2);
                    qaddk[m] = qaddv;
                    gs[(buf * 2 * MEQN + m) * NT + t] = G1;
                    gs[(buf * 2 * MEQN + MEQN + m) * NT + t] = G2;
                }
            }
            __syncthreads();
            if (row_out && col_out) {
#pragma unroll
                for (int m = 0; m < MEQN; m++) {
                    double G2l = gs[(buf * 2 * MEQN + MEQN + m) * NT + t - 1];
                    double G1r = gs[(buf * 2 * MEQN + m) * NT + t + 1];
                    double q = YQ(qxslot + m);
                    if (!CAPA) {
                        q = q + dtdx * G2l;
                        q = q + mainE[m];
                        q = q - dtdx * G1r;
                    } else { // step2.f:227-234 (+ step2qcor.f:244-245)
                        q = q + div1(dtdx * G2l, cap_2);
                        q = q + qaddk[m] - mainE[m];
                        if (RP::QCOR) q = q - div1(dtdy * qcv[m], cap_2);
                        q = q - div1(dtdx * G1r, cap_2);
                    }
                    A.qout[m * A.mstride + oidx] = q;
                }
            }
            buf ^= 1;
        }

        // shift the window
        dy_2 = dy_1; dy_1 = dy_k; cap_2 = cap_1; cap_1 = cap_k;
#pragma unroll
        for (int m = 0; m < MEQN; m++) {
            qm2[m] = qm1[m]; qm1[m] = qk[m];
            YS(SL::AP2 + m) = apdq1[m];
            YS(SL::F2 + m) = F1[m];
            if (TRANS) { YS(SL::BMP2 + m) = bmp1[m]; YS(SL::BPP2 + m) = bpp1[m]; }
        }
#pragma unroll
        for (int mw = 0; mw < MW; mw++) { sm1[mw] = s[mw]; norm1[mw] = normk[mw]; dot1[mw] = dotk[mw]; }
    }
#undef YS
#undef YQ
    if (ICFL) cfl = dtdy * __longlong_as_double((long long)smax);
    cfl_commit(cfl, A.cfl_bits);
}

// ---------------------------------------------------------------------------
// 1-D step (step1.f:4-142).  Same staging as the x-engine, one "row".
// Update order: q - dtdx*apdq(i), then - dtdx*amdq(i+1), then - dtdx*(f(i+1)-f(i)).
// ---------------------------------------------------------------------------
template <class RP, int NT, bool CAPA = false>
__global__ void __launch_bounds__(NT) step1_kernel(const SweepArgs A)
{
    constexpr int MEQN = RP::MEQN, MW = RP::MWAVES, NROE = RP::NROE;
    constexpr int NC = NT - 3;
    constexpr int QS = NT + 1;
    extern __shared__ double sm[];
    double *qs = sm;
    double *ws = qs + MEQN * QS;
    double *xs = ws + MEQN * MW * NT; // [2*MEQN][NT] amdq, f

    const int t = threadIdx.x;
    const int i0 = 1 + blockIdx.x * NC;
    const int ii = i0 - 1 + t;
    const int mbc = A.mbc;
    const int imax = A.mx + mbc;
    const int cload = min(i0 - 2 + t, imax) + mbc - 1;
    const int cload2 = min(i0 - 2 + NT, imax) + mbc - 1;
    const bool cell_ok = (t >= 1) && (t <= NC) && (ii <= A.mx);
    const bool iface_ok = (ii >= 1) && (ii <= A.mx + 1) && (t >= 1) && (t <= NT - 2);
    const bool order2 = (A.order != 1);
    double dtdx, dtdy_unused;
    load_dt(A, dtdx, dtdy_unused);
    double dtdx_l = dtdx; // cell ii-1
    if (CAPA) { // step1.f:62-73: dtdx(i) = dt / (dx * aux(mcapa, i))
        const double *cap = A.aux + (long long)(A.mcapa - 1) * A.amstride;
        const double dt = A.dt_dev ? __ldg(A.dt_dev) : A.dt;
        dtdx_l = div1(dt, A.dx * __ldg(&cap[min(max(ii - 1, 1 - mbc), imax) + mbc - 1]));
        dtdx = div1(dt, A.dx * __ldg(&cap[min(max(ii, 1 - mbc), imax) + mbc - 1]));
    }

#pragma unroll
    for (int m = 0; m < MEQN; m++) {
        qs[m * QS + t] = A.qin[m * A.mstride + cload];
        if (t == 0) qs[m * QS + NT] = A.qin[m * A.mstride + cload2];
    }
    __syncthreads();
    double l[MEQN], rr[MEQN];
#pragma unroll
    for (int m = 0; m < MEQN; m++) { l[m] = qs[m * QS + t]; rr[m] = qs[m * QS + t + 1]; }
    double wave[MEQN][MW], s[MW], amdq[MEQN], apdq[MEQN], roe[NROE];
    // step1.f:80: rp1 gets the one aux array as both auxl and auxr; 1-D arrays have no rows
    const AuxCell nocell{nullptr, 0};
    const AuxCell axl = (RP::MAUX > 0) ? AuxCell{A.aux + (min(max(ii - 1, 1 - mbc), imax) + mbc - 1), A.amstride} : nocell;
    const AuxCell axr = (RP::MAUX > 0) ? AuxCell{A.aux + (min(max(ii, 1 - mbc), imax) + mbc - 1), A.amstride} : nocell;
    with_arith([&](auto &ar) { RP::solve(ar, A.rp, l, rr, axl, axr, wave, s, amdq, apdq, roe); });
    double cfl = 0.0;
    if (iface_ok) {
#pragma unroll
        for (int mw = 0; mw < MW; mw++) cfl = dmax2(dmax2(cfl, dtdx * s[mw]), -dtdx_l * s[mw]);
    }
    if (order2) {
#pragma unroll
        for (int m = 0; m < MEQN; m++)
#pragma unroll
            for (int mw = 0; mw < MW; mw++)
                if (RP::nz(m, mw)) ws[(m * MW + mw) * NT + t] = wave[m][mw];
    }
    __syncthreads();
    double f[MEQN];
#pragma unroll
    for (int m = 0; m < MEQN; m++) f[m] = 0.0;
    if (order2 && t >= 1 && t <= NT - 2) {
        double wnorm2[MW], dotu[MW]; // dot product with the upwind neighbour only (see the x-engine)
#pragma unroll
        for (int mw = 0; mw < MW; mw++) {
            const int nb = (s[mw] > 0.0) ? t - 1 : t + 1;
            double n2 = 0.0, du = 0.0;
#pragma unroll
            for (int m = 0; m < MEQN; m++) {
                if (RP::nz(m, mw)) {
                    double w = wave[m][mw];
                    n2 = n2 + w * w;
                    du = du + ws[(m * MW + mw) * NT + nb] * w;
                }
            }
            wnorm2[mw] = n2; dotu[mw] = du;
        }
        with_arith([&](auto &ar) {
            if (!ar.FAST) {
#pragma unroll
                for (int m = 0; m < MEQN; m++)
#pragma unroll
                    for (int mw = 0; mw < MW; mw++)
                        if (RP::nz(m, mw)) wave[m][mw] = ws[(m * MW + mw) * NT + t];
            }
            limit_waves<RP>(ar, wave, s, wnorm2, dotu, dotu, A.mthlim);
        });
        // step1.f:121-128
        double dtdxave = 0.5 * (dtdx_l + dtdx);
#pragma unroll
        for (int m = 0; m < MEQN; m++)
#pragma unroll
            for (int mw = 0; mw < MW; mw++)
                if (RP::nz(m, mw))
                    f[m] = f[m] + 0.5 * lead_factor<RP>(s[mw]) * (1.0 - fabs(s[mw]) * dtdxave) * wave[m][mw];
    }
#pragma unroll
    for (int m = 0; m < MEQN; m++) {
        xs[m * NT + t] = amdq[m];
        xs[(MEQN + m) * NT + t] = f[m];
    }
    __syncthreads();
    if (cell_ok) {
#pragma unroll
        for (int m = 0; m < MEQN; m++) {
            double q = rr[m] - dtdx * apdq[m];
            q = q - dtdx * xs[m * NT + t + 1];
            if (order2) q = q - dtdx * (xs[(MEQN + m) * NT + t + 1] - f[m]);
            A.qout[m * A.mstride + (ii + mbc - 1)] = q;
        }
    }
    cfl_commit(cfl, A.cfl_bits);
}

// rp_kpp.cuh -- a USER-SUPPLIED Riemann solver, compiled into a variant of libclawb200.so with
//
//     python -m pyclaw_b200.build --user-rp examples/user_rp/rp_kpp.cuh --name kpp
//
// and bound from Python with  solver.rp = pyclaw.riemann.from_header(...)   (examples/kpp.py).
// This is the plugin seam of the reference -- any rpn2 / rpt2 Fortran file named by RP_SOURCE in
// the application's Makefile (Makefile.rules:1-26; apps/kpp/Makefile: rpn2_kpp.f + rpt2_dummy.f)
// -- re-hosted: a header with one struct template `RpUser<IXY>` exposing the interface every
// solver of rp.cuh has (MEQN, MWAVES, NROE, nz, solve, transverse).
//
// The KPP problem (Kurganov, Petrova, Popov 2007):  q_t + (sin q)_x + (cos q)_y = 0, a scalar
// law with a non-convex flux.  Two-wave HLL solver: the wave speeds are the exact extrema of
// f'(q) over the interval spanned by the two states (f' = cos q in x, g' = -sin q in y), the
// middle state follows from conservation.  The reference links the external
// clawpack/riemann rpn2_kpp.f (not in its tree, no golden file): parity for this solver is
// "GPU vs the same formulas in numpy" (tests/test_gpu_user_rp.py).
#pragma once

template <int IXY>
struct RpUser {
    static constexpr int MEQN = 1, MWAVES = 2, NROE = 1;
    static constexpr int X_MINB = 4, Y_MINB = 4; // CTAs per SM the sweeps are compiled for
    static constexpr int MAUX = 0;               // aux components the solver reads
    static constexpr bool QCOR = false;
    __host__ __device__ static constexpr bool nz(int, int) { return true; } // no structural zeros

    // does [a, b] contain p + 2 pi k for some integer k ?
    __device__ __forceinline__ static bool hits(double a, double b, double p)
    {
        const double twopi = 6.283185307179586476925286766559;
        return floor((b - p) / twopi) >= ceil((a - p) / twopi);
    }

    template <class AR>
    __device__ __forceinline__ static void solve(AR &ar, const RpParams &P, const double (&l)[1],
                                                 const double (&r)[1], const AuxCell &, const AuxCell &,
                                                 double (&wave)[1][2], double (&s)[2], double (&amdq)[1],
                                                 double (&apdq)[1], double (&roe)[NROE])
    {
        const double pi = 3.141592653589793238462643383279;
        const double ul = l[0], ur = r[0];
        const double a = fmin(ul, ur), b = fmax(ul, ur);
        double fl, fr, smin, smax;
        if (IXY == 1) { // f = sin q, f' = cos q: maximum 1 at 2 pi k, minimum -1 at pi + 2 pi k
            fl = sin(ul); fr = sin(ur);
            const double ca = cos(a), cb = cos(b);
            smax = hits(a, b, 0.0) ? 1.0 : fmax(ca, cb);
            smin = hits(a, b, pi) ? -1.0 : fmin(ca, cb);
        } else {        // g = cos q, g' = -sin q: maximum 1 at 3 pi / 2, minimum -1 at pi / 2
            fl = cos(ul); fr = cos(ur);
            const double sa = -sin(a), sb = -sin(b);
            smax = hits(a, b, 1.5 * pi) ? 1.0 : fmax(sa, sb);
            smin = hits(a, b, 0.5 * pi) ? -1.0 : fmin(sa, sb);
        }
        s[0] = smin;
        s[1] = smax;
        double um = ul; // equal speeds <=> equal states: no waves
        if (smax > smin) um = ar.div(smax * ur - smin * ul - (fr - fl), smax - smin);
        wave[0][0] = um - ul;
        wave[0][1] = ur - um;
        amdq[0] = fmin(smin, 0.0) * wave[0][0] + fmin(smax, 0.0) * wave[0][1];
        apdq[0] = fmax(smin, 0.0) * wave[0][0] + fmax(smax, 0.0) * wave[0][1];
        roe[0] = 0.0;
    }

    // rpt2_dummy.f: no transverse propagation
    template <class AR>
    __device__ __forceinline__ static void transverse(AR &, const RpParams &, const double (&)[NROE],
                                                      const double (&)[1], const AuxCell &, const AuxCell &,
                                                      const AuxCell &, const double (&)[1], double (&bm)[1],
                                                      double (&bp)[1])
    {
        bm[0] = 0.0;
        bp[0] = 0.0;
    }
};

// rp_advection_user.cuh -- the built-in 2-D advection solver (rpn2_advection.f / rpt2_advection.f)
// written as a user plugin: tests/test_gpu_user_rp.py checks that a run on this header is bit for
// bit the run on pyclaw.riemann.advection, i.e. the seam adds nothing of its own.
// Parameters: P.p[0] = u, P.p[1] = v   (from aux_global through from_header(param_names=["u","v"])).
#pragma once

template <int IXY>
struct RpUser {
    static constexpr int MEQN = 1, MWAVES = 1, NROE = 1;
    static constexpr int X_MINB = 6, Y_MINB = 6;
    static constexpr int MAUX = 0;
    static constexpr bool QCOR = false;
    __host__ __device__ static constexpr bool nz(int, int) { return true; }

    template <class AR>
    __device__ __forceinline__ static void solve(AR &, const RpParams &P, const double (&l)[1], const double (&r)[1],
                                                 const AuxCell &, const AuxCell &, double (&wave)[1][1],
                                                 double (&s)[1], double (&amdq)[1], double (&apdq)[1],
                                                 double (&roe)[NROE])
    {
        wave[0][0] = r[0] - l[0];
        s[0] = (IXY == 2) ? P.p[1] : P.p[0];
        amdq[0] = ((s[0] < 0.0) ? s[0] : 0.0) * wave[0][0];
        apdq[0] = ((s[0] > 0.0) ? s[0] : 0.0) * wave[0][0];
        roe[0] = 0.0;
    }

    template <class AR>
    __device__ __forceinline__ static void transverse(AR &, const RpParams &P, const double (&)[NROE],
                                                      const double (&)[1], const AuxCell &, const AuxCell &,
                                                      const AuxCell &, const double (&asdq)[1], double (&bm)[1],
                                                      double (&bp)[1])
    {
        const double stran = (IXY == 2) ? P.p[0] : P.p[1];
        bm[0] = ((stran < 0.0) ? stran : 0.0) * asdq[0];
        bp[0] = ((stran > 0.0) ? stran : 0.0) * asdq[0];
    }
};

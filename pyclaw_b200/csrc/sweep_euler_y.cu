// sweep_euler_y.cu -- y-engine instantiations for the Euler 5-wave Roe solver.
#include "launch.cuh"

int claw_y_euler(bool trans, const SweepArgs &A, cudaStream_t st)
{
    if (A.mcapa > 0)
        return trans ? launch_y<RpEuler5<2>, true, true>(A, st) : launch_y<RpEuler5<2>, false, true>(A, st);
    return trans ? launch_y<RpEuler5<2>, true>(A, st) : launch_y<RpEuler5<2>, false>(A, st);
}

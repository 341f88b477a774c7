// sweep_euler_x.cu -- x-engine instantiations for the Euler 5-wave Roe solver.
#include "launch.cuh"

int claw_x_euler(bool trans, const SweepArgs &A, cudaStream_t st)
{
    if (A.mcapa > 0)
        return trans ? launch_x<RpEuler5<1>, true, true>(A, st) : launch_x<RpEuler5<1>, false, true>(A, st);
    return trans ? launch_x<RpEuler5<1>, true>(A, st) : launch_x<RpEuler5<1>, false>(A, st);
}

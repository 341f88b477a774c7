// sweep_sphere.cu -- classic sweeps for shallow water on the sphere (16 aux, capacity function,
// step2qcor correction).
#include "launch.cuh"

int claw_x_sphere(bool trans, const SweepArgs &A, cudaStream_t st)
{
    return trans ? launch_x<RpSphere<1>, true, true>(A, st) : launch_x<RpSphere<1>, false, true>(A, st);
}
int claw_y_sphere(bool trans, const SweepArgs &A, cudaStream_t st)
{
    return trans ? launch_y<RpSphere<2>, true, true>(A, st) : launch_y<RpSphere<2>, false, true>(A, st);
}

// sweep_fused.cu -- the single-pass unsplit step (fused.cuh) for the light Riemann solvers.
#include "launch.cuh"
#include "fused.cuh"

constexpr int FNT = 128;

template <class RPX, class RPY>
static int launch_fused(SweepArgs A, cudaStream_t st)
{
    constexpr int NC = FNT - 4;
    constexpr int M = RPX::MEQN, MW = RPX::MWAVES;
    size_t smem = sizeof(double) * (2 * M * (FNT + 1) + M * MW * FNT + 4 * M * FNT + 4 * M * FNT);
    auto k = fused_step2_kernel<RPX, RPY, FNT>;
    CUDA_OK(set_smem(k, smem));
    const int ncols = A.ihi - A.ilo + 1, nrows = A.jhi - A.jlo + 1;
    const int strips = (ncols + NC - 1) / NC;
    A.rows_per_cta = pick_rows(nrows, strips);
    dim3 grid(strips, (nrows + A.rows_per_cta - 1) / A.rows_per_cta);
    k<<<grid, FNT, smem, st>>>(A);
    CUDA_OK(cudaGetLastError());
    return 0;
}

bool claw_fused_available(int rp_id, const SweepArgs &A)
{
    if (A.mcapa > 0 || A.trans < 0) return false;
    return rp_id == CLAWB200_RP_ACOUSTICS || rp_id == CLAWB200_RP_ADVECTION || rp_id == CLAWB200_RP_SHALLOW;
}

int claw_fused(int rp_id, const SweepArgs &A, cudaStream_t st)
{
    switch (rp_id) {
    case CLAWB200_RP_ACOUSTICS: return launch_fused<RpAcoustics<2, 1>, RpAcoustics<2, 2>>(A, st);
    case CLAWB200_RP_ADVECTION: return launch_fused<RpAdvection<2, 1>, RpAdvection<2, 2>>(A, st);
    case CLAWB200_RP_SHALLOW: return launch_fused<RpShallow<1>, RpShallow<2>>(A, st);
    default: return fail(CLAWB200_ERR_UNSUPPORTED, "the single-pass step is not compiled for this solver");
    }
}

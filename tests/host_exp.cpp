// Host build of the device exp (sparsergps_b200/csrc/kmath.cuh) so its accuracy can be tested without a GPU.
#define SRGP_KMATH_HOST 1
#include "../sparsergps_b200/csrc/kmath.cuh"
extern "C" void host_exp_nonpos(const double *x, long n, double *out)
{
    for (long i = 0; i < n; i++) out[i] = srgp::exp_nonpos(x[i]);
}

// todo.cu -- entry points declared in include/srgp.h whose kernels are not built yet.
#include "common.cuh"
using namespace srgp;
#define SRGP_TODO(name) do { set_error(name ": not implemented yet"); return SRGP_ERR_STATE; } while (0)
extern "C" {
int srgp_laplace_newton(srgp_ctx *, int, int, const double *, int64_t, const double *, double, const double *, double, double, double, int, double, double *, double *, int *, double *, double *, double *) { SRGP_TODO("srgp_laplace_newton"); }
int srgp_laplace_grad(srgp_ctx *, int, int, const double *, int64_t, double, const double *, double, double, double, const double *, double *) { SRGP_TODO("srgp_laplace_grad"); }
}

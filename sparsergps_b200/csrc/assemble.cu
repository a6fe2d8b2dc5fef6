// assemble.cu -- K1 (covariance assembly) and K2 (d Sigma / d log theta), materialising, column-major output.
//
// Replaces the double loops of make_cov_matC / make_cov_mat_ardC (reference src/covariance_functionsC.cpp:72-252)
// and dsig_dthetaC / dsig_dtheta_ardC (src/covariance_function_derivativesC.cpp:307-722).
//
// Layout: one thread per output row i (its d coordinates live in registers), a CTA covers 256 rows x 32
// columns, the 32 x_pred rows of the tile are staged in shared memory and read as broadcasts.  Consecutive
// threads write consecutive rows of a column => 256-byte coalesced streaming stores.  The kernel is bound by
// the HBM write of 8*n1*n2 bytes (plus the FP64 pipe for exp); nothing is read twice.
#include "common.cuh"
#include "fastexp.cuh"
#include "gauss.cuh"

namespace srgp {

enum { MODE_COV = 0, MODE_DSIGMA = 1, MODE_DL = 2, MODE_DLC = 3, MODE_DTAU = 4, MODE_ZERO = 5 };

struct AsmParams {
    double sigma2;           // sigma^2
    double c_exp;            // sqexp: -1/(2 l^2); exp: -1/l; ard: unused
    double inv_l2;           // sqexp: 1/l^2; exp: 1/l
    double tau2_delta;       // tau^2 + delta (self-covariance diagonal)
    double two_tau2;         // 2 tau^2
    int comp;                // ard: 0-based component for MODE_DLC
    double invl[SRGP_MAX_D]; // ard: 1/l_c
};

constexpr int ASM_ROWS = 256;
constexpr int ASM_COLS = 32;

template <int KT, int MODE, int DT>
__global__ void __launch_bounds__(ASM_ROWS)
assemble_kernel(const double *__restrict__ x, int64_t n1, const double *__restrict__ xb, int64_t n2, int d_rt,
                AsmParams p, int self, double *__restrict__ out, int64_t ldo)
{
    extern __shared__ double su[];  // [ASM_COLS][d]
    __shared__ double etab[EXP_TAB_DOUBLES];
    exp_tab_load(etab, threadIdx.x, ASM_ROWS);
    const int d = DT > 0 ? DT : d_rt;
    const int64_t i = (int64_t)blockIdx.x * ASM_ROWS + threadIdx.x;
    double xi[DT > 0 ? DT : 1];
    if (DT > 0 && i < n1) {
#pragma unroll
        for (int c = 0; c < DT; c++) xi[c] = x[i + n1 * c];
    }
    const int64_t col_tiles = (n2 + ASM_COLS - 1) / ASM_COLS;
    for (int64_t jt = blockIdx.y; jt < col_tiles; jt += gridDim.y) {
        const int64_t j0 = jt * ASM_COLS;
        __syncthreads();
        for (int t = threadIdx.x; t < ASM_COLS * d; t += ASM_ROWS) {
            int jj = t / d, c = t - jj * d;
            int64_t j = j0 + jj;
            su[t] = (j < n2) ? xb[j + n2 * c] : 0.0;
        }
        __syncthreads();
        if (i >= n1) continue;
        const int jmax = (int)((n2 - j0) < ASM_COLS ? (n2 - j0) : ASM_COLS);
#pragma unroll 4
        for (int jj = 0; jj < jmax; jj++) {
            const double *u = su + jj * d;
            double acc = 0.0;   // sqexp: sum d^2; ard: sum (d/l)^2; exp cov: sum |d|
            double dc2 = 0.0;   // ard MODE_DLC: (d_c / l_c)^2
            bool alleq = true;
            if (DT > 0) {
#pragma unroll
                for (int c = 0; c < DT; c++) {
                    double a = xi[c], b = u[c];
                    double df = a - b;
                    if (MODE == MODE_DTAU) alleq = alleq && (a == b);
                    if (KT == SRGP_ARD) {
                        double t = df * p.invl[c];
                        if (MODE == MODE_DLC && c == p.comp) dc2 = t * t;
                        acc = fma(t, t, acc);
                    } else if (KT == SRGP_EXP && MODE == MODE_COV) {
                        acc += fabs(df);
                    } else {
                        acc = fma(df, df, acc);
                    }
                }
            } else {
                for (int c = 0; c < d; c++) {
                    double a = x[i + n1 * c], b = u[c];
                    double df = a - b;
                    if (MODE == MODE_DTAU) alleq = alleq && (a == b);
                    if (KT == SRGP_ARD) {
                        double t = df * p.invl[c];
                        if (MODE == MODE_DLC && c == p.comp) dc2 = t * t;
                        acc = fma(t, t, acc);
                    } else if (KT == SRGP_EXP && MODE == MODE_COV) {
                        acc += fabs(df);
                    } else {
                        acc = fma(df, df, acc);
                    }
                }
            }
            double v;
            if (MODE == MODE_DTAU) {
                v = alleq ? p.two_tau2 : 0.0;
            } else if (MODE == MODE_ZERO) {
                v = 0.0;
            } else if (KT == SRGP_ARD) {
                double k = p.sigma2 * exp_tab(-0.5 * acc, etab);
                v = (MODE == MODE_COV) ? k : (MODE == MODE_DSIGMA) ? 2.0 * k : k * dc2;
            } else if (KT == SRGP_SQEXP) {
                double k = p.sigma2 * exp_tab(p.c_exp * acc, etab);
                v = (MODE == MODE_COV) ? k : (MODE == MODE_DSIGMA) ? 2.0 * k : k * (acc * p.inv_l2);
            } else {  // SRGP_EXP: covariance uses the L1 distance, derivatives the L2 distance (quirk Q8)
                if (MODE == MODE_COV) {
                    v = p.sigma2 * exp_tab(p.c_exp * acc, etab);
                } else {
                    double r = sqrt(acc);
                    double k = p.sigma2 * exp_tab(p.c_exp * r, etab);
                    v = (MODE == MODE_DSIGMA) ? 2.0 * k : k * (r * p.inv_l2);
                }
            }
            const int64_t j = j0 + jj;
            if (MODE == MODE_COV && self && i == j) v = v + p.tau2_delta;
            __stcs(out + i + ldo * j, v);
        }
    }
}

template <int KT, int MODE>
static int launch_d(srgp_ctx *ctx, cudaStream_t st, const double *x, int64_t n1, const double *xb, int64_t n2,
                    int d, const AsmParams &p, int self, double *out, int64_t ldo)
{
    dim3 block(ASM_ROWS);
    int64_t col_tiles = ceil_div(n2, ASM_COLS);
    dim3 grid((unsigned)ceil_div(n1, ASM_ROWS), (unsigned)(col_tiles < 65535 ? col_tiles : 65535));
    size_t smem = sizeof(double) * ASM_COLS * d;
    KernelScope ks(ctx, SRGP_PROF_ASSEMBLE, st);
#define SRGP_ASM_CASE(D)                                                                                  \
    case D:                                                                                               \
        assemble_kernel<KT, MODE, D><<<grid, block, smem, st>>>(x, n1, xb, n2, d, p, self, out, ldo);       \
        break;
    switch (d) {
        SRGP_ASM_CASE(1)
        SRGP_ASM_CASE(2)
        SRGP_ASM_CASE(3)
        SRGP_ASM_CASE(4)
        SRGP_ASM_CASE(5)
        SRGP_ASM_CASE(6)
        SRGP_ASM_CASE(7)
        SRGP_ASM_CASE(8)
    default:
        assemble_kernel<KT, MODE, 0><<<grid, block, smem, st>>>(x, n1, xb, n2, d, p, self, out, ldo);
    }
#undef SRGP_ASM_CASE
    SRGP_LAUNCH_CHECK();
    return SRGP_OK;
}

template <int KT>
static int launch_mode(srgp_ctx *ctx, cudaStream_t st, int mode, const double *x, int64_t n1, const double *xb,
                       int64_t n2, int d, const AsmParams &p, int self, double *out, int64_t ldo)
{
    switch (mode) {
    case MODE_COV: return launch_d<KT, MODE_COV>(ctx, st, x, n1, xb, n2, d, p, self, out, ldo);
    case MODE_DSIGMA: return launch_d<KT, MODE_DSIGMA>(ctx, st, x, n1, xb, n2, d, p, self, out, ldo);
    case MODE_DL: return launch_d<KT, MODE_DL>(ctx, st, x, n1, xb, n2, d, p, self, out, ldo);
    case MODE_DLC: return launch_d<KT, MODE_DLC>(ctx, st, x, n1, xb, n2, d, p, self, out, ldo);
    case MODE_DTAU: return launch_d<KT, MODE_DTAU>(ctx, st, x, n1, xb, n2, d, p, self, out, ldo);
    default: return launch_d<KT, MODE_ZERO>(ctx, st, x, n1, xb, n2, d, p, self, out, ldo);
    }
}

static int check_common(srgp_ctx *ctx, int kernel, const double *x, int64_t n1, const double *x_pred, int64_t n2,
                        int d, const double *l, const double *out)
{
    if (!ctx || !x || !l || !out || n1 <= 0 || d <= 0 || (x_pred && n2 <= 0)) {
        set_error("bad argument (null pointer or non-positive size)");
        return SRGP_ERR_ARG;
    }
    if (d > SRGP_MAX_D) {
        set_error("input dimension d=%d exceeds SRGP_MAX_D=%d", d, SRGP_MAX_D);
        return SRGP_ERR_ARG;
    }
    if (kernel != SRGP_SQEXP && kernel != SRGP_EXP && kernel != SRGP_ARD) {
        set_error("Error: invalid covariance function");
        return SRGP_ERR_UNKNOWN_KERNEL;
    }
    return SRGP_OK;
}

static void fill_params(AsmParams &p, int kernel, int d, double sigma, const double *l, double tau, double delta,
                        int comp0)
{
    p.sigma2 = sigma * sigma;
    p.tau2_delta = tau * tau + delta;
    p.two_tau2 = 2.0 * tau * tau;
    p.comp = comp0;
    p.c_exp = 0.0;
    p.inv_l2 = 0.0;
    for (int c = 0; c < SRGP_MAX_D; c++) p.invl[c] = 0.0;
    if (kernel == SRGP_SQEXP) {
        p.c_exp = -1.0 / (2.0 * l[0] * l[0]);
        p.inv_l2 = 1.0 / (l[0] * l[0]);
    } else if (kernel == SRGP_EXP) {
        p.c_exp = -1.0 / l[0];
        p.inv_l2 = 1.0 / l[0];
    } else {
        for (int c = 0; c < d; c++) p.invl[c] = 1.0 / l[c];
    }
}

int assemble_dev(srgp_ctx *ctx, int kernel, int mode, int comp0, const double *x_dev, int64_t n1,
                 const double *xp_dev, int64_t n2, int d, double sigma, const double *l, double tau, double delta,
                 double *out_dev)
{
    AsmParams p;
    fill_params(p, kernel, d, sigma, l, tau, delta, comp0);
    const int self = (xp_dev == nullptr);
    const double *xb = self ? x_dev : xp_dev;
    const int64_t nb = self ? n1 : n2;
    if (kernel == SRGP_SQEXP)
        return launch_mode<SRGP_SQEXP>(ctx, ctx->stream, mode, x_dev, n1, xb, nb, d, p, self, out_dev, n1);
    if (kernel == SRGP_EXP)
        return launch_mode<SRGP_EXP>(ctx, ctx->stream, mode, x_dev, n1, xb, nb, d, p, self, out_dev, n1);
    return launch_mode<SRGP_ARD>(ctx, ctx->stream, mode, x_dev, n1, xb, nb, d, p, self, out_dev, n1);
}

// Self-covariance K(x, x) + nugget * I with an explicit output leading dimension (the m x m stage keeps its
// matrices padded to ld = round_up(m, 128)).
int assemble_dev_ld(srgp_ctx *ctx, cudaStream_t st, int kernel, const double *x_dev, int64_t n1, int d, double sigma,
                    const double *l, double nugget, double *out_dev, int64_t ldo)
{
    AsmParams p;
    fill_params(p, kernel, d, sigma, l, 0.0, nugget, 0);
    if (kernel == SRGP_SQEXP)
        return launch_mode<SRGP_SQEXP>(ctx, st, MODE_COV, x_dev, n1, x_dev, n1, d, p, 1, out_dev, ldo);
    if (kernel == SRGP_EXP)
        return launch_mode<SRGP_EXP>(ctx, st, MODE_COV, x_dev, n1, x_dev, n1, d, p, 1, out_dev, ldo);
    return launch_mode<SRGP_ARD>(ctx, st, MODE_COV, x_dev, n1, x_dev, n1, d, p, 1, out_dev, ldo);
}

// par -> kernel mode, with the reference's dispatch rules.
static int par_to_mode(int kernel, int par, int comp0, int d, bool cross, int *mode)
{
    if (kernel == SRGP_EXP && cross && par != SRGP_PAR_SIGMA && par != SRGP_PAR_L) {
        *mode = MODE_ZERO;  // quirk Q9: `return mat` at derivativesC.cpp:520 precedes the tau branch
        return SRGP_OK;
    }
    if (par == SRGP_PAR_SIGMA) { *mode = MODE_DSIGMA; return SRGP_OK; }
    if (par == SRGP_PAR_TAU) { *mode = MODE_DTAU; return SRGP_OK; }
    if (kernel == SRGP_ARD && par == SRGP_PAR_LC && comp0 >= 0 && comp0 < d) { *mode = MODE_DLC; return SRGP_OK; }
    if (kernel != SRGP_ARD && par == SRGP_PAR_L) { *mode = MODE_DL; return SRGP_OK; }
    set_error("Error: invalid parameter name for chosen covariance function");
    return SRGP_ERR_UNKNOWN_PAR;
}

static int host_call(srgp_ctx *ctx, int kernel, int mode, int comp0, const double *x, int64_t n1,
                     const double *x_pred, int64_t n2, int d, double sigma, const double *l, double tau,
                     double delta, double *out)
{
    SRGP_TRY(use_device(ctx));
    const int64_t nb = x_pred ? n2 : n1;
    SRGP_TRY(ctx->in_x.reserve(sizeof(double) * n1 * d));
    SRGP_CUDA(cudaMemcpyAsync(ctx->in_x.p, x, sizeof(double) * n1 * d, cudaMemcpyHostToDevice, ctx->stream));
    const double *xp_dev = nullptr;
    if (x_pred) {
        SRGP_TRY(ctx->in_xp.reserve(sizeof(double) * n2 * d));
        SRGP_CUDA(cudaMemcpyAsync(ctx->in_xp.p, x_pred, sizeof(double) * n2 * d, cudaMemcpyHostToDevice,
                                  ctx->stream));
        xp_dev = ctx->in_xp.d();
    }
    SRGP_TRY(ctx->out_mat.reserve(sizeof(double) * (size_t)n1 * (size_t)nb));
    SRGP_TRY(assemble_dev(ctx, kernel, mode, comp0, ctx->in_x.d(), n1, xp_dev, n2, d, sigma, l, tau, delta,
                          ctx->out_mat.d()));
    SRGP_CUDA(cudaMemcpyAsync(out, ctx->out_mat.p, sizeof(double) * (size_t)n1 * (size_t)nb,
                              cudaMemcpyDeviceToHost, ctx->stream));
    SRGP_CUDA(cudaStreamSynchronize(ctx->stream));
    return SRGP_OK;
}

}  // namespace srgp

using namespace srgp;

extern "C" int srgp_make_cov_mat(srgp_ctx *ctx, int kernel, const double *x, int64_t n1, const double *x_pred,
                                 int64_t n2, int d, double sigma, const double *l, double tau, double delta,
                                 double *out)
{
    SRGP_TRY(check_common(ctx, kernel, x, n1, x_pred, n2, d, l, out));
    return host_call(ctx, kernel, MODE_COV, 0, x, n1, x_pred, n2, d, sigma, l, tau, delta, out);
}

extern "C" int srgp_dsig_dtheta(srgp_ctx *ctx, int kernel, int par, int comp0, const double *x, int64_t n1,
                                const double *x_pred, int64_t n2, int d, double sigma, const double *l, double tau,
                                double *out)
{
    SRGP_TRY(check_common(ctx, kernel, x, n1, x_pred, n2, d, l, out));
    int mode;
    SRGP_TRY(par_to_mode(kernel, par, comp0, d, x_pred != nullptr, &mode));
    return host_call(ctx, kernel, mode, comp0, x, n1, x_pred, n2, d, sigma, l, tau, 0.0, out);
}

extern "C" int srgp_make_cov_mat_dev(srgp_ctx *ctx, int kernel, const double *x_dev, int64_t n1,
                                     const double *x_pred_dev, int64_t n2, int d, double sigma, const double *l,
                                     double tau, double delta, double *out_dev)
{
    SRGP_TRY(check_common(ctx, kernel, x_dev, n1, x_pred_dev, n2, d, l, out_dev));
    SRGP_TRY(use_device(ctx));
    return assemble_dev(ctx, kernel, MODE_COV, 0, x_dev, n1, x_pred_dev, n2, d, sigma, l, tau, delta, out_dev);
}

extern "C" int srgp_dsig_dtheta_dev(srgp_ctx *ctx, int kernel, int par, int comp0, const double *x_dev, int64_t n1,
                                    const double *x_pred_dev, int64_t n2, int d, double sigma, const double *l,
                                    double tau, double *out_dev)
{
    SRGP_TRY(check_common(ctx, kernel, x_dev, n1, x_pred_dev, n2, d, l, out_dev));
    int mode;
    SRGP_TRY(par_to_mode(kernel, par, comp0, d, x_pred_dev != nullptr, &mode));
    SRGP_TRY(use_device(ctx));
    return assemble_dev(ctx, kernel, mode, comp0, x_dev, n1, x_pred_dev, n2, d, sigma, l, tau, 0.0, out_dev);
}

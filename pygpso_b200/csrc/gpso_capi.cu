// gpso_b200 C ABI (see include/gpso_b200.h): handle, device memory, and the launch sequences of the three pipelines
//   fit     : scale -> Gram -> blocked Cholesky -> L^-1 (recursive doubling) -> K_y^-1 -> alpha -> LML + gradient
//   predict : per candidate window: cross-covariance (+mean) -> triangular product + column sum of squares -> finalise
//   explore : leaf generation on the device -> predict -> UCB arg-max
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC
#include "../../include/gpso_b200.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "common.cuh"
#include "gemm_core.cuh"
#include "kern_cov.cuh"
#include "kern_dense.cuh"
#include "kern_leaves.cuh"
#include "kern_predict.cuh"

using namespace gpso;

static thread_local std::string g_last_error;

static int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

#define CU_TRY(call)                                                                                          \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess) {                                                                             \
            char buf__[512];                                                                                  \
            snprintf(buf__, sizeof buf__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return fail(e__ == cudaErrorMemoryAllocation ? GPSO_E_NOMEM : GPSO_E_CUDA, buf__);                \
        }                                                                                                     \
    } while (0)

#define GP_TRY(call)             \
    do {                         \
        int rc__ = (call);       \
        if (rc__ != 0) return rc__; \
    } while (0)

namespace {

constexpr int MAX_LS = LEAF_MAXD;  // maximum input dimension supported
constexpr double NOISE_FLOOR = 1.0e-6;
constexpr long long WINDOW_BYTES = 2LL << 30;  // rolling cross-covariance window budget (2 GiB)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes, bool zero = false) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            char b[256];
            snprintf(b, sizeof b, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
            return fail(GPSO_E_NOMEM, b);
        }
        cap = bytes;
        if (zero) cudaMemset(p, 0, bytes);
        return 0;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T* as() const {
        return reinterpret_cast<T*>(p);
    }
};

double softplus(double u) { return u > 0 ? u + log1p(exp(-u)) : log1p(exp(u)); }
double sigmoid(double u) { return 1.0 / (1.0 + exp(-u)); }

}  // namespace

struct gpso_handle {
    int device = 0, kernel_id = KERNEL_MATERN52, ard = 0, mean_id = GPSO_MEAN_CONSTANT;
    int N = 0, d = 0, Np = 0, nb = 0;
    int nsm = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_used[2] = {nullptr, nullptr}, ev_t0 = nullptr, ev_t1 = nullptr;
    bool used_pending[2] = {false, false};
    // data + fitted state
    DevBuf X, y, Xs, ls, alpha;
    DevBuf K, Linv, LinvT, T, Kinv;
    DevBuf resid, a, logdet, scalars, gpart, gout, info, counter;
    // predict workspaces
    DevBuf KsT, part, wmean, blockbest, running, cand[2], leaves, omean, ovar;
    long long window_override = 0;
    // host copies of the hyper-parameters in force
    double ls_host[MAX_LS] = {0}, variance = 1.0, noise = 1.0, c0 = 0.0;
    bool have_data = false, factorized = false;
    double factor_nlml = 0.0;
    long long launches = 0;
    double last_ms[4] = {0, 0, 0, 0};
    // optional per-stage profiling: 4 events per window on the launch stream
    bool profile = false;
    std::vector<cudaEvent_t> prof_events;
    size_t prof_used = 0;
    long long last_windows = 0;
    int n_ls() const { return ard ? d : 1; }
    int n_params() const { return n_ls() + 2 + (mean_id == GPSO_MEAN_CONSTANT ? 1 : 0); }
};

// ---------------------------------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------------------------------
static int set_device(gpso_handle* h) {
    CU_TRY(cudaSetDevice(h->device));
    return 0;
}

static int check_launch(gpso_handle* h, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        char b[256];
        snprintf(b, sizeof b, "launch of %s failed: %s", what, cudaGetErrorString(e));
        return fail(GPSO_E_CUDA, b);
    }
    h->launches++;
    return 0;
}

template <int KID>
static void launch_gram(gpso_handle* h, cudaStream_t st) {
    dim3 grid(h->Np / CT, h->Np / CT);
    gram_kernel<KID><<<grid, 256, 0, st>>>(h->Xs.as<double>(), h->N, h->d, h->Np, h->variance, h->noise, h->K.as<double>());
}

template <int KID>
static void launch_crosscov(gpso_handle* h, cudaStream_t st, const double* Xc, long long Mw, long long Mw_pad) {
    size_t sm = (size_t)(XG * h->d + 8 * XG) * sizeof(double);
    crosscov_kernel<KID><<<(unsigned)(Mw_pad / XG), 256, sm, st>>>(Xc, Mw, h->d, h->ls.as<double>(), h->n_ls(), h->Xs.as<double>(),
                                                                  h->alpha.as<double>(), h->N, h->Np, h->variance, h->c0,
                                                                  h->KsT.as<double>(), h->wmean.as<double>());
}

template <int KID>
static void launch_grad(gpso_handle* h, cudaStream_t st, int nblk, int stride) {
    if (!h->ard) {
        lml_grad_kernel<KID, false><<<nblk, 256, 0, st>>>(h->Xs.as<double>(), h->alpha.as<double>(), h->Kinv.as<double>(), h->N,
                                                         h->d, h->Np, h->variance, 0, 1, h->gpart.as<double>(), stride);
        h->launches++;
    } else {
        for (int dim0 = 0; dim0 < h->d; dim0 += GRAD_DCH) {
            int nd = std::min(GRAD_DCH, h->d - dim0);
            lml_grad_kernel<KID, true><<<nblk, 256, 0, st>>>(h->Xs.as<double>(), h->alpha.as<double>(), h->Kinv.as<double>(),
                                                            h->N, h->d, h->Np, h->variance, dim0, nd, h->gpart.as<double>(),
                                                            stride);
            h->launches++;
        }
    }
}

#define DISPATCH_KID(h, fn, ...)                                           \
    switch ((h)->kernel_id) {                                              \
        case KERNEL_MATERN12: fn<KERNEL_MATERN12>(__VA_ARGS__); break;     \
        case KERNEL_MATERN32: fn<KERNEL_MATERN32>(__VA_ARGS__); break;     \
        case KERNEL_MATERN52: fn<KERNEL_MATERN52>(__VA_ARGS__); break;     \
        default: fn<KERNEL_SE>(__VA_ARGS__); break;                        \
    }

static int configure_kernels() {
    CU_TRY(cudaFuncSetAttribute(dense_gemm_kernel<MODE_CHOL_PANEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    CU_TRY(cudaFuncSetAttribute(dense_gemm_kernel<MODE_CHOL_TRAIL>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    CU_TRY(cudaFuncSetAttribute(dense_gemm_kernel<MODE_TRTRI_XT>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    CU_TRY(cudaFuncSetAttribute(dense_gemm_kernel<MODE_TRTRI_Y>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    CU_TRY(cudaFuncSetAttribute(dense_gemm_kernel<MODE_LAUUM>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    CU_TRY(cudaFuncSetAttribute(predict_trmm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    CU_TRY(cudaFuncSetAttribute(diag_factor_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM_BYTES));
    return 0;
}

// (re)allocate everything that depends on the problem shape
static int ensure_shape(gpso_handle* h, int N, int d) {
    int Np = ((N + TB - 1) / TB) * TB;
    size_t mat = (size_t)Np * Np * sizeof(double);
    GP_TRY(h->X.ensure((size_t)N * d * sizeof(double)));
    GP_TRY(h->y.ensure((size_t)Np * sizeof(double)));
    GP_TRY(h->Xs.ensure((size_t)d * Np * sizeof(double)));
    GP_TRY(h->ls.ensure(MAX_LS * sizeof(double)));
    GP_TRY(h->alpha.ensure((size_t)Np * sizeof(double)));
    GP_TRY(h->resid.ensure((size_t)Np * sizeof(double)));
    GP_TRY(h->a.ensure((size_t)Np * sizeof(double)));
    GP_TRY(h->K.ensure(mat, true));
    GP_TRY(h->Linv.ensure(mat, true));
    GP_TRY(h->LinvT.ensure(mat, true));
    GP_TRY(h->logdet.ensure((size_t)(Np / TB) * sizeof(double)));
    GP_TRY(h->scalars.ensure(16 * sizeof(double)));
    GP_TRY(h->info.ensure(sizeof(int)));
    GP_TRY(h->counter.ensure(sizeof(int)));
    GP_TRY(h->running.ensure(sizeof(BestRec)));
    h->N = N;
    h->d = d;
    h->Np = Np;
    h->nb = Np / TB;
    return 0;
}

static int upload_lengthscales(gpso_handle* h, cudaStream_t st) {
    CU_TRY(cudaMemcpyAsync(h->ls.p, h->ls_host, sizeof(double) * h->n_ls(), cudaMemcpyHostToDevice, st));
    return 0;
}

// Gram -> Cholesky -> inverse factor -> [K_y^-1] -> a, alpha -> scalars.  Uses h->ls_host/variance/noise/c0.
static int factor_pipeline(gpso_handle* h, cudaStream_t st, bool need_kinv) {
    const int Np = h->Np, nb = h->nb;
    GP_TRY(upload_lengthscales(h, st));
    CU_TRY(cudaMemsetAsync(h->info.p, 0, sizeof(int), st));
    scale_inputs_kernel<<<(Np + 255) / 256, 256, 0, st>>>(h->X.as<double>(), h->ls.as<double>(), h->n_ls(), h->N, h->d, Np,
                                                          h->Xs.as<double>());
    GP_TRY(check_launch(h, "scale_inputs"));
    DISPATCH_KID(h, launch_gram, h, st);
    GP_TRY(check_launch(h, "gram"));

    DenseParams P;
    P.K = h->K.as<double>();
    P.Linv = h->Linv.as<double>();
    P.LinvT = h->LinvT.as<double>();
    P.T = nullptr;
    P.Kinv = nullptr;
    P.Np = Np;
    P.nb = nb;
    P.p = 0;
    P.s = 0;
    for (int p = 0; p < nb; p++) {
        diag_factor_inverse_kernel<<<1, 256, DIAG_SMEM_BYTES, st>>>(P.K, P.Linv, P.LinvT, Np, p, h->N, h->logdet.as<double>(),
                                                                    h->info.as<int>());
        GP_TRY(check_launch(h, "diag_factor_inverse"));
        int nt = nb - 1 - p;
        if (nt > 0) {
            P.p = p;
            dense_gemm_kernel<MODE_CHOL_PANEL><<<nt, GTHREADS, GEMM_SMEM_BYTES, st>>>(P);
            GP_TRY(check_launch(h, "chol_panel"));
            dense_gemm_kernel<MODE_CHOL_TRAIL><<<nt*(nt + 1) / 2, GTHREADS, GEMM_SMEM_BYTES, st>>>(P);
            GP_TRY(check_launch(h, "chol_trailing"));
        }
    }
    if (nb > 1) {
        GP_TRY(h->T.ensure((size_t)Np * Np * sizeof(double), true));
        P.T = h->T.as<double>();
        for (int s = 1; s < nb; s *= 2) {
            int cnt = 0;
            for (int q = 0; 2 * q * s < nb; q++) cnt += s * trtri_pair_vtiles(nb, s, q);
            if (cnt == 0) continue;
            P.s = s;
            dense_gemm_kernel<MODE_TRTRI_XT><<<cnt, GTHREADS, GEMM_SMEM_BYTES, st>>>(P);
            GP_TRY(check_launch(h, "trtri_xt"));
            dense_gemm_kernel<MODE_TRTRI_Y><<<cnt, GTHREADS, GEMM_SMEM_BYTES, st>>>(P);
            GP_TRY(check_launch(h, "trtri_y"));
        }
    }
    if (need_kinv) {
        GP_TRY(h->Kinv.ensure((size_t)Np * Np * sizeof(double), true));
        P.Kinv = h->Kinv.as<double>();
        dense_gemm_kernel<MODE_LAUUM><<<nb*(nb + 1) / 2, GTHREADS, GEMM_SMEM_BYTES, st>>>(P);
        GP_TRY(check_launch(h, "lauum"));
    }
    residual_kernel<<<(Np + 255) / 256, 256, 0, st>>>(h->y.as<double>(), h->c0, h->N, Np, h->resid.as<double>());
    GP_TRY(check_launch(h, "residual"));
    tri_matvec_kernel<true><<<(Np + 7) / 8, 256, 0, st>>>(P.Linv, h->resid.as<double>(), Np, h->a.as<double>());
    GP_TRY(check_launch(h, "trimv_lower"));
    tri_matvec_kernel<false><<<(Np + 7) / 8, 256, 0, st>>>(P.LinvT, h->a.as<double>(), Np, h->alpha.as<double>());
    GP_TRY(check_launch(h, "trimv_upper"));
    lml_scalars_kernel<<<1, 256, 0, st>>>(h->a.as<double>(), h->alpha.as<double>(), h->logdet.as<double>(), h->N, nb,
                                          h->scalars.as<double>());
    GP_TRY(check_launch(h, "lml_scalars"));
    return 0;
}

static double nlml_from_scalars(const gpso_handle* h, const double* sc) {
    return 0.5 * sc[0] + 0.5 * h->N * log(2.0 * M_PI) + sc[1];
}

// ---------------------------------------------------------------------------------------------------------------------
// library
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int gpso_version(void) { return 1; }
extern "C" const char* gpso_last_error(void) { return g_last_error.c_str(); }
extern "C" int gpso_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(GPSO_E_NOGPU, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    return n;
}

extern "C" int gpso_create(int device, int kernel_id, int ard, int mean_id, gpso_handle** out) {
    if (!out) return fail(GPSO_E_BADARG, "gpso_create: out is null");
    *out = nullptr;
    if (kernel_id < 0 || kernel_id > 3) return fail(GPSO_E_BADARG, "gpso_create: unknown kernel id");
    if (mean_id != GPSO_MEAN_ZERO && mean_id != GPSO_MEAN_CONSTANT) return fail(GPSO_E_BADARG, "gpso_create: unknown mean id");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return fail(GPSO_E_NOGPU, "gpso_create: no CUDA device available");
    if (device < 0 || device >= n) return fail(GPSO_E_BADARG, "gpso_create: device index out of range");
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        char b[256];
        snprintf(b, sizeof b, "gpso_create: device %d (%s, sm_%d%d) is not a Blackwell sm_100 GPU", device, prop.name, prop.major,
                 prop.minor);
        return fail(GPSO_E_NOGPU, b);
    }
    CU_TRY(cudaSetDevice(device));
    gpso_handle* h = new gpso_handle();
    h->device = device;
    h->kernel_id = kernel_id;
    h->ard = ard ? 1 : 0;
    h->mean_id = mean_id;
    h->nsm = prop.multiProcessorCount;
    CU_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        CU_TRY(cudaEventCreateWithFlags(&h->ev_copy[i], cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&h->ev_used[i], cudaEventDisableTiming));
    }
    CU_TRY(cudaEventCreate(&h->ev_t0));
    CU_TRY(cudaEventCreate(&h->ev_t1));
    GP_TRY(configure_kernels());
    *out = h;
    return 0;
}

extern "C" int gpso_destroy(gpso_handle* h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    cudaStreamSynchronize(h->copy_stream);
    DevBuf* bufs[] = {&h->X, &h->y, &h->Xs, &h->ls, &h->alpha, &h->K, &h->Linv, &h->LinvT, &h->T, &h->Kinv, &h->resid, &h->a,
                      &h->logdet, &h->scalars, &h->gpart, &h->gout, &h->info, &h->counter, &h->KsT, &h->part, &h->wmean,
                      &h->blockbest, &h->running, &h->cand[0], &h->cand[1], &h->leaves, &h->omean, &h->ovar};
    for (DevBuf* b : bufs) b->release();
    for (int i = 0; i < 2; i++) {
        cudaEventDestroy(h->ev_copy[i]);
        cudaEventDestroy(h->ev_used[i]);
    }
    cudaEventDestroy(h->ev_t0);
    cudaEventDestroy(h->ev_t1);
    for (cudaEvent_t e : h->prof_events) cudaEventDestroy(e);
    cudaStreamDestroy(h->stream);
    cudaStreamDestroy(h->copy_stream);
    delete h;
    return 0;
}

extern "C" int gpso_set_data(gpso_handle* h, const double* X_host, const double* y_host, int N, int d) {
    if (!h || !X_host || !y_host) return fail(GPSO_E_BADARG, "gpso_set_data: null argument");
    if (N <= 0 || d <= 0) return fail(GPSO_E_BADARG, "gpso_set_data: N and d must be positive");
    if (d > MAX_LS) return fail(GPSO_E_BADARG, "gpso_set_data: input dimension above the supported maximum (64)");
    GP_TRY(set_device(h));
    CU_TRY(cudaStreamSynchronize(h->stream));
    GP_TRY(ensure_shape(h, N, d));
    CU_TRY(cudaMemcpyAsync(h->X.p, X_host, sizeof(double) * (size_t)N * d, cudaMemcpyHostToDevice, h->stream));
    CU_TRY(cudaMemsetAsync(h->y.p, 0, sizeof(double) * h->Np, h->stream));
    CU_TRY(cudaMemcpyAsync(h->y.p, y_host, sizeof(double) * (size_t)N, cudaMemcpyHostToDevice, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    h->have_data = true;
    h->factorized = false;
    return 0;
}

static int load_theta(gpso_handle* h, const double* theta, int p, const char* who) {
    if (p != h->n_params()) {
        char b[256];
        snprintf(b, sizeof b, "%s: expected %d hyper-parameters, got %d", who, h->n_params(), p);
        return fail(GPSO_E_BADARG, b);
    }
    int nl = h->n_ls();
    for (int i = 0; i < nl; i++) {
        if (!(theta[i] > 0.0)) return fail(GPSO_E_BADARG, std::string(who) + ": lengthscale must be positive");
        h->ls_host[i] = theta[i];
    }
    h->variance = theta[nl];
    h->noise = theta[nl + 1];
    if (!(h->variance > 0.0) || !(h->noise > 0.0)) return fail(GPSO_E_BADARG, std::string(who) + ": variances must be positive");
    h->c0 = (h->mean_id == GPSO_MEAN_CONSTANT) ? theta[nl + 2] : 0.0;
    return 0;
}

extern "C" int gpso_neg_lml_grad(gpso_handle* h, const double* u, int p, double* f_host, double* grad_host) {
    if (!h || !u || !f_host || !grad_host) return fail(GPSO_E_BADARG, "gpso_neg_lml_grad: null argument");
    if (!h->have_data) return fail(GPSO_E_STATE, "gpso_neg_lml_grad: call gpso_set_data first");
    if (p != h->n_params()) return fail(GPSO_E_BADARG, "gpso_neg_lml_grad: wrong number of hyper-parameters");
    GP_TRY(set_device(h));
    const int nl = h->n_ls();
    std::vector<double> theta(p);
    for (int i = 0; i < nl; i++) theta[i] = softplus(u[i]);
    theta[nl] = softplus(u[nl]);
    theta[nl + 1] = NOISE_FLOOR + softplus(u[nl + 1]);
    if (h->mean_id == GPSO_MEAN_CONSTANT) theta[nl + 2] = u[nl + 2];
    GP_TRY(load_theta(h, theta.data(), p, "gpso_neg_lml_grad"));
    h->factorized = false;
    cudaStream_t st = h->stream;
    CU_TRY(cudaEventRecord(h->ev_t0, st));
    GP_TRY(factor_pipeline(h, st, true));
    const int nb64 = h->Np / CT;
    const int nblk = nb64 * (nb64 + 1) / 2;
    const int stride = 2 + nl;
    GP_TRY(h->gpart.ensure((size_t)nblk * stride * sizeof(double)));
    GP_TRY(h->gout.ensure((size_t)(stride + 4) * sizeof(double)));
    DISPATCH_KID(h, launch_grad, h, st, nblk, stride);
    if (cudaGetLastError() != cudaSuccess) return fail(GPSO_E_CUDA, "launch of lml_grad failed");
    reduce_partials_kernel<<<stride, 256, 0, st>>>(h->gpart.as<double>(), nblk, stride, h->gout.as<double>());
    GP_TRY(check_launch(h, "reduce_partials"));
    CU_TRY(cudaEventRecord(h->ev_t1, st));
    double sc[3];
    std::vector<double> g(stride);
    int info = 0;
    CU_TRY(cudaMemcpyAsync(sc, h->scalars.p, sizeof sc, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(g.data(), h->gout.p, sizeof(double) * stride, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(&info, h->info.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev_t0, h->ev_t1);
    h->last_ms[0] = ms;
    h->last_ms[1] = h->last_ms[2] = h->last_ms[3] = 0;
    if (info > 0) {
        char b[128];
        snprintf(b, sizeof b, "Gram matrix is not positive definite (pivot %d)", info);
        fail(info, b);
        return info;
    }
    *f_host = nlml_from_scalars(h, sc);
    // dLML/dtheta = 0.5 * sum W * dK/dtheta ; chain through softplus: dtheta/du = sigmoid(u)
    for (int i = 0; i < nl; i++) grad_host[i] = -(0.5 * g[2 + i] / h->ls_host[i]) * sigmoid(u[i]);
    grad_host[nl] = -(0.5 * g[0] / h->variance) * sigmoid(u[nl]);
    grad_host[nl + 1] = -(0.5 * g[1]) * sigmoid(u[nl + 1]);
    if (h->mean_id == GPSO_MEAN_CONSTANT) grad_host[nl + 2] = -sc[2];
    return 0;
}

extern "C" int gpso_factorize(gpso_handle* h, const double* theta_host, int p) {
    if (!h || !theta_host) return fail(GPSO_E_BADARG, "gpso_factorize: null argument");
    if (!h->have_data) return fail(GPSO_E_STATE, "gpso_factorize: call gpso_set_data first");
    GP_TRY(set_device(h));
    GP_TRY(load_theta(h, theta_host, p, "gpso_factorize"));
    h->factorized = false;
    cudaStream_t st = h->stream;
    GP_TRY(factor_pipeline(h, st, false));
    double sc[3];
    int info = 0;
    CU_TRY(cudaMemcpyAsync(sc, h->scalars.p, sizeof sc, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(&info, h->info.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    if (info > 0) {
        char b[128];
        snprintf(b, sizeof b, "Gram matrix is not positive definite (pivot %d)", info);
        fail(info, b);
        return info;
    }
    h->factor_nlml = nlml_from_scalars(h, sc);
    h->factorized = true;
    return 0;
}

extern "C" int gpso_factor_lml(gpso_handle* h, double* lml_host) {
    if (!h || !lml_host) return fail(GPSO_E_BADARG, "gpso_factor_lml: null argument");
    if (!h->factorized) return fail(GPSO_E_STATE, "gpso_factor_lml: call gpso_factorize first");
    *lml_host = -h->factor_nlml;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// predict
// ---------------------------------------------------------------------------------------------------------------------
static long long window_size(const gpso_handle* h, long long M) {
    long long group = (long long)PRED_GROUP * TB;
    long long maxw = h->window_override > 0 ? h->window_override : WINDOW_BYTES / ((long long)h->Np * 8);
    maxw = std::max(group, (maxw / group) * group);
    long long need = ((M + TB - 1) / TB) * TB;
    return std::min(maxw, need);
}

static int ensure_window(gpso_handle* h, long long W) {
    GP_TRY(h->KsT.ensure((size_t)W * h->Np * sizeof(double)));
    GP_TRY(h->part.ensure((size_t)W * h->nb * sizeof(double)));
    GP_TRY(h->wmean.ensure((size_t)W * sizeof(double)));
    GP_TRY(h->blockbest.ensure((size_t)((W + 255) / 256) * sizeof(BestRec)));
    return 0;
}

static int prof_mark(gpso_handle* h, cudaStream_t st) {
    if (!h->profile) return 0;
    if (h->prof_used == h->prof_events.size()) {
        cudaEvent_t e;
        CU_TRY(cudaEventCreate(&e));
        h->prof_events.push_back(e);
    }
    CU_TRY(cudaEventRecord(h->prof_events[h->prof_used++], st));
    return 0;
}

// sums the per-window stage times recorded by prof_mark (call after the stream has been synchronised)
static void prof_collect(gpso_handle* h) {
    h->last_ms[1] = h->last_ms[2] = h->last_ms[3] = 0.0;
    if (!h->profile) return;
    for (size_t i = 0; i + 3 < h->prof_used; i += 4) {
        float a = 0, b = 0, c = 0;
        cudaEventElapsedTime(&a, h->prof_events[i], h->prof_events[i + 1]);
        cudaEventElapsedTime(&b, h->prof_events[i + 1], h->prof_events[i + 2]);
        cudaEventElapsedTime(&c, h->prof_events[i + 2], h->prof_events[i + 3]);
        h->last_ms[1] += a;
        h->last_ms[2] += b;
        h->last_ms[3] += c;
    }
    h->prof_used = 0;
}

// one window, candidates already on the device.  mode 0: mean/var -> out_mean/out_var (device); mode 1: running best
static int run_window(gpso_handle* h, cudaStream_t st, const double* Xc_dev, long long Mw, long long idx0, int mode,
                      double varsigma, double* out_mean, double* out_var, bool first) {
    long long Mw_pad = ((Mw + TB - 1) / TB) * TB;
    h->last_windows++;
    GP_TRY(prof_mark(h, st));
    DISPATCH_KID(h, launch_crosscov, h, st, Xc_dev, Mw, Mw_pad);
    GP_TRY(check_launch(h, "crosscov"));
    PredictParams P;
    P.Linv = h->Linv.as<double>();
    P.KsT = h->KsT.as<double>();
    P.part = h->part.as<double>();
    P.Np = h->Np;
    P.nb = h->nb;
    P.nct = (int)(Mw_pad / TB);
    P.counter = h->counter.as<int>();
    CU_TRY(cudaMemsetAsync(h->counter.p, 0, sizeof(int), st));
    int grid = std::min(h->nsm, P.nct * P.nb);
    GP_TRY(prof_mark(h, st));
    predict_trmm_kernel<<<grid, GTHREADS, GEMM_SMEM_BYTES, st>>>(P);
    GP_TRY(check_launch(h, "predict_trmm"));
    GP_TRY(prof_mark(h, st));
    int fb = (int)((Mw + 255) / 256);
    predict_finalize_kernel<<<fb, 256, 0, st>>>(h->part.as<double>(), h->wmean.as<double>(), h->nb, (int)Mw_pad, Mw, idx0,
                                                h->variance, h->noise, varsigma, mode, out_mean, out_var,
                                                h->blockbest.as<BestRec>());
    GP_TRY(check_launch(h, "predict_finalize"));
    if (mode == 1) {
        best_merge_kernel<<<1, 256, 0, st>>>(h->blockbest.as<BestRec>(), fb, h->running.as<BestRec>(), first ? 1 : 0);
        GP_TRY(check_launch(h, "best_merge"));
    }
    GP_TRY(prof_mark(h, st));
    return 0;
}

static int predict_common_checks(gpso_handle* h, const void* a, long long M, const char* who) {
    if (!h || !a) return fail(GPSO_E_BADARG, std::string(who) + ": null argument");
    if (M <= 0) return fail(GPSO_E_BADARG, std::string(who) + ": M must be positive");
    if (!h->factorized) return fail(GPSO_E_STATE, std::string(who) + ": call gpso_factorize first");
    h->prof_used = 0;
    h->last_windows = 0;
    return set_device(h);
}

static int fetch_best(gpso_handle* h, cudaStream_t st, double* result_host) {
    BestRec r;
    CU_TRY(cudaMemcpyAsync(&r, h->running.p, sizeof r, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    prof_collect(h);
    result_host[0] = (double)r.idx;
    result_host[1] = r.mean;
    result_host[2] = r.var;
    result_host[3] = r.ucb;
    return 0;
}

// device-resident candidates, all windows on `st`
static int run_dev(gpso_handle* h, cudaStream_t st, const double* Xc_dev, long long M, int mode, double varsigma,
                   double* mean_dev, double* var_dev) {
    long long W = window_size(h, M);
    GP_TRY(ensure_window(h, W));
    for (long long off = 0; off < M; off += W) {
        long long Mw = std::min(W, M - off);
        GP_TRY(run_window(h, st, Xc_dev + off * h->d, Mw, off, mode, varsigma, mean_dev ? mean_dev + off : nullptr,
                          var_dev ? var_dev + off : nullptr, off == 0));
    }
    return 0;
}

extern "C" int gpso_predict_y_dev(gpso_handle* h, const double* Xc_dev, int64_t M, double* mean_dev, double* var_dev,
                                  void* stream) {
    GP_TRY(predict_common_checks(h, Xc_dev, M, "gpso_predict_y_dev"));
    if (!mean_dev || !var_dev) return fail(GPSO_E_BADARG, "gpso_predict_y_dev: null output");
    return run_dev(h, (cudaStream_t)stream, Xc_dev, M, 0, 0.0, mean_dev, var_dev);
}

extern "C" int gpso_ucb_argmax_dev(gpso_handle* h, const double* Xc_dev, int64_t M, double varsigma, double* result_host,
                                   void* stream) {
    GP_TRY(predict_common_checks(h, Xc_dev, M, "gpso_ucb_argmax_dev"));
    if (!result_host) return fail(GPSO_E_BADARG, "gpso_ucb_argmax_dev: null output");
    cudaStream_t st = (cudaStream_t)stream;
    CU_TRY(cudaEventRecord(h->ev_t0, st));
    GP_TRY(run_dev(h, st, Xc_dev, M, 1, varsigma, nullptr, nullptr));
    CU_TRY(cudaEventRecord(h->ev_t1, st));
    GP_TRY(fetch_best(h, st, result_host));
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev_t0, h->ev_t1);
    h->last_ms[0] = ms;
    return 0;
}

// host-resident candidates: windows are staged through two device buffers; the H2D copy of window i+1 (copy stream)
// overlaps the kernels of window i (compute stream)
static int run_host(gpso_handle* h, const double* Xc_host, long long M, int mode, double varsigma, double* mean_host,
                    double* var_host) {
    long long W = window_size(h, M);
    GP_TRY(ensure_window(h, W));
    const int d = h->d;
    size_t cbytes = (size_t)W * d * sizeof(double);
    GP_TRY(h->cand[0].ensure(cbytes));
    GP_TRY(h->cand[1].ensure(cbytes));
    if (mode == 0) {
        GP_TRY(h->omean.ensure((size_t)W * sizeof(double)));
        GP_TRY(h->ovar.ensure((size_t)W * sizeof(double)));
    }
    cudaStream_t st = h->stream, cs = h->copy_stream;
    CU_TRY(cudaEventRecord(h->ev_t0, st));
    h->used_pending[0] = h->used_pending[1] = false;
    long long nwin = (M + W - 1) / W;
    // prefetch window 0
    {
        long long Mw = std::min(W, M);
        CU_TRY(cudaMemcpyAsync(h->cand[0].p, Xc_host, (size_t)Mw * d * sizeof(double), cudaMemcpyHostToDevice, cs));
        CU_TRY(cudaEventRecord(h->ev_copy[0], cs));
    }
    for (long long w = 0; w < nwin; w++) {
        int b = (int)(w & 1);
        long long off = w * W;
        long long Mw = std::min(W, M - off);
        CU_TRY(cudaStreamWaitEvent(st, h->ev_copy[b], 0));
        GP_TRY(run_window(h, st, h->cand[b].as<double>(), Mw, off, mode, varsigma, h->omean.as<double>(), h->ovar.as<double>(),
                          w == 0));
        CU_TRY(cudaEventRecord(h->ev_used[b], st));
        h->used_pending[b] = true;
        if (w + 1 < nwin) {  // stage the next window while this one computes
            int nb2 = b ^ 1;
            long long off2 = (w + 1) * W;
            long long Mw2 = std::min(W, M - off2);
            if (h->used_pending[nb2]) CU_TRY(cudaStreamWaitEvent(cs, h->ev_used[nb2], 0));
            CU_TRY(cudaMemcpyAsync(h->cand[nb2].p, Xc_host + off2 * d, (size_t)Mw2 * d * sizeof(double), cudaMemcpyHostToDevice, cs));
            CU_TRY(cudaEventRecord(h->ev_copy[nb2], cs));
        }
        if (mode == 0) {
            CU_TRY(cudaMemcpyAsync(mean_host + off, h->omean.p, (size_t)Mw * sizeof(double), cudaMemcpyDeviceToHost, st));
            CU_TRY(cudaMemcpyAsync(var_host + off, h->ovar.p, (size_t)Mw * sizeof(double), cudaMemcpyDeviceToHost, st));
        }
    }
    CU_TRY(cudaEventRecord(h->ev_t1, st));
    return 0;
}

extern "C" int gpso_predict_y_host(gpso_handle* h, const double* Xc_host, int64_t M, double* mean_host, double* var_host) {
    GP_TRY(predict_common_checks(h, Xc_host, M, "gpso_predict_y_host"));
    if (!mean_host || !var_host) return fail(GPSO_E_BADARG, "gpso_predict_y_host: null output");
    GP_TRY(run_host(h, Xc_host, M, 0, 0.0, mean_host, var_host));
    CU_TRY(cudaStreamSynchronize(h->stream));
    prof_collect(h);
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev_t0, h->ev_t1);
    h->last_ms[0] = ms;
    return 0;
}

extern "C" int gpso_ucb_argmax_host(gpso_handle* h, const double* Xc_host, int64_t M, double varsigma, double* result_host) {
    GP_TRY(predict_common_checks(h, Xc_host, M, "gpso_ucb_argmax_host"));
    if (!result_host) return fail(GPSO_E_BADARG, "gpso_ucb_argmax_host: null output");
    GP_TRY(run_host(h, Xc_host, M, 1, varsigma, nullptr, nullptr));
    GP_TRY(fetch_best(h, h->stream, result_host));
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev_t0, h->ev_t1);
    h->last_ms[0] = ms;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// leaves
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int64_t gpso_grow_count(int depth) {
    if (depth < 0 || depth > 38) return -1;
    int64_t n = 0, w = 1;
    for (int l = 0; l < depth; l++) {
        n += w;
        w *= 3;
    }
    return n;
}

static int grow_launch(const double* bounds_host, int d, int depth, double* out_dev, cudaStream_t st, DevBuf& bdev) {
    if (!bounds_host || !out_dev) return fail(GPSO_E_BADARG, "gpso_grow_leaves: null argument");
    if (d <= 0 || d > LEAF_MAXD) return fail(GPSO_E_BADARG, "gpso_grow_leaves: dimension must be in 1..64");
    if (depth < 1 || depth > 20) return fail(GPSO_E_BADARG, "gpso_grow_leaves: depth must be in 1..20");
    for (int j = 0; j < d; j++)
        if (!(bounds_host[2 * j + 1] > bounds_host[2 * j])) return fail(GPSO_E_BADARG, "gpso_grow_leaves: need hi > lo per dimension");
    GP_TRY(bdev.ensure(sizeof(double) * 2 * LEAF_MAXD));
    CU_TRY(cudaMemcpyAsync(bdev.p, bounds_host, sizeof(double) * 2 * d, cudaMemcpyHostToDevice, st));
    long long rows = gpso_grow_count(depth);
    grow_leaves_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(bdev.as<double>(), d, depth, rows, out_dev);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(GPSO_E_CUDA, std::string("launch of grow_leaves failed: ") + cudaGetErrorString(e));
    return 0;
}

extern "C" int gpso_grow_leaves_dev(int device, const double* bounds_host, int d, int depth, double* out_dev, void* stream) {
    CU_TRY(cudaSetDevice(device));
    DevBuf b;
    int rc = grow_launch(bounds_host, d, depth, out_dev, (cudaStream_t)stream, b);
    if (rc == 0) {
        cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);  // bounds buffer is freed below
        if (e != cudaSuccess) rc = fail(GPSO_E_CUDA, cudaGetErrorString(e));
    }
    b.release();
    return rc;
}

extern "C" int gpso_grow_leaves_host(int device, const double* bounds_host, int d, int depth, double* out_host) {
    if (!out_host) return fail(GPSO_E_BADARG, "gpso_grow_leaves_host: null output");
    CU_TRY(cudaSetDevice(device));
    long long rows = gpso_grow_count(depth);
    if (rows <= 0) return fail(GPSO_E_BADARG, "gpso_grow_leaves_host: bad depth");
    DevBuf out, b;
    GP_TRY(out.ensure((size_t)rows * d * sizeof(double)));
    int rc = grow_launch(bounds_host, d, depth, out.as<double>(), 0, b);
    if (rc == 0) {
        cudaError_t e = cudaMemcpy(out_host, out.p, (size_t)rows * d * sizeof(double), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail(GPSO_E_CUDA, cudaGetErrorString(e));
    }
    out.release();
    b.release();
    return rc;
}

extern "C" int gpso_grow_ucb_argmax(gpso_handle* h, const double* bounds_host, int d, int depth, double varsigma,
                                    double* result_host) {
    if (!h || !bounds_host || !result_host) return fail(GPSO_E_BADARG, "gpso_grow_ucb_argmax: null argument");
    if (!h->factorized) return fail(GPSO_E_STATE, "gpso_grow_ucb_argmax: call gpso_factorize first");
    h->prof_used = 0;
    h->last_windows = 0;
    if (d != h->d) return fail(GPSO_E_BADARG, "gpso_grow_ucb_argmax: dimension differs from the training data");
    GP_TRY(set_device(h));
    long long rows = gpso_grow_count(depth);
    if (rows <= 0) return fail(GPSO_E_BADARG, "gpso_grow_ucb_argmax: bad depth");
    GP_TRY(h->leaves.ensure((size_t)rows * d * sizeof(double) + sizeof(double) * 2 * LEAF_MAXD));
    // bounds live at the tail of the leaves buffer
    double* bdev = h->leaves.as<double>() + (size_t)rows * d;
    cudaStream_t st = h->stream;
    for (int j = 0; j < d; j++)
        if (!(bounds_host[2 * j + 1] > bounds_host[2 * j])) return fail(GPSO_E_BADARG, "gpso_grow_ucb_argmax: need hi > lo per dimension");
    if (depth < 1 || depth > 20) return fail(GPSO_E_BADARG, "gpso_grow_ucb_argmax: depth must be in 1..20");
    CU_TRY(cudaEventRecord(h->ev_t0, st));
    CU_TRY(cudaMemcpyAsync(bdev, bounds_host, sizeof(double) * 2 * d, cudaMemcpyHostToDevice, st));
    grow_leaves_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(bdev, d, depth, rows, h->leaves.as<double>());
    GP_TRY(check_launch(h, "grow_leaves"));
    GP_TRY(run_dev(h, st, h->leaves.as<double>(), rows, 1, varsigma, nullptr, nullptr));
    CU_TRY(cudaEventRecord(h->ev_t1, st));
    GP_TRY(fetch_best(h, st, result_host));
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev_t0, h->ev_t1);
    h->last_ms[0] = ms;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// multi-GPU state exchange: [ header (128 doubles) | Xs (d*Np) | alpha (Np) | Linv (Np*Np) ]
// ---------------------------------------------------------------------------------------------------------------------
constexpr int STATE_HEADER = 128;

extern "C" int gpso_state_bytes(gpso_handle* h, int N, int d, int64_t* bytes) {
    if (!h || !bytes || N <= 0 || d <= 0) return fail(GPSO_E_BADARG, "gpso_state_bytes: bad argument");
    long long Np = ((N + TB - 1) / TB) * TB;
    *bytes = (int64_t)sizeof(double) * (STATE_HEADER + (long long)d * Np + Np + Np * Np);
    return 0;
}

extern "C" int gpso_export_state_dev(gpso_handle* h, void* dst_dev, int64_t bytes, void* stream) {
    if (!h || !dst_dev) return fail(GPSO_E_BADARG, "gpso_export_state_dev: null argument");
    if (!h->factorized) return fail(GPSO_E_STATE, "gpso_export_state_dev: call gpso_factorize first");
    GP_TRY(set_device(h));
    int64_t need = 0;
    gpso_state_bytes(h, h->N, h->d, &need);
    if (bytes < need) return fail(GPSO_E_BADARG, "gpso_export_state_dev: destination too small");
    cudaStream_t st = (cudaStream_t)stream;
    double hdr[STATE_HEADER] = {0};
    hdr[0] = h->N;
    hdr[1] = h->d;
    hdr[2] = h->variance;
    hdr[3] = h->noise;
    hdr[4] = h->c0;
    hdr[5] = h->factor_nlml;
    hdr[6] = h->kernel_id;
    hdr[7] = h->ard;
    for (int i = 0; i < h->n_ls(); i++) hdr[16 + i] = h->ls_host[i];
    double* dst = (double*)dst_dev;
    CU_TRY(cudaStreamSynchronize(h->stream));
    CU_TRY(cudaMemcpyAsync(dst, hdr, sizeof hdr, cudaMemcpyHostToDevice, st));
    size_t Np = h->Np;
    CU_TRY(cudaMemcpyAsync(dst + STATE_HEADER, h->Xs.p, sizeof(double) * h->d * Np, cudaMemcpyDeviceToDevice, st));
    CU_TRY(cudaMemcpyAsync(dst + STATE_HEADER + h->d * Np, h->alpha.p, sizeof(double) * Np, cudaMemcpyDeviceToDevice, st));
    CU_TRY(cudaMemcpyAsync(dst + STATE_HEADER + h->d * Np + Np, h->Linv.p, sizeof(double) * Np * Np, cudaMemcpyDeviceToDevice, st));
    CU_TRY(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int gpso_import_state_dev(gpso_handle* h, const void* src_dev, int64_t bytes, int N, int d, void* stream) {
    if (!h || !src_dev) return fail(GPSO_E_BADARG, "gpso_import_state_dev: null argument");
    if (N <= 0 || d <= 0 || d > MAX_LS) return fail(GPSO_E_BADARG, "gpso_import_state_dev: bad shape");
    GP_TRY(set_device(h));
    int64_t need = 0;
    gpso_state_bytes(h, N, d, &need);
    if (bytes < need) return fail(GPSO_E_BADARG, "gpso_import_state_dev: source too small");
    cudaStream_t st = (cudaStream_t)stream;
    CU_TRY(cudaStreamSynchronize(h->stream));
    GP_TRY(ensure_shape(h, N, d));
    const double* src = (const double*)src_dev;
    double hdr[STATE_HEADER];
    CU_TRY(cudaMemcpyAsync(hdr, src, sizeof hdr, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    if ((int)hdr[0] != N || (int)hdr[1] != d) return fail(GPSO_E_BADARG, "gpso_import_state_dev: header shape mismatch");
    if ((int)hdr[6] != h->kernel_id || (int)hdr[7] != h->ard) return fail(GPSO_E_BADARG, "gpso_import_state_dev: kernel mismatch");
    h->variance = hdr[2];
    h->noise = hdr[3];
    h->c0 = hdr[4];
    h->factor_nlml = hdr[5];
    for (int i = 0; i < h->n_ls(); i++) h->ls_host[i] = hdr[16 + i];
    size_t Np = h->Np;
    CU_TRY(cudaMemcpyAsync(h->Xs.p, src + STATE_HEADER, sizeof(double) * d * Np, cudaMemcpyDeviceToDevice, st));
    CU_TRY(cudaMemcpyAsync(h->alpha.p, src + STATE_HEADER + d * Np, sizeof(double) * Np, cudaMemcpyDeviceToDevice, st));
    CU_TRY(cudaMemcpyAsync(h->Linv.p, src + STATE_HEADER + d * Np + Np, sizeof(double) * Np * Np, cudaMemcpyDeviceToDevice, st));
    GP_TRY(upload_lengthscales(h, st));
    CU_TRY(cudaStreamSynchronize(st));
    h->have_data = false;  // no raw training data on this rank: predict only
    h->factorized = true;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// introspection
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int64_t gpso_launch_count(gpso_handle* h) { return h ? h->launches : 0; }

extern "C" int gpso_debug_fetch(gpso_handle* h, int which, double* out_host, int64_t count) {
    if (!h || !out_host) return fail(GPSO_E_BADARG, "gpso_debug_fetch: null argument");
    GP_TRY(set_device(h));
    CU_TRY(cudaStreamSynchronize(h->stream));
    const int N = h->N, Np = h->Np;
    if (which == 3) {
        if (count < N) return fail(GPSO_E_BADARG, "gpso_debug_fetch: buffer too small");
        CU_TRY(cudaMemcpy(out_host, h->alpha.p, sizeof(double) * N, cudaMemcpyDeviceToHost));
        return 0;
    }
    const DevBuf* src = which == 1 ? &h->K : which == 2 ? &h->Linv : which == 4 ? &h->Kinv : nullptr;
    if (!src || !src->p) return fail(GPSO_E_BADARG, "gpso_debug_fetch: unknown or unavailable matrix");
    if (count < (int64_t)N * N) return fail(GPSO_E_BADARG, "gpso_debug_fetch: buffer too small");
    CU_TRY(cudaMemcpy2D(out_host, sizeof(double) * N, src->p, sizeof(double) * Np, sizeof(double) * N, N, cudaMemcpyDeviceToHost));
    for (int i = 0; i < N; i++)
        for (int j = i + 1; j < N; j++) out_host[(size_t)i * N + j] = 0.0;  // only the lower triangle is defined
    return 0;
}

extern "C" int gpso_last_timing(gpso_handle* h, double* out_ms4) {
    if (!h || !out_ms4) return fail(GPSO_E_BADARG, "gpso_last_timing: null argument");
    for (int i = 0; i < 4; i++) out_ms4[i] = h->last_ms[i];
    return 0;
}

extern "C" int gpso_set_profile(gpso_handle* h, int enabled) {
    if (!h) return fail(GPSO_E_BADARG, "gpso_set_profile: null handle");
    h->profile = enabled != 0;
    return 0;
}

extern "C" int64_t gpso_last_windows(gpso_handle* h) { return h ? h->last_windows : 0; }

extern "C" int gpso_set_window(gpso_handle* h, int64_t candidates) {
    if (!h || candidates < 0) return fail(GPSO_E_BADARG, "gpso_set_window: bad argument");
    h->window_override = candidates;
    return 0;
}

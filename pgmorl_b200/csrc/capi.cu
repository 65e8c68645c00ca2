// Error plumbing and trivial entry points of the C ABI (include/pgmorl_b200.h).
#include <stdarg.h>

#include "common.cuh"

namespace pgm {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return PGM_ERR_CUDA;
}

}  // namespace pgm

extern "C" int pgm_abi_version(void) { return PGM_ABI_VERSION; }
extern "C" const char *pgm_last_error(void) { return pgm::g_err; }
extern "C" int pgm_n_par(int O, int A, int M) { return pgm::NetLayout(O, A, M).n_par; }

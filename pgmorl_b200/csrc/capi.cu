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

// Offsets of the 13 parameter tensors inside one flat vector, in named_parameters() order (pgmorl_b200/layout.py):
// the kernels and the Python state_dict slicing must agree on them (tests/test_cabi.py).
extern "C" int pgm_param_offsets(int O, int A, int M, int *out13) {
    if (!out13) return PGM_ERR_ARG;
    const pgm::NetLayout L(O, A, M);
    const int v[13] = {L.oW1a, L.ob1a, L.oW2a, L.ob2a, L.oW1c, L.ob1c, L.oW2c, L.ob2c, L.oWv, L.obv, L.oWmu, L.obmu, L.ols};
    for (int i = 0; i < 13; ++i) out13[i] = v[i];
    return PGM_OK;
}

// fastvim_b200 -- error reporting, launch accounting, argument checks (host side of the C ABI).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace fv {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}
int finish_launch(const char* what) {
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
        return 2;
    }
    return 0;
}
// SM count of the current device (cached per device; persistent kernels size their grids with it).
int sm_count() {
    static int cache[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cache[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cache[dev] = n;
    }
    return cache[dev];
}

static int g_pdl = -1;
bool pdl_enabled() {
    if (g_pdl < 0) {
        const char* e = getenv("FASTVIM_PDL");
        g_pdl = (e && e[0] == '0') ? 0 : 1;
    }
    return g_pdl != 0;
}

// The attribute is set by default only on the FastVim-T inference chain (fv_gemm_bf16_tn, fv_block_fwd, fv_gemm_out_norm), where it
// was measured to pay (+2.8 %, and the block -> out_proj dataflow needs it).  Every other kernel of the library is PDL-correct
// (griddepcontrol.wait before its first global access) but launches in plain stream order unless FASTVIM_PDL_ALL=1: with the
// multi-wave, non-persistent kernels of the wider models the early-resident dependents cost 1-2 % (FastVim-S/B, FastChannelVim).
static int g_pdl_all = -1;
bool pdl_all_enabled() {
    if (g_pdl_all < 0) {
        const char* e = getenv("FASTVIM_PDL_ALL");
        g_pdl_all = (e && e[0] == '1') ? 1 : 0;
    }
    return g_pdl_all != 0 && pdl_enabled();
}

int check_geom(const fv_geom* g, const char* who) {
    FV_REQUIRE(g != nullptr, "%s: null geometry", who);
    FV_REQUIRE(g->batch > 0 && g->dim > 0 && g->outer > 0 && g->pool > 0 && g->inner > 0,
               "%s: non-positive geometry (batch %d dim %d outer %d pool %d inner %d)", who, g->batch,
               g->dim, g->outer, g->pool, g->inner);
    FV_REQUIRE(g->dim % 4 == 0, "%s: dim (%d) must be a multiple of 4", who, g->dim);
    FV_REQUIRE((int64_t)g->outer * g->pool * g->inner < (1ll << 30), "%s: sequence too long", who);
    return 0;
}

}  // namespace fv

extern "C" const char* fv_last_error(void) { return fv::g_err; }
extern "C" int fv_set_pdl_all(int on) {
    const int prev = fv::g_pdl_all > 0 ? 1 : 0;
    fv::g_pdl_all = on ? 1 : 0;
    return prev;
}
extern "C" int fv_set_pdl(int on) {
    const int prev = fv::pdl_enabled() ? 1 : 0;
    fv::g_pdl = on ? 1 : 0;
    return prev;
}
extern "C" int fv_version(void) { return 100; }
extern "C" int64_t fv_launch_count(void) { return fv::g_launches; }
extern "C" void fv_reset_launch_count(void) { fv::g_launches = 0; }

// glg_abi.cu - error plumbing of the C ABI (include/glg_b200.h).
#include <stdarg.h>
#include <string.h>

#include "glg_common.cuh"

namespace glg {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int launch_status(const char* what)
{
    const cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return GLG_OK;
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return GLG_ERR_LAUNCH;
}

}  // namespace glg

extern "C" const char* glg_last_error(void) { return glg::g_error; }
extern "C" int glg_abi_version(void) { return 7; }

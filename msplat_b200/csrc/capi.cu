// msplat_b200/csrc/capi.cu -- C-ABI plumbing shared by all entry points (include/msplat_b200.h).
// Error convention: every entry point returns 0 on success, a positive cudaError_t if a launch
// failed, or a negative MSB_ERR_* validation code; msb_last_error() returns a thread-local
// description.  The library keeps no other mutable global state, never allocates or frees
// device memory and never synchronises the device.
#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace msb {

static thread_local char g_err[512] = "";

int set_error(int code, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return MSB_OK;
}

}  // namespace msb

extern "C" {

const char* msb_last_error(void) { return msb::g_err; }

int msb_version(void) { return 100; }

// Number of SMs of the current device (used by the host side to size persistent grids).
int msb_sm_count(void) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    return n;
}

}  // extern "C"

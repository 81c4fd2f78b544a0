// api.cu — handle lifetime and error reporting for the abcdls C ABI (include/abcdls.h).
#include "common.cuh"
#include <stdlib.h>

extern "C" {

int abcdls_version(void) { return ABCDLS_VERSION_MAJOR * 100 + ABCDLS_VERSION_MINOR; }

int abcdls_create(abcdls_handle_t* out, int device) {
    if (!out) return ABCDLS_E_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        cudaGetLastError();
        return ABCDLS_E_NODEVICE;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return ABCDLS_E_CUDA;
    if (prop.major != 10) return ABCDLS_E_NODEVICE;  // kernels are built for sm_100a only
    abcdls_handle_s* h = new abcdls_handle_s();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    const char* e = getenv("ABCDLS_NO_TMA");
    h->no_tma = (e && e[0] == '1') ? 1 : 0;
    *out = h;
    return ABCDLS_OK;
}

int abcdls_destroy(abcdls_handle_t h) {
    delete h;
    return ABCDLS_OK;
}

const char* abcdls_last_error(abcdls_handle_t h) { return h ? h->err.c_str() : "null handle"; }

int abcdls_sm_count(abcdls_handle_t h) { return h ? h->sm_count : 0; }

}  // extern "C"

// afsk_api.cu — library-level entry points of include/afsk_b200.h: errors, device info and the
// minimal memory/stream helpers that let a host binding (ctypes) run without another CUDA wrapper.
#include <stdarg.h>
#include <string.h>

#include "afsk_common.cuh"

static thread_local char g_err[512] = "";

void afsk_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" {

int afsk_abi_version(void) { return AFSK_ABI_VERSION; }

const char *afsk_last_error(void) { return g_err; }

int afsk_device_count(int *count)
{
    if (!count) return AFSK_E_ARG;
    *count = 0;
    AFSK_CUDA(cudaGetDeviceCount(count));
    return AFSK_OK;
}

int afsk_device_info(int device, int *sm_count, size_t *mem_bytes, int *cc)
{
    cudaDeviceProp p;
    AFSK_CUDA(cudaGetDeviceProperties(&p, device));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (mem_bytes) *mem_bytes = p.totalGlobalMem;
    if (cc) *cc = p.major * 10 + p.minor;
    return AFSK_OK;
}

int afsk_malloc(int device, size_t bytes, void **dptr)
{
    if (!dptr) return AFSK_E_ARG;
    AfskDeviceGuard g(device);
    if (!g.ok) { afsk_set_error("cannot select device %d", device); return AFSK_E_CUDA; }
    AFSK_CUDA(cudaMalloc(dptr, bytes ? bytes : 16));
    return AFSK_OK;
}

int afsk_free(int device, void *dptr)
{
    AfskDeviceGuard g(device);
    if (!g.ok) return AFSK_E_CUDA;
    AFSK_CUDA(cudaFree(dptr));
    return AFSK_OK;
}

int afsk_host_alloc(size_t bytes, void **hptr)
{
    if (!hptr) return AFSK_E_ARG;
    AFSK_CUDA(cudaMallocHost(hptr, bytes ? bytes : 16));
    return AFSK_OK;
}

int afsk_host_free(void *hptr)
{
    AFSK_CUDA(cudaFreeHost(hptr));
    return AFSK_OK;
}

int afsk_memcpy_h2d(int device, void *dst, const void *src, size_t bytes, void *stream)
{
    AfskDeviceGuard g(device);
    if (!g.ok) return AFSK_E_CUDA;
    AFSK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return AFSK_OK;
}

int afsk_memcpy_d2h(int device, void *dst, const void *src, size_t bytes, void *stream)
{
    AfskDeviceGuard g(device);
    if (!g.ok) return AFSK_E_CUDA;
    AFSK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return AFSK_OK;
}

int afsk_memset(int device, void *dst, int value, size_t bytes, void *stream)
{
    AfskDeviceGuard g(device);
    if (!g.ok) return AFSK_E_CUDA;
    AFSK_CUDA(cudaMemsetAsync(dst, value, bytes, (cudaStream_t)stream));
    return AFSK_OK;
}

int afsk_stream_create(int device, void **stream)
{
    if (!stream) return AFSK_E_ARG;
    AfskDeviceGuard g(device);
    if (!g.ok) return AFSK_E_CUDA;
    cudaStream_t s;
    AFSK_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = (void *)s;
    return AFSK_OK;
}

int afsk_stream_destroy(int device, void *stream)
{
    AfskDeviceGuard g(device);
    if (!g.ok) return AFSK_E_CUDA;
    AFSK_CUDA(cudaStreamDestroy((cudaStream_t)stream));
    return AFSK_OK;
}

int afsk_stream_wait_stream(int device, void *waiter, void *signaller)
{
    AfskDeviceGuard g(device);
    if (!g.ok) return AFSK_E_CUDA;
    cudaEvent_t ev;
    AFSK_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    cudaError_t e = cudaEventRecord(ev, (cudaStream_t)signaller);
    if (e == cudaSuccess) e = cudaStreamWaitEvent((cudaStream_t)waiter, ev, 0);
    cudaEventDestroy(ev);                    // released once the recorded work has completed
    if (e != cudaSuccess) { afsk_set_error("afsk_stream_wait_stream: %s", cudaGetErrorString(e)); return AFSK_E_CUDA; }
    return AFSK_OK;
}

int afsk_stream_sync(int device, void *stream)
{
    AfskDeviceGuard g(device);
    if (!g.ok) return AFSK_E_CUDA;
    AFSK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return AFSK_OK;
}

int afsk_tone_lengths(int baud, int *bit_frames, int *mark_len, int *space_len)
{
    int bf, ml, sl;
    if (!afsk_tone_geometry(baud, &bf, &ml, &sl)) {
        afsk_set_error("Invalid baud rate.");
        return AFSK_E_BAUD;
    }
    if (bit_frames) *bit_frames = bf;
    if (mark_len) *mark_len = ml;
    if (space_len) *space_len = sl;
    return AFSK_OK;
}

}  // extern "C"

// tools/hbm_read_peak.cu — read-only HBM ceilings on this B200, for the k_demod roofline discussion.
//
// MEASURED_PEAKS.json's hbm_gbs is a COPY (read + write).  k_demod only reads (2 bits written per
// 40+ samples), so the relevant ceiling is the read-only one.  This tool measures it two ways over a
// buffer larger than L2, with CUDA events:
//   ldg   : grid-stride 128-bit ld.global.nc loads, xor-reduced (no shared memory)
//   bulk  : the k_demod fetch structure with the compute removed — persistent CTAs (2 per SM), one
//           producer lane issuing 1-D TMA bulk copies (cp.async.bulk) into a 4-stage shared-memory
//           ring, consumer warps only wait on the full barrier, read one word and release the stage
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/hbm_read_peak tools/hbm_read_peak.cu
//   gpurun_out/hbm_read_peak [GiB] [tile_bytes] [stages] [ctas_per_sm] [misalign_bytes] [producer_sleep_ns] [consumer_spin_iters]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(512) k_ldg(const uint4 *__restrict__ src, size_t nvec, uint32_t *sink)
{
    uint32_t acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < nvec; i += 4 * stride) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++)
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(src + i + u * stride));
#pragma unroll
        for (int u = 0; u < 4; u++) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    for (; i < nvec; i += stride) { const uint4 v = src[i]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 0x12345678u) *sink = acc;
}

__global__ void __launch_bounds__(288, 2) k_bulk(const uint8_t *__restrict__ src, long long ntiles, int tile_bytes, int S,
                                                 uint32_t *sink, int misalign, int sleep_ns, int spin_iters)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)S * (tile_bytes + 128));
    uint64_t *empty = full + 8;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < S; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[s])), "r"(8));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long per = (ntiles + gridDim.x - 1) / gridDim.x;
    const long long lo = blockIdx.x * per, hi = min(ntiles, lo + per);
    if (lo >= hi) return;
    auto try_wait = [](uint64_t *bar, uint32_t parity) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        return ok != 0;
    };
    if (warp == 8) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (long long it = lo; it < hi; ++it) {
                if (it - lo >= S) while (!try_wait(&empty[s], ph ^ 1u)) { if (sleep_ns) __nanosleep(sleep_ns); }
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(tile_bytes + (misalign ? 16 : 0)) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(smem + (size_t)s * (tile_bytes + 128))), "l"(src + it * tile_bytes + misalign), "r"(tile_bytes + (misalign ? 16 : 0)),
                               "r"(smem_u32(&full[s])) : "memory");
                if (++s == S) { s = 0; ph ^= 1u; }
            }
        }
        return;
    }
    int s = 0; uint32_t ph = 0, acc = 0;
    for (long long it = lo; it < hi; ++it) {
        while (!try_wait(&full[s], ph)) { }
        acc ^= reinterpret_cast<const uint32_t *>(smem + (size_t)s * (tile_bytes + 128))[tid];
        for (int k = 0; k < spin_iters; k++) acc = acc * 1664525u + 1013904223u;     // stand-in for consumer work
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
        if (++s == S) { s = 0; ph ^= 1u; }
    }
    if (acc == 0x12345678u) *sink = acc;
}

int main(int argc, char **argv)
{
    const double gib = argc > 1 ? atof(argv[1]) : 4.6;
    const int tile = argc > 2 ? atoi(argv[2]) : 20480;
    const int S = argc > 3 ? atoi(argv[3]) : 4;
    const int per_sm = argc > 4 ? atoi(argv[4]) : 2;
    const int misalign = argc > 5 ? atoi(argv[5]) : 0;
    const int sleep_ns = argc > 6 ? atoi(argv[6]) : 64;
    const int spin = argc > 7 ? atoi(argv[7]) : 0;
    const size_t bytes = ((size_t)(gib * (1ull << 30)) / tile) * tile;
    uint8_t *buf; uint32_t *sink;
    CK(cudaMalloc(&buf, bytes)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(buf, 1, bytes));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = 20;
    float ms;
    // ldg
    for (int w = 0; w < 3; w++) k_ldg<<<p.multiProcessorCount * 4, 512>>>((const uint4 *)buf, bytes / 16, sink);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; i++) k_ldg<<<p.multiProcessorCount * 4, 512>>>((const uint4 *)buf, bytes / 16, sink);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("{\"kernel\": \"ldg128 read-only\", \"bytes\": %zu, \"ms\": %.4f, \"GBps\": %.1f}\n", bytes, ms / iters, bytes / (ms / iters) / 1e6);
    // bulk ring
    const size_t smem = (size_t)S * (tile + 128) + 128;
    CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long ntiles = bytes / tile;
    for (int w = 0; w < 3; w++) k_bulk<<<p.multiProcessorCount * per_sm, 288, smem>>>(buf, ntiles - 1, tile, S, sink, misalign, sleep_ns, spin);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; i++) k_bulk<<<p.multiProcessorCount * per_sm, 288, smem>>>(buf, ntiles - 1, tile, S, sink, misalign, sleep_ns, spin);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("{\"kernel\": \"tma bulk ring read-only\", \"tile_bytes\": %d, \"stages\": %d, \"ctas_per_sm\": %d, \"misalign\": %d, \"sleep_ns\": %d, \"spin\": %d, \"bytes\": %zu, \"ms\": %.4f, \"GBps\": %.1f}\n",
           tile, S, per_sm, misalign, sleep_ns, spin, bytes, ms / iters, bytes / (ms / iters) / 1e6);
    CK(cudaGetLastError());
    return 0;
}

// Legality / behaviour test: a 2-D tensor map whose rows OVERLAP in memory (dim0 = 511 int16, row stride 512 B)
// so that a (256 x R) box at (x0 = s % 256, y0 = s / 256) is R*256 CONTIGUOUS samples starting at the
// arbitrary sample index s, landed 128-byte aligned in shared memory by the TMA unit.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int kRows = 64, kW = 256;

__global__ void k_load(const __grid_constant__ CUtensorMap tm, int x0, int y0, int16_t *out)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar), dst = (uint32_t)__cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(kRows * kW * 2) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(dst), "l"(&tm), "r"(x0), "r"(y0), "r"(bar_a) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar_a) : "memory");
    }
    const int16_t *s = reinterpret_cast<const int16_t *>(smem);
    for (int i = threadIdx.x; i < kRows * kW; i += blockDim.x) out[i] = s[i];
}

int main()
{
    const long long N = 1 << 22;
    std::vector<int16_t> h(N);
    for (long long i = 0; i < N; i++) h[i] = (int16_t)((i * 2654435761u) >> 13);
    int16_t *d = nullptr, *d_out = nullptr;
    cudaMalloc(&d, N * 2 + 1024);
    cudaMalloc(&d_out, kRows * kW * 2);
    cudaMemcpy(d, h.data(), N * 2, cudaMemcpyHostToDevice);
    EncodeFn enc = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&enc, cudaEnableDefault, &qr) != cudaSuccess || !enc) {
        printf("no cuTensorMapEncodeTiled\n");
        return 2;
    }
    cudaFuncSetAttribute(k_load, cudaFuncAttributeMaxDynamicSharedMemorySize, kRows * kW * 2);
    int bad_total = 0;
    std::vector<int16_t> out(kRows * kW);
    // variant 0: rows overlap (dim0 = 511, stride 512 B); 1: no overlap (dim0 = 256): unaligned start -> zero fill past the row;
    // 2: 1-D tensor, box 256
    for (int variant = 0; variant < 3; variant++) {
        CUtensorMap tm;
        const cuuint64_t gdim[2] = {variant == 0 ? 511u : (variant == 1 ? 256u : (cuuint64_t)N), (cuuint64_t)(N / 256 - 1)};
        const cuuint64_t gstride[1] = {512};
        const cuuint32_t box[2] = {kW, (cuuint32_t)(variant == 2 ? 1 : kRows)}, estr[2] = {1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, variant == 2 ? 1 : 2, d, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("variant %d encode -> %d\n", variant, (int)r);
        if (r != CUDA_SUCCESS) continue;
        if (variant == 2) continue;   // kernel is 2-D only
        const long long starts[] = {0, 8, 256, 264, 1, 7, 255, 257, 1000003, 4000001};
        for (long long s : starts) {
            k_load<<<1, 256, kRows * kW * 2>>>(tm, (int)(s % 256), (int)(s / 256), d_out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("variant %d start %lld: launch: %s\n", variant, s, cudaGetErrorString(e)); return 3; }
            cudaMemcpy(out.data(), d_out, kRows * kW * 2, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int i = 0; i < kRows * kW; i++) bad += out[i] != h[s + i];
            printf("variant %d start %lld: %d mismatches (out[255..257] = %d %d %d want %d %d %d)\n", variant, s, bad, out[255], out[256], out[257],
                   h[s + 255], h[s + 256], h[s + 257]);
            if (variant == 0) bad_total += bad;
        }
    }
    printf(bad_total ? "FAIL\n" : "PASS\n");
    return bad_total != 0;
}

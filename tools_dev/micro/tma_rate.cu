// Micro-benchmark: per-SM TMA ingest rate (L2 -> shared memory) as a function of the box's inner extent (32 / 64 / 128 bytes per row).
// One thread per CTA streams boxes of ROWS x INNER bytes from an L2-resident [rows][256 B] array into a 4-slot shared-memory ring and
// waits for each slot's mbarrier before re-using it; no consumer.  Prints bytes per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_rate tma_rate.cu -lcuda && ./tma_rate
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int INNER, int ROWS, int SLOTS>
__global__ void __launch_bounds__(32, 1) k(const __grid_constant__ CUtensorMap map, int iters, int total_rows, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bars[SLOTS];
    constexpr int BYTES = INNER * ROWS;
    if (threadIdx.x == 0) {
        for (int s = 0; s < SLOTS; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        int row = (blockIdx.x * 977) % (total_rows - ROWS);
        for (int it = 0; it < iters + SLOTS; ++it) {
            const int s = it % SLOTS;
            if (it >= SLOTS) {      // wait for the previous load of this slot
                const uint32_t ph = ((it / SLOTS) - 1) & 1;
                uint32_t ok = 0;
                while (!ok)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(ok)
                                 : "r"(smem_u32(&bars[s])), "r"(ph)
                                 : "memory");
            }
            if (it < iters) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(BYTES) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                                 smem_u32(smem + s * BYTES)),
                             "l"(&map), "r"(smem_u32(&bars[s])), "r"(0), "r"(row)
                             : "memory");
                row += ROWS;
                if (row + ROWS > total_rows) row = 0;
            }
        }
        cycles[blockIdx.x] = clock64() - t0;
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int INNER, int ROWS, int SLOTS>
void run(EncodeFn enc, void* buf, int total_rows, int grid, long long* d_cycles) {
    CUtensorMap map;
    cuuint64_t dims[2] = {256, (cuuint64_t)total_rows};
    cuuint64_t strides[1] = {256};
    cuuint32_t box[2] = {(cuuint32_t)INNER, (cuuint32_t)ROWS};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapSwizzle sw = INNER == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : INNER == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        printf("encode failed\n");
        return;
    }
    const int smem = SLOTS * INNER * ROWS + 1024;
    cudaFuncSetAttribute(k<INNER, ROWS, SLOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 2000;
    for (int rep = 0; rep < 2; ++rep) k<INNER, ROWS, SLOTS><<<grid, 32, smem>>>(map, iters, total_rows, d_cycles);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, d_cycles, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < grid; ++i) mean += (double)h[i];
    mean /= grid;
    printf("inner %3d B x %3d rows (%5.1f KB box), %d slots, grid %3d: %.1f B/clk/SM, %.3f rows/clk/SM  [%s]\n", INNER, ROWS, INNER * ROWS / 1024.0, SLOTS,
           grid, (double)iters * INNER * ROWS / mean, (double)iters * ROWS / mean, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeFn enc = (EncodeFn)fn;
    const int total_rows = 1 << 17;      // 32 MiB: L2 resident
    void* buf;
    cudaMalloc(&buf, (size_t)total_rows * 256);
    cudaMemset(buf, 1, (size_t)total_rows * 256);
    long long* d_cycles;
    cudaMalloc(&d_cycles, 148 * sizeof(long long));
    for (int grid : {148, 1}) {
        run<32, 256, 4>(enc, buf, total_rows, grid, d_cycles);
        run<64, 256, 4>(enc, buf, total_rows, grid, d_cycles);
        run<128, 256, 4>(enc, buf, total_rows, grid, d_cycles);
        run<64, 128, 8>(enc, buf, total_rows, grid, d_cycles);
        run<128, 128, 8>(enc, buf, total_rows, grid, d_cycles);
        run<128, 64, 8>(enc, buf, total_rows, grid, d_cycles);
    }
    return 0;
}

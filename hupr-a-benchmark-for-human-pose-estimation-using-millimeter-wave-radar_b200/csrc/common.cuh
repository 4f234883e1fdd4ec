// Shared device helpers for the hupr_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "hupr_internal.h"
#ifndef HUPR_OK
#define HUPR_OK 0
#define HUPR_ERR_BAD_ARG (-1)
#define HUPR_ERR_ALIGNMENT (-2)
#define HUPR_ERR_CUDA (-3)
#define HUPR_ERR_ARCH (-4)
#define HUPR_ERR_WORKSPACE (-5)
#endif

namespace hupr {

void note_launches(int n);   // capi.cu: kernel-launch counter behind hupr_launch_count()
// Per-device host-side caches (capi.cu).  A process may drive several GPUs (one after the other or from several threads), and both the
// architecture check and cudaFuncSetAttribute(MaxDynamicSharedMemorySize) are per-device state, so nothing here is cached process-wide.
constexpr int kMaxDevices = 64;
int device_index();            // current CUDA device ordinal, or -1
int device_check_sm100();      // HUPR_OK on an sm_100 device, HUPR_ERR_ARCH otherwise (HUPR_ERR_CUDA if the query fails)
int device_sm_count();         // multiprocessor count of the current device (0 if the query fails)
// Programmatic dependent launch (PDL).  Kernels of the latency-bound inference chain are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization (launch_k below, when pdl_enabled()): the next kernel of the stream may start —
// block scheduling, barrier init, TMEM allocation, tensor-map prefetch — while its predecessor is still running.  Contract for every
// kernel launched through launch_k: all threads execute pdl_wait() before the first access to global memory (it returns once the
// predecessor grid has completed and its writes are visible), then pdl_launch_dependents() so the successor's prologue overlaps this
// kernel's body.  Both instructions are no-ops in a kernel launched without the attribute.
bool pdl_enabled();            // capi.cu: HUPR_PDL environment switch / hupr_set_pdl()
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
static inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    if (pdl_enabled()) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);      // errors surface through cudaGetLastError() at the call site
}
#endif
// Opt-in dynamic shared memory for `func` on the current device, once per (kernel, device): `flags` is a per-kernel static array.
template <typename F>
static inline int ensure_smem_optin(F func, int bytes, bool (&flags)[kMaxDevices]) {
    const int dev = device_index();
    if (dev < 0 || dev >= kMaxDevices) return HUPR_ERR_CUDA;
    if (!flags[dev]) {
        if (cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return HUPR_ERR_CUDA;
        flags[dev] = true;
    }
    return HUPR_OK;
}

__host__ __device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
// a + w*b
__host__ __device__ __forceinline__ float2 cfma(float2 w, float2 b, float2 a) {
    return make_float2(fmaf(-w.y, b.y, fmaf(w.x, b.x, a.x)), fmaf(w.y, b.x, fmaf(w.x, b.y, a.y)));
}
__host__ __device__ __forceinline__ float2 mul_neg_i(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)
__host__ __device__ __forceinline__ float2 mul_pos_i(float2 a) { return make_float2(-a.y, a.x); }   // a * (+i)

// exp(-2*pi*i * num / den)
__device__ __forceinline__ float2 twiddle(int num, int den) {
    float s, c;
    sincospif(-2.0f * (float)num / (float)den, &s, &c);
    return make_float2(c, s);
}

// Forward 4-point DFT (w4 = -i), natural order in/out.
__host__ __device__ __forceinline__ void fft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = csub(a1, a3);
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    a1 = make_float2(t1.x + t3.y, t1.y - t3.x);   // t1 - i*t3
    a3 = make_float2(t1.x - t3.y, t1.y + t3.x);   // t1 + i*t3
}

// Forward 8-point DFT, natural order in/out.
__host__ __device__ __forceinline__ void fft8(float2 (&x)[8]) {
    const float r = 0.70710678118654752440f;
    float2 e0 = x[0], e1 = x[2], e2 = x[4], e3 = x[6];
    float2 o0 = x[1], o1 = x[3], o2 = x[5], o3 = x[7];
    fft4(e0, e1, e2, e3);
    fft4(o0, o1, o2, o3);
    // o_k *= w8^k
    float2 p1 = make_float2((o1.x + o1.y) * r, (o1.y - o1.x) * r);      // *(1-i)/sqrt2
    float2 p2 = mul_neg_i(o2);
    float2 p3 = make_float2((o3.y - o3.x) * r, -(o3.x + o3.y) * r);     // *(-1-i)/sqrt2
    x[0] = cadd(e0, o0); x[4] = csub(e0, o0);
    x[1] = cadd(e1, p1); x[5] = csub(e1, p1);
    x[2] = cadd(e2, p2); x[6] = csub(e2, p2);
    x[3] = cadd(e3, p3); x[7] = csub(e3, p3);
}

// Forward 16-point DFT (radix 4x4, decimation in time), natural order in/out.
__host__ __device__ __forceinline__ void fft16(float2 (&x)[16]) {
    const float r = 0.70710678118654752440f;
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;   // cos/sin(pi/8)
    // stage A: P_t[k_lo] = FFT4 over s of x[4s+t]   (stored in place at x[4*k_lo + t])
#pragma unroll
    for (int t = 0; t < 4; ++t) fft4(x[t], x[4 + t], x[8 + t], x[12 + t]);
    // twiddle R_t[k_lo] = P_t[k_lo] * w16^(t*k_lo)
    x[5]  = cmul(x[5],  make_float2(c1, -s1));                       // w16^1
    x[6]  = make_float2((x[6].x + x[6].y) * r, (x[6].y - x[6].x) * r);       // w16^2
    x[7]  = cmul(x[7],  make_float2(s1, -c1));                       // w16^3
    x[9]  = make_float2((x[9].x + x[9].y) * r, (x[9].y - x[9].x) * r);       // w16^2
    x[10] = mul_neg_i(x[10]);                                        // w16^4
    x[11] = make_float2((x[11].y - x[11].x) * r, -(x[11].x + x[11].y) * r);  // w16^6
    x[13] = cmul(x[13], make_float2(s1, -c1));                       // w16^3
    x[14] = make_float2((x[14].y - x[14].x) * r, -(x[14].x + x[14].y) * r);  // w16^6
    x[15] = cmul(x[15], make_float2(-c1, s1));                       // w16^9
    // stage B: X[k_lo + 4*k_hi] = FFT4 over t of R_t[k_lo]
    float2 y[16];
#pragma unroll
    for (int kl = 0; kl < 4; ++kl) {
        float2 b0 = x[4 * kl], b1 = x[4 * kl + 1], b2 = x[4 * kl + 2], b3 = x[4 * kl + 3];
        fft4(b0, b1, b2, b3);
        y[kl] = b0; y[kl + 4] = b1; y[kl + 8] = b2; y[kl + 12] = b3;
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = y[k];
}

// ---- mbarrier / bulk-copy (TMA engine, non-tensor form) helpers -----------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// 1-D bulk async copy global -> shared (bytes multiple of 16, both 16-B aligned); completes on mbarrier.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// Map a local shared::cta address to the same offset in CTA `rank` of the cluster (shared::cluster window).
__device__ __forceinline__ uint32_t mapa(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f2(uint32_t addr, float2 v) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void st_global_cs_f4(void* p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

}  // namespace hupr

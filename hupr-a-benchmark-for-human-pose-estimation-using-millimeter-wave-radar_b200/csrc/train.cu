// Training-mode element-wise / reduction kernels on channels-last bf16 hi/lo tensors (sm_100a; HBM-bound).
// They implement what autograd + nn.BatchNorm3d(training) + nn.PReLU + F.interpolate backward do around the convolutions of
// /root/reference/models/layers.py (BasicBlock3D :40-70, BasicBlock2D :8-38, Encoder3D :186-217, attention :126-133):
//   hupr_channel_sums    per-channel double-precision sums  S1 = sum f1, S2 = sum f2  for four (f1, f2) modes:
//                          STATS   (z, z^2)                       -> BatchNorm batch statistics
//                          BN_BWD  (g', g' * zhat)                -> dbeta, dgamma and the two means BN backward needs  (g' = g * [y > 0])
//                          PRELU   (g * s * [s <= 0], g)          -> dslope (summed over channels by the caller), bias gradient
//   hupr_affine_act      out = act( s1*z + h1 (+ s2*r + h2) )      BN apply (+ residual branch BN) + ReLU / PReLU
//   hupr_bn_bwd_apply    dz = k1 * (g' - k2 - zhat * k3)
//   hupr_act_bwd         ds = g * (s > 0 ? 1 : slope)             PReLU / ReLU backward (mask from the saved pre- or post-activation)
//   hupr_accumulate      out = split(a (+ b) (+ f32))             gradient fan-in
//   hupr_resample_linear_bwd   adjoint of hupr_resample_linear (atomic float adds into a zero-filled fp32 tensor)
//   hupr_softmax_bwd_rows      dS = P * (dP - sum_m P dP)          (layers.py:131 backward)
#include "common.cuh"
#include "split.cuh"

namespace hupr {

struct TView {                 // channels-last view: element (pos, ch) at pos*ld + off + ch
    const __nv_bfloat16* hi;
    const __nv_bfloat16* lo;
    int ld, off;
};
struct TViewOut {
    __nv_bfloat16* hi;
    __nv_bfloat16* lo;
    int ld, off;
};

__device__ __forceinline__ void tv_load8(const TView& t, size_t pos, int ch, float (&v)[8]) {
    const size_t o = pos * t.ld + t.off + ch;
    load8(t.hi + o, t.lo ? t.lo + o : nullptr, v);
}
__device__ __forceinline__ void tv_store8(const TViewOut& t, size_t pos, int ch, const float (&v)[8]) {
    const size_t o = pos * t.ld + t.off + ch;
    store8(t.hi + o, t.lo ? t.lo + o : nullptr, v);
}

// Raw (undecoded) 8-channel load: both planes are requested back to back with no control flow in between, so a caller can put the loads
// of several tensors / positions in flight before the first conversion (tv_load8's `if (lo)` splits the basic block and the compiler
// then keeps one load pair in flight per thread: channel_sums measured 0.40 of the HBM peak that way, ncu r02_chsums).
struct Raw8 { uint4 h, l; };
__device__ __forceinline__ Raw8 tv_raw8(const TView& t, size_t pos, int ch) {
    const size_t o = pos * t.ld + t.off + ch;
    Raw8 r;
    r.h = __ldg(reinterpret_cast<const uint4*>(t.hi + o));
    r.l = __ldg(reinterpret_cast<const uint4*>((t.lo ? t.lo : t.hi) + o));      // single-plane tensors: a harmless second read of hi, masked below
    if (!t.lo) r.l = make_uint4(0u, 0u, 0u, 0u);
    return r;
}
__device__ __forceinline__ void raw_decode8(const Raw8& r, float (&v)[8]) {
    const uint32_t hw[4] = {r.h.x, r.h.y, r.h.z, r.h.w}, lw[4] = {r.l.x, r.l.y, r.l.z, r.l.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = bf16_lo_f(hw[i]) + bf16_lo_f(lw[i]);
        v[2 * i + 1] = bf16_hi_f(hw[i]) + bf16_hi_f(lw[i]);
    }
}

enum { SUMS_STATS = 0, SUMS_BN_BWD = 1, SUMS_PRELU = 2 };

// grid.x CTAs of 256 threads; thread = (channel group of 8, position lane); each CTA walks a strided range of positions and
// adds its per-channel partial sums (double) with atomics: c*2 atomics per CTA.
template <int MODE>
__global__ void __launch_bounds__(256)
channel_sums_kernel(TView a, TView b, TView m, const float* __restrict__ mean, const float* __restrict__ rstd, long long positions,
                    int c, double* __restrict__ s1, double* __restrict__ s2) {
    extern __shared__ double sred[];                 // [2][lanes][c]
    const int groups = c >> 3;
    const int lanes = 256 / groups;
    const int cg = threadIdx.x % groups, lane = threadIdx.x / groups;
    double acc1[8], acc2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc1[i] = acc2[i] = 0.0;
    float mu[8], rs[8];
    if (MODE == SUMS_BN_BWD) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { mu[i] = __ldg(mean + cg * 8 + i); rs[i] = __ldg(rstd + cg * 8 + i); }
    }
    if (lane < lanes) {
        // two positions per iteration: all loads of both are issued before the (double-precision) accumulation chain consumes them
        const long long stride = (long long)gridDim.x * lanes;
        long long pos = (long long)blockIdx.x * lanes + lane;
        auto accumulate = [&](float (&x)[8], float (&y)[8], float (&k)[8]) {
            if (MODE == SUMS_STATS) {
#pragma unroll
                for (int i = 0; i < 8; ++i) { acc1[i] += x[i]; acc2[i] += (double)x[i] * x[i]; }
            } else if (MODE == SUMS_BN_BWD) {       // a = g, b = z, m = activation output (mask), may be absent
                if (m.hi) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) x[i] = k[i] > 0.f ? x[i] : 0.f;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) { acc1[i] += x[i]; acc2[i] += (double)x[i] * ((y[i] - mu[i]) * rs[i]); }
            } else {                                 // PRELU: a = g, b = pre-activation s (may be absent -> only sum g)
                if (b.hi) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc1[i] += y[i] > 0.f ? 0.0 : (double)x[i] * y[i];
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) acc2[i] += x[i];
            }
        };
        const bool has_b = (MODE != SUMS_STATS) && b.hi != nullptr, has_m = (MODE == SUMS_BN_BWD) && m.hi != nullptr;
        const TView& vb = has_b ? b : a;                   // absent operands alias `a`: the loads stay unconditional, the values unused
        const TView& vm = has_m ? m : a;
        // software-pipelined over positions: the six 16-byte loads of position p + stride are issued BEFORE the conversion / double-precision
        // accumulation chain of position p runs (~300 instructions), so every warp keeps loads in flight while it computes.  (Issuing the
        // loads of two positions and then consuming both left the memory system idle during the arithmetic: 16 resident warps per SM at
        // 127 registers alternate between waiting and computing — ncu: 0.40 of the HBM peak, every stall a long-scoreboard one.)
        if (pos < positions) {
            Raw8 ca = tv_raw8(a, (size_t)pos, cg * 8), cb = ca, cm = ca;
            if (MODE != SUMS_STATS) cb = tv_raw8(vb, (size_t)pos, cg * 8);
            if (MODE == SUMS_BN_BWD) cm = tv_raw8(vm, (size_t)pos, cg * 8);
            for (;;) {
                const long long nxt = pos + stride;
                const bool more = nxt < positions;
                const size_t np = (size_t)(more ? nxt : pos);      // the last iteration re-reads its own (cached) position: loads stay unconditional
                Raw8 na = tv_raw8(a, np, cg * 8), nb = na, nm = na;
                if (MODE != SUMS_STATS) nb = tv_raw8(vb, np, cg * 8);
                if (MODE == SUMS_BN_BWD) nm = tv_raw8(vm, np, cg * 8);
                float x0[8], y0[8], k0[8];
                raw_decode8(ca, x0); raw_decode8(cb, y0); raw_decode8(cm, k0);
                accumulate(x0, y0, k0);
                if (!more) break;
                ca = na; cb = nb; cm = nm;
                pos = nxt;
            }
        }
    }
    double* r1 = sred;
    double* r2 = sred + lanes * c;
    if (lane < lanes) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { r1[lane * c + cg * 8 + i] = acc1[i]; r2[lane * c + cg * 8 + i] = acc2[i]; }
    }
    __syncthreads();
    for (int ch = threadIdx.x; ch < c; ch += 256) {
        double t1 = 0.0, t2 = 0.0;
        for (int l = 0; l < lanes; ++l) { t1 += r1[l * c + ch]; t2 += r2[l * c + ch]; }
        atomicAdd(s1 + ch, t1);
        atomicAdd(s2 + ch, t2);
    }
}

// out = act(s1*z + h1 + (s2*r + h2)); scale/shift pointers may be null (1 / 0); slope null = identity.
__global__ void __launch_bounds__(256)
affine_act_kernel(TView z, const float* __restrict__ s1, const float* __restrict__ h1, TView r, const float* __restrict__ s2,
                  const float* __restrict__ h2, const float* __restrict__ slope, TViewOut out, long long positions, int c) {
    const int groups = c >> 3;
    const long long total = positions * groups;
    for (long long gid = (long long)blockIdx.x * 256 + threadIdx.x; gid < total; gid += (long long)gridDim.x * 256) {
        const int cg = (int)(gid % groups);
        const size_t pos = (size_t)(gid / groups);
        float x[8], y[8];
        tv_load8(z, pos, cg * 8, x);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], s1 ? __ldg(s1 + cg * 8 + i) : 1.f, h1 ? __ldg(h1 + cg * 8 + i) : 0.f);
        if (r.hi) {
            tv_load8(r, pos, cg * 8, y);
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] += fmaf(y[i], s2 ? __ldg(s2 + cg * 8 + i) : 1.f, h2 ? __ldg(h2 + cg * 8 + i) : 0.f);
        }
        if (slope) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = x[i] > 0.f ? x[i] : x[i] * __ldg(slope + cg * 8 + i);
        }
        tv_store8(out, pos, cg * 8, x);
    }
}

// dz = k1 * (g' - k2 - zhat*k3), g' = g*[mask > 0], zhat = (z - mean)*rstd; k1 = gamma*rstd, k2 = mean(g'), k3 = mean(g' zhat).
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(TView g, TView z, TView m, const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ k1,
                    const float* __restrict__ k2, const float* __restrict__ k3, TViewOut out, long long positions, int c) {
    const int groups = c >> 3;
    const long long total = positions * groups;
    for (long long gid = (long long)blockIdx.x * 256 + threadIdx.x; gid < total; gid += (long long)gridDim.x * 256) {
        const int cg = (int)(gid % groups);
        const size_t pos = (size_t)(gid / groups);
        float x[8], y[8], k[8];
        tv_load8(g, pos, cg * 8, x);
        tv_load8(z, pos, cg * 8, y);
        if (m.hi) {
            tv_load8(m, pos, cg * 8, k);
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = k[i] > 0.f ? x[i] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int ch = cg * 8 + i;
            const float zh = (y[i] - __ldg(mean + ch)) * __ldg(rstd + ch);
            x[i] = __ldg(k1 + ch) * (x[i] - __ldg(k2 + ch) - zh * __ldg(k3 + ch));
        }
        tv_store8(out, pos, cg * 8, x);
    }
}

// ds = g * (s > 0 ? 1 : slope[ch])
__global__ void __launch_bounds__(256)
act_bwd_kernel(TView g, TView s, const float* __restrict__ slope, TViewOut out, long long positions, int c) {
    const int groups = c >> 3;
    const long long total = positions * groups;
    for (long long gid = (long long)blockIdx.x * 256 + threadIdx.x; gid < total; gid += (long long)gridDim.x * 256) {
        const int cg = (int)(gid % groups);
        const size_t pos = (size_t)(gid / groups);
        float x[8], y[8];
        tv_load8(g, pos, cg * 8, x);
        tv_load8(s, pos, cg * 8, y);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = y[i] > 0.f ? x[i] : x[i] * __ldg(slope + cg * 8 + i);
        tv_store8(out, pos, cg * 8, x);
    }
}

// out = split(a + b + f)  (b and f optional; f = float [positions][f_ld] at f_off)
__global__ void __launch_bounds__(256)
accumulate_kernel(TView a, TView b, const float* __restrict__ f, int f_ld, int f_off, TViewOut out, long long positions, int c) {
    const int groups = c >> 3;
    const long long total = positions * groups;
    for (long long gid = (long long)blockIdx.x * 256 + threadIdx.x; gid < total; gid += (long long)gridDim.x * 256) {
        const int cg = (int)(gid % groups);
        const size_t pos = (size_t)(gid / groups);
        float x[8], y[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = 0.f;
        if (a.hi) tv_load8(a, pos, cg * 8, x);
        if (b.hi) {
            tv_load8(b, pos, cg * 8, y);
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] += y[i];
        }
        if (f) {
            const float4* fp = reinterpret_cast<const float4*>(f + pos * f_ld + f_off + cg * 8);
            const float4 u = __ldg(fp), v = __ldg(fp + 1);
            x[0] += u.x; x[1] += u.y; x[2] += u.z; x[3] += u.w; x[4] += v.x; x[5] += v.y; x[6] += v.z; x[7] += v.w;
        }
        tv_store8(out, pos, cg * 8, x);
    }
}

// out[pos] = sum_ch a[pos][ch] * (b[pos][ch] - c[pos][ch]); one warp per position, c multiple of 8 (c.hi may be null).
__global__ void __launch_bounds__(256)
rowdot_kernel(TView a, TView b, TView c, float* __restrict__ out, long long positions, int cn) {
    const long long warp = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= positions) return;
    float acc = 0.f;
    for (int g = lane; g < (cn >> 3); g += 32) {
        float x[8], y[8], z[8];
        tv_load8(a, (size_t)warp, g * 8, x);
        tv_load8(b, (size_t)warp, g * 8, y);
        if (c.hi) {
            tv_load8(c, (size_t)warp, g * 8, z);
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] -= z[i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc = fmaf(x[i], y[i], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[warp] = acc;
}

// Per-channel BatchNorm bookkeeping of the training forward in ONE launch (it was ~20 element-wise launches on [C] vectors):
// batch mean / biased variance from the double sums, rstd, the fused affine (scale, shift) of the normalise kernel, and the
// running-statistics update with momentum (unbiased variance), exactly as nn.BatchNorm3d does in train mode.
__global__ void bn_finalize_kernel(const double* __restrict__ s1, const double* __restrict__ s2, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, float* __restrict__ mean, float* __restrict__ rstd,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, long long* __restrict__ num_batches_tracked, int c) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch == 0 && num_batches_tracked) *num_batches_tracked += 1;
    if (ch >= c) return;
    const double m = s1[ch] / count;
    double var = s2[ch] / count - m * m;
    var = var > 0.0 ? var : 0.0;
    const float mf = (float)m;
    const float rs = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[ch] * rs;
    mean[ch] = mf;
    rstd[ch] = rs;
    scale[ch] = sc;
    shift[ch] = beta[ch] - mf * sc;
    if (running_mean) running_mean[ch] = running_mean[ch] * (1.0f - momentum) + momentum * mf;
    if (running_var) {
        const double unbiased = var * (count / (count > 1.0 ? count - 1.0 : 1.0));
        running_var[ch] = running_var[ch] * (1.0f - momentum) + momentum * (float)unbiased;
    }
}

// ... and of the backward: k2 = mean(g'), k3 = mean(g' zhat) for hupr_bn_bwd_apply, dgamma = sum g' zhat, dbeta = sum g'.
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ t1, const double* __restrict__ t2, double count, float* __restrict__ k2,
                                       float* __restrict__ k3, float* __restrict__ dgamma, float* __restrict__ dbeta, int c) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= c) return;
    k2[ch] = (float)(t1[ch] / count);
    k3[ch] = (float)(t2[ch] / count);
    dgamma[ch] = (float)t2[ch];
    dbeta[ch] = (float)t1[ch];
}

struct ResampleBwdParams {
    int n, di, hi, wi, dout, ho, wo, c8;
    int g_ld, g_off, o_ld, o_off;
    float sd, sh, sw;
    int vec4;
};

// Adjoint of resample_kernel: every OUTPUT element of the forward scatters its gradient to its (up to) 8 sources.
__global__ void __launch_bounds__(256)
resample_bwd_kernel(TView g, float* __restrict__ din, const ResampleBwdParams p) {
    const size_t total = (size_t)p.n * p.dout * p.ho * p.wo * p.c8;
    for (size_t gid = (size_t)blockIdx.x * 256 + threadIdx.x; gid < total; gid += (size_t)gridDim.x * 256) {
        size_t t = gid;
        const int cg = (int)(t % p.c8); t /= p.c8;
        const int ow = (int)(t % p.wo); t /= p.wo;
        const int oh = (int)(t % p.ho); t /= p.ho;
        const int od = (int)(t % p.dout);
        const int n = (int)(t / p.dout);
        const float fd = p.sd * od, fh = p.sh * oh, fw = p.sw * ow;
        const int d0 = (int)fd, h0 = (int)fh, w0 = (int)fw;
        const int d1 = d0 + (d0 < p.di - 1), h1 = h0 + (h0 < p.hi - 1), w1 = w0 + (w0 < p.wi - 1);
        const float ld1 = fd - d0, lh1 = fh - h0, lw1 = fw - w0;
        const float wd[2] = {p.di == 1 ? 1.f : 1.f - ld1, ld1}, wh[2] = {1.f - lh1, lh1}, ww[2] = {1.f - lw1, lw1};
        const int ds[2] = {d0, d1}, hs[2] = {h0, h1}, ws[2] = {w0, w1};
        float v[8];
        tv_load8(g, gid / p.c8, cg * 8, v);
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            if (a == 1 && p.di == 1) break;
#pragma unroll
            for (int b = 0; b < 2; ++b) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const float wgt = wd[a] * wh[b] * ww[c];
                    if (wgt == 0.f) continue;
                    float* dst = din + ((((size_t)n * p.di + ds[a]) * p.hi + hs[b]) * p.wi + ws[c]) * p.o_ld + p.o_off + cg * 8;
                    if (p.vec4) {       // 16-byte aligned rows: two vector reductions instead of eight scalar atomics
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(wgt * v[0]), "f"(wgt * v[1]), "f"(wgt * v[2]),
                                     "f"(wgt * v[3])
                                     : "memory");
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(wgt * v[4]), "f"(wgt * v[5]), "f"(wgt * v[6]),
                                     "f"(wgt * v[7])
                                     : "memory");
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) atomicAdd(dst + i, wgt * v[i]);
                    }
                }
            }
        }
    }
}

// dS = P * (dP - rowdot), rowdot = sum_m P[m] dP[m]; one CTA per row, cols <= 4096, cols % 4 == 0.
__global__ void __launch_bounds__(256)
softmax_bwd_rows_kernel(const __nv_bfloat16* __restrict__ p_hi, const __nv_bfloat16* __restrict__ p_lo, const float* __restrict__ dp,
                        __nv_bfloat16* __restrict__ ds_hi, __nv_bfloat16* __restrict__ ds_lo, int cols) {
    __shared__ float red[8];
    const size_t row = blockIdx.x;
    const int nvec = cols >> 2, tid = threadIdx.x;
    const uint2* ph = reinterpret_cast<const uint2*>(p_hi + row * cols);
    const uint2* pl = reinterpret_cast<const uint2*>(p_lo + row * cols);
    const float4* dpr = reinterpret_cast<const float4*>(dp + row * cols);
    float4 pv[4], gv[4];
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = tid + k * 256;
        if (i < nvec) {
            const uint2 h = __ldg(ph + i), l = __ldg(pl + i);
            pv[k] = make_float4(bf16_lo_f(h.x) + bf16_lo_f(l.x), bf16_hi_f(h.x) + bf16_hi_f(l.x), bf16_lo_f(h.y) + bf16_lo_f(l.y),
                                bf16_hi_f(h.y) + bf16_hi_f(l.y));
            gv[k] = __ldg(dpr + i);
            dot += pv[k].x * gv[k].x + pv[k].y * gv[k].y + pv[k].z * gv[k].z + pv[k].w * gv[k].w;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if ((tid & 31) == 0) red[tid >> 5] = dot;
    __syncthreads();
    dot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) dot += red[w];
    uint2* oh = reinterpret_cast<uint2*>(ds_hi + row * cols);
    uint2* ol = reinterpret_cast<uint2*>(ds_lo + row * cols);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = tid + k * 256;
        if (i < nvec) {
            uint32_t h0, l0, h1, l1;
            split2(pv[k].x * (gv[k].x - dot), pv[k].y * (gv[k].y - dot), h0, l0);
            split2(pv[k].z * (gv[k].z - dot), pv[k].w * (gv[k].w - dot), h1, l1);
            oh[i] = make_uint2(h0, h1);
            ol[i] = make_uint2(l0, l1);
        }
    }
}

static int train_check_sm100() {
    return device_check_sm100();      // cached per device (capi.cu)
}

static inline TView mk_view(const hupr_tensor_view* v) {
    TView t;
    if (v) { t.hi = (const __nv_bfloat16*)v->hi; t.lo = (const __nv_bfloat16*)v->lo; t.ld = v->ld; t.off = v->ch_off; }
    else { t.hi = nullptr; t.lo = nullptr; t.ld = 0; t.off = 0; }
    return t;
}
static inline TViewOut mk_out(const hupr_tensor_view* v) {
    TViewOut t;
    t.hi = (__nv_bfloat16*)v->hi; t.lo = (__nv_bfloat16*)v->lo; t.ld = v->ld; t.off = v->ch_off;
    return t;
}
static inline bool view_ok(const hupr_tensor_view* v, int c, bool required) {
    if (!v || !v->hi) return !required;
    if (v->ld % 8 || v->ch_off % 8 || v->ch_off < 0 || v->ch_off + c > v->ld) return false;
    return (((uintptr_t)v->hi | (uintptr_t)v->lo) & 15) == 0;
}
static inline unsigned ew_blocks(long long total) {
    long long b = (total + 255) / 256;
    if (b > 148LL * 32) b = 148LL * 32;
    return (unsigned)(b < 1 ? 1 : b);
}
static inline int finish(int n) {
    note_launches(n);
    return cudaGetLastError() == cudaSuccess ? HUPR_OK : HUPR_ERR_CUDA;
}

}  // namespace hupr

using namespace hupr;

extern "C" int hupr_channel_sums(int mode, const hupr_tensor_view* a, const hupr_tensor_view* b, const hupr_tensor_view* mask,
                                 const float* mean, const float* rstd, long long positions, int c, double* s1, double* s2, void* stream) {
    if (positions <= 0 || c <= 0 || c % 8 || c > 2048 || !s1 || !s2) return HUPR_ERR_BAD_ARG;
    if (!view_ok(a, c, true) || !view_ok(b, c, mode == SUMS_BN_BWD) || !view_ok(mask, c, false)) return HUPR_ERR_BAD_ARG;
    if (mode == SUMS_BN_BWD && (!mean || !rstd)) return HUPR_ERR_BAD_ARG;
    int rc = train_check_sm100();
    if (rc != HUPR_OK) return rc;
    const int groups = c / 8;
    if (groups > 256) return HUPR_ERR_BAD_ARG;
    const int lanes = 256 / groups;
    const size_t smem = (size_t)2 * lanes * c * sizeof(double);
    long long blocks = (positions + lanes - 1) / lanes;
    if (blocks > 148 * 6) blocks = 148 * 6;
    cudaStream_t s = (cudaStream_t)stream;
    const TView va = mk_view(a), vb = mk_view(b && b->hi ? b : nullptr), vm = mk_view(mask && mask->hi ? mask : nullptr);
    if (mode == SUMS_STATS) channel_sums_kernel<SUMS_STATS><<<(unsigned)blocks, 256, smem, s>>>(va, vb, vm, mean, rstd, positions, c, s1, s2);
    else if (mode == SUMS_BN_BWD) channel_sums_kernel<SUMS_BN_BWD><<<(unsigned)blocks, 256, smem, s>>>(va, vb, vm, mean, rstd, positions, c, s1, s2);
    else if (mode == SUMS_PRELU) channel_sums_kernel<SUMS_PRELU><<<(unsigned)blocks, 256, smem, s>>>(va, vb, vm, mean, rstd, positions, c, s1, s2);
    else return HUPR_ERR_BAD_ARG;
    return finish(1);
}

extern "C" int hupr_affine_act(const hupr_tensor_view* z, const float* scale1, const float* shift1, const hupr_tensor_view* r,
                               const float* scale2, const float* shift2, const float* slope, const hupr_tensor_view* out, long long positions,
                               int c, void* stream) {
    if (positions <= 0 || c <= 0 || c % 8 || !view_ok(z, c, true) || !view_ok(r, c, false) || !view_ok(out, c, true)) return HUPR_ERR_BAD_ARG;
    int rc = train_check_sm100();
    if (rc != HUPR_OK) return rc;
    affine_act_kernel<<<ew_blocks(positions * (c / 8)), 256, 0, (cudaStream_t)stream>>>(mk_view(z), scale1, shift1, mk_view(r && r->hi ? r : nullptr),
                                                                                        scale2, shift2, slope, mk_out(out), positions, c);
    return finish(1);
}

extern "C" int hupr_bn_bwd_apply(const hupr_tensor_view* g, const hupr_tensor_view* z, const hupr_tensor_view* mask, const float* mean,
                                 const float* rstd, const float* k1, const float* k2, const float* k3, const hupr_tensor_view* out,
                                 long long positions, int c, void* stream) {
    if (positions <= 0 || c <= 0 || c % 8 || !view_ok(g, c, true) || !view_ok(z, c, true) || !view_ok(mask, c, false) || !view_ok(out, c, true))
        return HUPR_ERR_BAD_ARG;
    if (!mean || !rstd || !k1 || !k2 || !k3) return HUPR_ERR_BAD_ARG;
    int rc = train_check_sm100();
    if (rc != HUPR_OK) return rc;
    bn_bwd_apply_kernel<<<ew_blocks(positions * (c / 8)), 256, 0, (cudaStream_t)stream>>>(
        mk_view(g), mk_view(z), mk_view(mask && mask->hi ? mask : nullptr), mean, rstd, k1, k2, k3, mk_out(out), positions, c);
    return finish(1);
}

extern "C" int hupr_act_bwd(const hupr_tensor_view* g, const hupr_tensor_view* s, const float* slope, const hupr_tensor_view* out,
                            long long positions, int c, void* stream) {
    if (positions <= 0 || c <= 0 || c % 8 || !slope || !view_ok(g, c, true) || !view_ok(s, c, true) || !view_ok(out, c, true)) return HUPR_ERR_BAD_ARG;
    int rc = train_check_sm100();
    if (rc != HUPR_OK) return rc;
    act_bwd_kernel<<<ew_blocks(positions * (c / 8)), 256, 0, (cudaStream_t)stream>>>(mk_view(g), mk_view(s), slope, mk_out(out), positions, c);
    return finish(1);
}

extern "C" int hupr_accumulate(const hupr_tensor_view* a, const hupr_tensor_view* b, const float* f, int f_ld, int f_off,
                               const hupr_tensor_view* out, long long positions, int c, void* stream) {
    if (positions <= 0 || c <= 0 || c % 8 || !view_ok(a, c, false) || !view_ok(b, c, false) || !view_ok(out, c, true)) return HUPR_ERR_BAD_ARG;
    if (f && (f_ld % 4 || f_off % 4 || f_off + c > f_ld || ((uintptr_t)f & 15))) return HUPR_ERR_BAD_ARG;
    int rc = train_check_sm100();
    if (rc != HUPR_OK) return rc;
    accumulate_kernel<<<ew_blocks(positions * (c / 8)), 256, 0, (cudaStream_t)stream>>>(mk_view(a && a->hi ? a : nullptr), mk_view(b && b->hi ? b : nullptr),
                                                                                        f, f_ld, f_off, mk_out(out), positions, c);
    return finish(1);
}

extern "C" int hupr_bn_finalize(const double* s1, const double* s2, long long count, const float* gamma, const float* beta, float eps, float momentum,
                                float* mean, float* rstd, float* scale, float* shift, float* running_mean, float* running_var,
                                long long* num_batches_tracked, int c, void* stream) {
    if (!s1 || !s2 || !gamma || !beta || !mean || !rstd || !scale || !shift || c <= 0 || count <= 0) return HUPR_ERR_BAD_ARG;
    int rc = train_check_sm100();
    if (rc != HUPR_OK) return rc;
    bn_finalize_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(s1, s2, (double)count, gamma, beta, eps, momentum, mean, rstd, scale, shift,
                                                                          running_mean, running_var, num_batches_tracked, c);
    return finish(1);
}

extern "C" int hupr_bn_bwd_finalize(const double* t1, const double* t2, long long count, float* k2, float* k3, float* dgamma, float* dbeta, int c,
                                    void* stream) {
    if (!t1 || !t2 || !k2 || !k3 || !dgamma || !dbeta || c <= 0 || count <= 0) return HUPR_ERR_BAD_ARG;
    int rc = train_check_sm100();
    if (rc != HUPR_OK) return rc;
    bn_bwd_finalize_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(t1, t2, (double)count, k2, k3, dgamma, dbeta, c);
    return finish(1);
}

extern "C" int hupr_rowdot(const hupr_tensor_view* a, const hupr_tensor_view* b, const hupr_tensor_view* c, float* out, long long positions,
                           int c_n, void* stream) {
    if (positions < 0 || c_n <= 0 || c_n % 8 || !out) return HUPR_ERR_BAD_ARG;
    if (positions == 0) return HUPR_OK;
    if (!view_ok(a, c_n, true) || !view_ok(b, c_n, true) || !view_ok(c, c_n, false)) return HUPR_ERR_BAD_ARG;
    int rc = train_check_sm100();
    if (rc != HUPR_OK) return rc;
    const long long blocks = (positions + 7) / 8;
    if (blocks > 2147483647LL) return HUPR_ERR_BAD_ARG;
    rowdot_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(mk_view(a), mk_view(b), mk_view(c && c->hi ? c : nullptr), out, positions, c_n);
    return finish(1);
}

extern "C" int hupr_resample_linear_bwd(const hupr_tensor_view* g, int n, int dout, int ho, int wo, int c, float* din, int di, int hi, int wi,
                                        int in_ld, int in_ch_off, void* stream) {
    if (!din || n <= 0 || di <= 0 || hi <= 0 || wi <= 0 || dout <= 0 || ho <= 0 || wo <= 0 || c <= 0 || c % 8) return HUPR_ERR_BAD_ARG;
    if (!view_ok(g, c, true) || in_ch_off < 0 || in_ch_off + c > in_ld) return HUPR_ERR_BAD_ARG;
    int rc = train_check_sm100();
    if (rc != HUPR_OK) return rc;
    ResampleBwdParams p;
    p.n = n; p.di = di; p.hi = hi; p.wi = wi; p.dout = dout; p.ho = ho; p.wo = wo; p.c8 = c / 8;
    p.g_ld = g->ld; p.g_off = g->ch_off; p.o_ld = in_ld; p.o_off = in_ch_off;
    p.sd = dout > 1 ? (float)(di - 1) / (float)(dout - 1) : 0.f;
    p.sh = ho > 1 ? (float)(hi - 1) / (float)(ho - 1) : 0.f;
    p.sw = wo > 1 ? (float)(wi - 1) / (float)(wo - 1) : 0.f;
    p.vec4 = (in_ld % 4 == 0 && in_ch_off % 4 == 0 && ((uintptr_t)din & 15) == 0) ? 1 : 0;
    resample_bwd_kernel<<<ew_blocks((long long)n * dout * ho * wo * p.c8), 256, 0, (cudaStream_t)stream>>>(mk_view(g), din, p);
    return finish(1);
}

extern "C" int hupr_softmax_bwd_rows(const void* p_hi, const void* p_lo, const float* dp, void* ds_hi, void* ds_lo, long long rows, int cols,
                                     void* stream) {
    if (rows < 0 || rows > 2147483647LL) return HUPR_ERR_BAD_ARG;
    if (rows == 0) return HUPR_OK;
    if (!p_hi || !p_lo || !dp || !ds_hi || !ds_lo || cols <= 0 || cols % 4 || cols > 4096) return HUPR_ERR_BAD_ARG;
    int rc = train_check_sm100();
    if (rc != HUPR_OK) return rc;
    softmax_bwd_rows_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)p_hi, (const __nv_bfloat16*)p_lo, dp,
                                                                            (__nv_bfloat16*)ds_hi, (__nv_bfloat16*)ds_lo, cols);
    return finish(1);
}

// ================================================================================================================================
// Backward pieces of the PRGCN head and of MNet (reference: autograd of models/gcn_networks.py:23-64, models/chirp_networks.py:17-21).
// ================================================================================================================================
namespace hupr {

constexpr int kTJ = 14;

// dbias[q][j] = sum_b dYt[(b, j)][q]      (dYt: bf16 split rows [(b, j)][1024])
__global__ void __launch_bounds__(256)
gcn_bias_grad_kernel(const __nv_bfloat16* __restrict__ g_hi, const __nv_bfloat16* __restrict__ g_lo, float* __restrict__ dbias, int batch) {
    const int gid = blockIdx.x * 256 + threadIdx.x;       // (j, q)
    if (gid >= kTJ * 1024) return;
    const int j = gid >> 10, q = gid & 1023;
    float s = 0.f;
    for (int b = 0; b < batch; ++b) {
        const size_t o = (size_t)(b * kTJ + j) * 1024 + q;
        s += __bfloat162float(g_hi[o]) + __bfloat162float(g_lo[o]);
    }
    dbias[q * kTJ + j] = s;
}

// Adjoint of gcn_heads (bilinear x2 of the 32x32 maps): d_pre float [B][14][64][64] -> dYt float rows [(b, j)][1024] (+=, zero-filled).
__global__ void __launch_bounds__(256)
gcn_heads_bwd_kernel(const float* __restrict__ d_pre, float* __restrict__ dyt, int batch) {
    const int gid = blockIdx.x * 256 + threadIdx.x;
    if (gid >= batch * kTJ * 4096) return;
    const int pos = gid & 4095, bj = gid >> 12;
    const int oh = pos >> 6, ow = pos & 63;
    const float scale = 31.0f / 63.0f;
    const float fh = scale * oh, fw = scale * ow;
    const int h0 = (int)fh, w0 = (int)fw;
    const int h1 = h0 + (h0 < 31), w1 = w0 + (w0 < 31);
    const float lh1 = fh - h0, lw1 = fw - w0, lh0 = 1.f - lh1, lw0 = 1.f - lw1;
    const float g = __ldg(d_pre + gid);
    float* m = dyt + (size_t)bj * 1024;
    atomicAdd(m + h0 * 32 + w0, lh0 * lw0 * g);
    atomicAdd(m + h0 * 32 + w1, lh0 * lw1 * g);
    atomicAdd(m + h1 * 32 + w0, lh1 * lw0 * g);
    atomicAdd(m + h1 * 32 + w1, lh1 * lw1 * g);
}

// Adjoint of gcn_nodes' node path: dSt0 (bf16 split rows [(b, j')][1024]) -> dx[b][node][k] = sum_j' A[k][j'] dSt0[(b, j')][node]
// -> scattered through the bilinear x0.5 stencil into d_logits float channels-last [B][4096][ld] (+=).
__global__ void __launch_bounds__(256)
gcn_nodes_bwd_kernel(const __nv_bfloat16* __restrict__ s_hi, const __nv_bfloat16* __restrict__ s_lo, const float* __restrict__ adj,
                     float* __restrict__ d_logits, int ld, int batch) {
    __shared__ float sA[kTJ * kTJ];
    if (threadIdx.x < kTJ * kTJ) sA[threadIdx.x] = __ldg(adj + threadIdx.x);
    __syncthreads();
    const int gid = blockIdx.x * 256 + threadIdx.x;
    if (gid >= batch * 1024) return;
    const int b = gid >> 10, node = gid & 1023;
    float ds[kTJ];
#pragma unroll
    for (int j = 0; j < kTJ; ++j) {
        const size_t o = (size_t)(b * kTJ + j) * 1024 + node;
        ds[j] = __bfloat162float(s_hi[o]) + __bfloat162float(s_lo[o]);
    }
    const int oh = node >> 5, ow = node & 31;
    const float scale = 63.0f / 31.0f;
    const float fh = scale * oh, fw = scale * ow;
    const int h0 = (int)fh, w0 = (int)fw;
    const int h1 = h0 + (h0 < 63), w1 = w0 + (w0 < 63);
    const float lh1 = fh - h0, lw1 = fw - w0, lh0 = 1.f - lh1, lw0 = 1.f - lw1;
    float* base = d_logits + (size_t)b * 4096 * ld;
#pragma unroll
    for (int k = 0; k < kTJ; ++k) {
        float dx = 0.f;
#pragma unroll
        for (int j = 0; j < kTJ; ++j) dx = fmaf(sA[k * kTJ + j], ds[j], dx);
        atomicAdd(base + (size_t)(h0 * 64 + w0) * ld + k, lh0 * lw0 * dx);
        atomicAdd(base + (size_t)(h0 * 64 + w1) * ld + k, lh0 * lw1 * dx);
        atomicAdd(base + (size_t)(h1 * 64 + w0) * ld + k, lh1 * lw0 * dx);
        atomicAdd(base + (size_t)(h1 * 64 + w1) * ld + k, lh1 * lw1 * dx);
    }
}

// MNet backward: the max over the 4 chirp pairs routes each output gradient to one pair; weights [32][2][2] and bias [32] gradients
// are reduced over all positions (double atomics, 160 per CTA).
__global__ void __launch_bounds__(256)
mnet_bwd_kernel(const float* __restrict__ vrdae, const float* __restrict__ weight, const float* __restrict__ bias,
                const __nv_bfloat16* __restrict__ g_hi, const __nv_bfloat16* __restrict__ g_lo, double* __restrict__ dw /* [32][4] */,
                double* __restrict__ db /* [32] */, int n_slots) {
    __shared__ float sW[32 * 4 + 32];
    __shared__ float sRed[8][160];
    if (threadIdx.x < 160) sW[threadIdx.x] = threadIdx.x < 128 ? __ldg(weight + threadIdx.x) : __ldg(bias + threadIdx.x - 128);
    __syncthreads();
    const size_t gid = (size_t)blockIdx.x * 256 + threadIdx.x;
    const bool live = gid < (size_t)n_slots * 4096;
    float m[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) m[j] = 0.f;
    if (live) {
        const size_t slot = gid / 4096;
        const int pos = (int)(gid % 4096);
        const float4* base = reinterpret_cast<const float4*>(vrdae + slot * (16 * 4096 * 8) + (size_t)pos * 8);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float4 a = __ldg(base + (size_t)j * (4096 * 2));
            const float4 b = __ldg(base + (size_t)j * (4096 * 2) + 1);
            m[j] = (((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w))) * 0.125f;
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // the thread's 32 output gradients = one 64-byte row per plane: eight 16-byte loads up front.  (Reading them as 2-byte scalars inside the
    // channel loop made every warp-level load touch 32 sectors for 64 useful bytes, 64 times: the kernel was bound by L1 sector requests,
    // 295 us per launch for 134 MB of input.)
    float gv[32];
    {
        const size_t row = live ? gid : 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float g8[8];
            load8(g_hi + row * 32 + q * 8, g_lo ? g_lo + row * 32 + q * 8 : nullptr, g8);
#pragma unroll
            for (int e = 0; e < 8; ++e) gv[q * 8 + e] = g8[e];
        }
    }
#pragma unroll
    for (int o = 0; o < 32; ++o) {
        float c[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        if (live) {
            const float w00 = sW[o * 4], w01 = sW[o * 4 + 1], w10 = sW[o * 4 + 2], w11 = sW[o * 4 + 3], bb = sW[128 + o];
            float best = -INFINITY;
            int bt = 0;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float y = fmaf(w11, m[9 + 2 * t], fmaf(w10, m[8 + 2 * t], fmaf(w01, m[2 * t + 1], w00 * m[2 * t]))) + bb;
                if (y > best) { best = y; bt = t; }       // first maximum wins (torch max_pool3d backward)
            }
            const float g = gv[o];
            float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                if (t == bt) { m0 = m[2 * t]; m1 = m[2 * t + 1]; m2 = m[8 + 2 * t]; m3 = m[9 + 2 * t]; }
            }
            c[0] = g * m0; c[1] = g * m1; c[2] = g * m2; c[3] = g * m3; c[4] = g;
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) {
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) c[k] += __shfl_xor_sync(0xffffffffu, c[k], s);
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 5; ++k) sRed[warp][o * 5 + k] = c[k];
        }
    }
    __syncthreads();
    if (threadIdx.x < 160) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += sRed[w][threadIdx.x];
        const int o = threadIdx.x / 5, k = threadIdx.x % 5;
        if (k < 4) atomicAdd(dw + o * 4 + k, (double)s);
        else atomicAdd(db + o, (double)s);
    }
}

}  // namespace hupr

extern "C" int hupr_gcn_bias_grad(const void* g_hi, const void* g_lo, float* dbias, int batch, void* stream) {
    if (batch <= 0 || !g_hi || !g_lo || !dbias) return HUPR_ERR_BAD_ARG;
    int rc = train_check_sm100();
    if (rc != HUPR_OK) return rc;
    gcn_bias_grad_kernel<<<(kTJ * 1024 + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)g_hi, (const __nv_bfloat16*)g_lo, dbias, batch);
    return finish(1);
}

extern "C" int hupr_gcn_heads_bwd(const float* d_pre, float* dyt, int batch, void* stream) {
    if (batch <= 0 || !d_pre || !dyt) return HUPR_ERR_BAD_ARG;
    int rc = train_check_sm100();
    if (rc != HUPR_OK) return rc;
    gcn_heads_bwd_kernel<<<(batch * kTJ * 4096 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_pre, dyt, batch);
    return finish(1);
}

extern "C" int hupr_gcn_nodes_bwd(const void* s_hi, const void* s_lo, const float* adj, float* d_logits, int ld, int batch, void* stream) {
    if (batch <= 0 || !s_hi || !s_lo || !adj || !d_logits || ld < kTJ) return HUPR_ERR_BAD_ARG;
    int rc = train_check_sm100();
    if (rc != HUPR_OK) return rc;
    gcn_nodes_bwd_kernel<<<(batch * 1024 + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)s_hi, (const __nv_bfloat16*)s_lo, adj,
                                                                                      d_logits, ld, batch);
    return finish(1);
}

extern "C" int hupr_mnet_bwd(const float* vrdae, const float* weight, const float* bias, const void* g_hi, const void* g_lo, double* dweight,
                             double* dbias, int n_slots, void* stream) {
    if (n_slots <= 0 || !vrdae || !weight || !bias || !g_hi || !dweight || !dbias) return HUPR_ERR_BAD_ARG;
    int rc = train_check_sm100();
    if (rc != HUPR_OK) return rc;
    const size_t total = (size_t)n_slots * 4096;
    mnet_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(vrdae, weight, bias, (const __nv_bfloat16*)g_hi,
                                                                                       (const __nv_bfloat16*)g_lo, dweight, dbias, n_slots);
    return finish(1);
}

// Micro-benchmarks, part 2: packed fp32 (FFMA2 / FADD2 / FMUL2, new on sm_100) against scalar FFMA for the
// FFT butterfly mix, alone and interleaved with shared-memory loads.  Answers: does packing two butterflies
// into one instruction raise the butterfly rate of an issue-bound FFT kernel on B200?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench2 ubench2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 2048;

__global__ void k_ffma2(float* out, float a, float b) {
    float2 r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = make_float2(threadIdx.x * 0.001f + i, i - threadIdx.x * 0.002f);
    const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.999f);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = __ffma2_rn(r[i], a2, b2);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += r[i].x + r[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// three distinct register-pair sources
__global__ void k_ffma2_3reg(float* out, float a0) {
    float2 r[8], a[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = make_float2(threadIdx.x * 0.001f + i, i - threadIdx.x * 0.002f);
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = make_float2(a0 + i * 1e-6f, a0 - i * 1e-6f);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = __ffma2_rn(r[(i + 3) & 7], a[i & 3], r[i]);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += r[i].x + r[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// scalar butterfly mix on 16 complex values (re[], im[]): 4 stages x 8 butterflies x 6 FMA
__global__ void k_bfly(float* out, float wr, float wi) {
    float re[16], im[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { re[i] = threadIdx.x * 0.001f + i; im[i] = i - threadIdx.x * 0.002f; }
    for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
        for (int s = 1; s < 16; s <<= 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if ((i & s) == 0) {
                    int j = i | s;
                    float tr = fmaf(-wi, im[j], fmaf(wr, re[j], re[i]));
                    float ti = fmaf(wi, re[j], fmaf(wr, im[j], im[i]));
                    re[j] = fmaf(2.f, re[i], -tr); im[j] = fmaf(2.f, im[i], -ti);
                    re[i] = tr; im[i] = ti;
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += re[i] + im[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// packed butterfly mix: the same 16 complex values as 8 pairs (x[i], x[i+8]); stages with span 1,2,4 are packed
// (6 FFMA2 per two butterflies), the span-8 stage works inside the pairs with scalar FMAs.
__global__ void k_bfly2(float* out, float wr, float wi) {
    float2 re[8], im[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        re[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.001f + i + 8);
        im[i] = make_float2(i - threadIdx.x * 0.002f, i + 8 - threadIdx.x * 0.002f);
    }
    const float2 WR = make_float2(wr, wr), WI = make_float2(wi, wi), NWI = make_float2(-wi, -wi), TWO = make_float2(2.f, 2.f);
    for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
        for (int s = 1; s < 8; s <<= 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if ((i & s) == 0) {
                    int j = i | s;
                    float2 tr = __ffma2_rn(NWI, im[j], __ffma2_rn(WR, re[j], re[i]));
                    float2 ti = __ffma2_rn(WI, re[j], __ffma2_rn(WR, im[j], im[i]));
                    float2 nr, ni;
                    nr.x = -tr.x; nr.y = -tr.y; ni.x = -ti.x; ni.y = -ti.y;
                    re[j] = __ffma2_rn(TWO, re[i], nr); im[j] = __ffma2_rn(TWO, im[i], ni);
                    re[i] = tr; im[i] = ti;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {   // span 8: inside the pair
            float tr = fmaf(-wi, im[i].y, fmaf(wr, re[i].y, re[i].x));
            float ti = fmaf(wi, re[i].y, fmaf(wr, im[i].y, im[i].x));
            re[i].y = fmaf(2.f, re[i].x, -tr); im[i].y = fmaf(2.f, im[i].x, -ti);
            re[i].x = tr; im[i].x = ti;
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += re[i].x + im[i].x + re[i].y + im[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// butterfly mixes with one conflict-free LDS.64 per butterfly pair folded in (the fused kernel's LSU : FP ratio)
template <bool kPacked>
__global__ void k_bfly_lds(float* out, float wr, float wi) {
    __shared__ float2 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float2(1e-3f * i, -1e-3f * i);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    float2 re[8], im[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        re[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.001f + i + 8);
        im[i] = make_float2(i - threadIdx.x * 0.002f, i + 8 - threadIdx.x * 0.002f);
    }
    const float2 WR = make_float2(wr, wr), WI = make_float2(wi, wi), NWI = make_float2(-wi, -wi), TWO = make_float2(2.f, 2.f);
    for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
        for (int s = 1; s < 16; s <<= 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (s < 8 ? (i & s) == 0 : (i & 1) == 0) {
                    const int j = s < 8 ? (i | s) : (i | 1);
                    const float2 t = sm[(lane + 32 * (i + s) + it) & 2047];
                    if (kPacked) {
                        float2 tr = __ffma2_rn(NWI, im[j], __ffma2_rn(WR, re[j], re[i]));
                        float2 ti = __ffma2_rn(WI, re[j], __ffma2_rn(WR, im[j], im[i]));
                        float2 nr, ni;
                        nr.x = -tr.x; nr.y = -tr.y; ni.x = -ti.x; ni.y = -ti.y;
                        re[j] = __ffma2_rn(TWO, re[i], nr); im[j] = __ffma2_rn(TWO, im[i], ni);
                        re[i] = __fadd2_rn(tr, t); im[i] = ti;
                    } else {
                        float trx = fmaf(-wi, im[j].x, fmaf(wr, re[j].x, re[i].x)), try_ = fmaf(-wi, im[j].y, fmaf(wr, re[j].y, re[i].y));
                        float tix = fmaf(wi, re[j].x, fmaf(wr, im[j].x, im[i].x)), tiy = fmaf(wi, re[j].y, fmaf(wr, im[j].y, im[i].y));
                        re[j].x = fmaf(2.f, re[i].x, -trx); re[j].y = fmaf(2.f, re[i].y, -try_);
                        im[j].x = fmaf(2.f, im[i].x, -tix); im[j].y = fmaf(2.f, im[i].y, -tiy);
                        re[i].x = trx + t.x; re[i].y = try_ + t.y; im[i].x = tix; im[i].y = tiy;
                    }
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += re[i].x + im[i].x + re[i].y + im[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F launch, int reps = 5) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("device %s sms=%d clock=%d kHz\n", p.name, p.multiProcessorCount, clk_khz);
    const int sms = p.multiProcessorCount;
    float* out; CK(cudaMalloc(&out, sizeof(float) * sms * 8 * 1024));
    for (int warps = 4; warps <= 32; warps *= 2) {
        const int thr = warps * 32 < 256 ? warps * 32 : 256;
        const int blk = warps * 32 < 256 ? sms : sms * (warps * 32 / 256);
        const double nthreads = (double)blk * thr;
        const double per = 1.0 / sms / (clk_khz * 1e3);
        float ms;
        ms = time_ms([&] { k_ffma2<<<blk, thr>>>(out, 1.0001f, 0.5f); });
        printf("warps/SM=%2d ffma2(2 src pairs const) %7.1f FMA/clk/SM  %6.1f inst/clk/SM\n", warps, nthreads * ITERS * 16 / (ms * 1e-3) * per,
               nthreads * ITERS * 8 / (ms * 1e-3) * per);
        ms = time_ms([&] { k_ffma2_3reg<<<blk, thr>>>(out, 1.0001f); });
        printf("warps/SM=%2d ffma2(3 reg pairs)       %7.1f FMA/clk/SM  %6.1f inst/clk/SM\n", warps, nthreads * ITERS * 16 / (ms * 1e-3) * per,
               nthreads * ITERS * 8 / (ms * 1e-3) * per);
        const double bf = nthreads * (ITERS / 4) * 4 * 8;   // butterflies
        ms = time_ms([&] { k_bfly<<<blk, thr>>>(out, 0.999f, 0.01f); });
        printf("warps/SM=%2d bfly scalar              %7.2f bfly/clk/SM\n", warps, bf / (ms * 1e-3) * per);
        ms = time_ms([&] { k_bfly2<<<blk, thr>>>(out, 0.999f, 0.01f); });
        printf("warps/SM=%2d bfly packed(3/4 stages)  %7.2f bfly/clk/SM\n", warps, bf / (ms * 1e-3) * per);
        ms = time_ms([&] { k_bfly_lds<false><<<blk, thr>>>(out, 0.999f, 0.01f); });
        printf("warps/SM=%2d bfly scalar + lds.64     %7.2f bfly/clk/SM\n", warps, bf / (ms * 1e-3) * per);
        ms = time_ms([&] { k_bfly_lds<true><<<blk, thr>>>(out, 0.999f, 0.01f); });
        printf("warps/SM=%2d bfly packed + lds.64     %7.2f bfly/clk/SM\n", warps, bf / (ms * 1e-3) * per);
    }
    cudaFree(out);
    return 0;
}

// Micro-benchmarks that decide the STFT route ("chosen by evidence").
// Measures, per SM and chip-wide on the B200 it runs on:
//   * FFMA issue rate (3-register form, independent chains)
//   * legacy mma.sync m16n8k16 f16->f32 and m16n8k8 tf32->f32 rate
//   * SHFL and LDS.64 rate
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;

__global__ void k_ffma(float* out, float a, float b) {
    float r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = fmaf(r[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FFMA where all three source operands are distinct registers (RF bandwidth test)
__global__ void k_ffma3(float* out, float a0) {
    float r[16], a[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = threadIdx.x * 0.001f + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = a0 + i * 1e-6f;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = fmaf(r[(i + 5) & 15], a[i & 3], r[i]);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_mma_f16(float* out) {
    uint32_t a[4] = {0x3c003c00u + threadIdx.x, 0x3c003c00u, 0x3c003c00u ^ threadIdx.x, 0x3c003c00u};
    uint32_t b[2] = {0x38003800u, 0x38003800u + threadIdx.x};
    float c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = threadIdx.x; c[i][2] = i * 2.f; c[i][3] = i + 1.f; }
    #pragma unroll 2
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_mma_tf32(float* out) {
    uint32_t a[4] = {0x3f800000u + threadIdx.x, 0x3f800000u, 0x3f800000u ^ threadIdx.x, 0x3f800000u};
    uint32_t b[2] = {0x3f000000u, 0x3f000000u + threadIdx.x};
    float c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = threadIdx.x; c[i][2] = i * 2.f; c[i][3] = i + 1.f; }
    #pragma unroll 2
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FFT-like mix: radix-2 butterflies with FMA twiddles on a 32-value register array
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

__global__ void k_shfl(float* out) {
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = threadIdx.x + i;
    int src = (32 - (threadIdx.x & 31)) & 31;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = __shfl_sync(0xffffffffu, r[i], src);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_lds64(float* out) {
    __shared__ float2 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float2(i, -i);
    __syncthreads();
    float2 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(0.f, 0.f);
    int base = threadIdx.x & 31;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float2 v = sm[(base + i * 32 + it) & 1023];
            acc[i].x += v.x; acc[i].y += v.y;
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
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
    printf("device %s sms=%d cc=%d.%d clock=%d kHz smem/blk optin=%zu regs/sm=%d\n", p.name, p.multiProcessorCount,
           p.major, p.minor, clk_khz, p.sharedMemPerBlockOptin, p.regsPerMultiprocessor);
    const int sms = p.multiProcessorCount;
    float* out; CK(cudaMalloc(&out, sizeof(float) * sms * 8 * 1024));
    for (int warps = 4; warps <= 32; warps *= 2) {
        const int threads = 256, blocks = sms * (warps * 32 / threads > 0 ? warps * 32 / threads : 1);
        const int thr = warps * 32 < 256 ? warps * 32 : 256;
        const int blk = warps * 32 < 256 ? sms : blocks;
        const double nthreads = (double)blk * thr;
        float ms;
        ms = time_ms([&] { k_ffma<<<blk, thr>>>(out, 1.0001f, 0.5f); });
        printf("warps/SM=%2d ffma(imm-ish)   %8.2f TFLOP/s  %6.1f FMA/clk/SM@%dMHz\n", warps, 2.0 * nthreads * ITERS * 16 / ms / 1e9,
               nthreads * ITERS * 16 / (ms * 1e-3) / sms / (clk_khz * 1e3), clk_khz / 1000);
        ms = time_ms([&] { k_ffma3<<<blk, thr>>>(out, 1.0001f); });
        printf("warps/SM=%2d ffma(3reg)      %8.2f TFLOP/s  %6.1f FMA/clk/SM\n", warps, 2.0 * nthreads * ITERS * 16 / ms / 1e9,
               nthreads * ITERS * 16 / (ms * 1e-3) / sms / (clk_khz * 1e3));
        ms = time_ms([&] { k_bfly<<<blk, thr>>>(out, 0.999f, 0.01f); });
        printf("warps/SM=%2d fft-bfly fma mix %8.2f Ginst/s  %6.1f thread-inst/clk/SM\n", warps, nthreads * (ITERS / 4) * 4 * 8 * 6 / ms / 1e6,
               nthreads * (ITERS / 4) * 4 * 8 * 6 / (ms * 1e-3) / sms / (clk_khz * 1e3));
        ms = time_ms([&] { k_mma_f16<<<blk, thr>>>(out); });
        printf("warps/SM=%2d mma.sync f16    %8.2f TFLOP/s  %6.1f MAC/clk/SM\n", warps, 2.0 * (nthreads / 32) * ITERS * 8 * 2048 / ms / 1e9,
               (nthreads / 32) * ITERS * 8 * 2048 / (ms * 1e-3) / sms / (clk_khz * 1e3));
        ms = time_ms([&] { k_mma_tf32<<<blk, thr>>>(out); });
        printf("warps/SM=%2d mma.sync tf32   %8.2f TFLOP/s  %6.1f MAC/clk/SM\n", warps, 2.0 * (nthreads / 32) * ITERS * 8 * 1024 / ms / 1e9,
               (nthreads / 32) * ITERS * 8 * 1024 / (ms * 1e-3) / sms / (clk_khz * 1e3));
        ms = time_ms([&] { k_shfl<<<blk, thr>>>(out); });
        printf("warps/SM=%2d shfl            %8.2f Gwarp-inst/s  %6.3f warp-inst/clk/SM\n", warps, (nthreads / 32) * ITERS * 8 / ms / 1e6,
               (nthreads / 32) * ITERS * 8 / (ms * 1e-3) / sms / (clk_khz * 1e3));
        ms = time_ms([&] { k_lds64<<<blk, thr>>>(out); });
        printf("warps/SM=%2d lds.64          %8.2f TB/s  %6.1f B/clk/SM\n", warps, nthreads * ITERS * 8 * 8 / ms / 1e9,
               nthreads * ITERS * 8 * 8 / (ms * 1e-3) / sms / (clk_khz * 1e3));
    }
    cudaFree(out);
    return 0;
}

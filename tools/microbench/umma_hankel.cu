// tcgen05 micro-benchmark for the "next" STFT route.  Two questions, answered on a B200:
//   1. Does a K-major, no-swizzle shared-memory descriptor with OVERLAPPING rows (leading-dimension byte offset 16, stride byte
//      offset 128: row f starts 8 bf16 after row f-1) read a Hankel matrix A[f][k] = P[8 f + k] straight out of a flat sample
//      array?  That is "framing for free": hop 256 = 8 * 32, so frame f of the stride-32 de-interleaved samples is exactly that.
//   2. How fast are the small-N MMAs this route needs (M = 128 frames, N = 16 / 32 real outputs, K = 16), i.e. is SS-mode
//      operand fetch (A re-read from shared memory for every MMA) the bound?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_hankel umma_hankel.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE matrix descriptor (cute::UMMA::SmemDescriptor): 8-row x 16-byte core matrices, rows 16 bytes apart;
// lbo = byte distance between the two core matrices of one K = 16 step, sbo = byte distance between 8-row groups.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version of sm_100
    return d;
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// bounded wait (a wrong descriptor must not hang the box): returns false after ~2^24 polls
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    for (int spin = 0; spin < (1 << 24) && !ok; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
    return ok != 0;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

constexpr int kFrames = 128, kK = 32, kN = 32;
constexpr int kPLen = 8 * (kFrames - 1) + kK;   // 1048 samples of the de-interleaved stream

// mode 0: Hankel A straight out of P (lbo 16, sbo 128); mode 1: A materialised frame by frame in the standard canonical layout.
// D[f][n] = sum_k A[f][k] * B[n][k], 128 x 32 x 32, two K = 16 MMAs.
__global__ void __launch_bounds__(128) k_check(const __nv_bfloat16* __restrict__ P, const __nv_bfloat16* __restrict__ B, float* __restrict__ D, int mode) {
    __shared__ __align__(128) __nv_bfloat16 sP[1088];
    __shared__ __align__(128) __nv_bfloat16 sA[kFrames * kK];     // standard layout: (f/8)*128 B + (k/8)*2048 B + (f%8)*16 B + (k%8)*2 B
    __shared__ __align__(128) __nv_bfloat16 sB[kN * kK];          // (n/8)*512 B + (k/8)*128 B + (n%8)*16 B + (k%8)*2 B
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 1088; i += 128) sP[i] = i < kPLen ? P[i] : __float2bfloat16(0.f);
    for (int i = tid; i < kFrames * kK; i += 128) {
        const int f = i / kK, k = i % kK;
        sA[(f / 8) * 64 + (k / 8) * 1024 + (f % 8) * 8 + (k % 8)] = P[8 * f + k];
    }
    for (int i = tid; i < kN * kK; i += 128) {
        const int n = i / kK, k = i % kK;
        sB[(n / 8) * 256 + (k / 8) * 64 + (n % 8) * 8 + (k % 8)] = B[i];
    }
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(32));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // operand tiles written with st.shared -> visible to the MMA's async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = make_idesc(128, kN);
        for (int ks = 0; ks < 2; ++ks) {   // K = 16 per MMA: two core-matrix columns
            const uint64_t ad = mode == 0 ? make_desc(smem_u32(sP) + ks * 32, 16, 128) : make_desc(smem_u32(sA) + ks * 4096, 2048, 128);
            const uint64_t bd = make_desc(smem_u32(sB) + ks * 256, 128, 512);
            mma_bf16_ss(tmem, ad, bd, idesc, ks > 0);
        }
        mma_commit(&bar);
    }
    const bool done = mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[32];
    if (!done && tid == 0) printf("k_check mode %d: MMA completion never arrived\n", mode);
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), r);
#pragma unroll
    for (int n = 0; n < 32; ++n) D[tid * kN + n] = __uint_as_float(r[n]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
}

// Issue `iters` MMAs of shape M x N x 16 per issuing warp and time them on the SM clock.  `nwarps` warps issue concurrently
// (lane 0 of each), every warp rotating over 4 accumulators of its own (column blocks of N), so consecutive MMAs are
// independent.  hankel = 1: A descriptor with overlapping rows (lbo 16, sbo 128); 0: standard canonical layout (lbo 2048).
__global__ void __launch_bounds__(128) k_rate(int M, int N, int hankel, int iters, int nwarps, long long* cycles) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar[4];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (8192 + 8192) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // finite bf16 pattern
    if (tid == 0) { for (int w = 0; w < 4; ++w) mbar_init(&bar[w], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    long long dt = 0;
    if (warp < nwarps && lane == 0) {
        const uint32_t idesc = make_idesc(M, N);
        const uint64_t ad = hankel ? make_desc(smem_u32(smem), 16, 128) : make_desc(smem_u32(smem), 2048, 128);
        const uint64_t bd = make_desc(smem_u32(smem + 8192), 128, 256);   // N x 16, K-major: two core-matrix columns 128 B apart
        const uint32_t d0 = tmem + (uint32_t)(warp * 4 * N);
        const long long t0 = clock64();
        for (int i = 0; i < iters; i += 4) {
            mma_bf16_ss(d0, ad, bd, idesc, i > 0);
            mma_bf16_ss(d0 + N, ad, bd, idesc, i > 0);
            mma_bf16_ss(d0 + 2 * N, ad, bd, idesc, i > 0);
            mma_bf16_ss(d0 + 3 * N, ad, bd, idesc, i > 0);
        }
        mma_commit(&bar[warp]);
        const bool done = mbar_wait(&bar[warp], 0);
        dt = done ? clock64() - t0 : -1;
    }
    if (tid == 0) cycles[blockIdx.x] = dt;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s sms=%d\n", prop.name, prop.multiProcessorCount);
    // ---------------- 1. correctness of the overlapping-row descriptor ----------------
    std::vector<__nv_bfloat16> hP(kPLen), hB(kN * kK);
    std::vector<float> fP(kPLen), fB(kN * kK);
    uint32_t s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xFFFF) / 65536.0f - 0.5f; };
    for (int i = 0; i < kPLen; ++i) { fP[i] = bf16_round(rnd()); hP[i] = __float2bfloat16(fP[i]); }
    for (int i = 0; i < kN * kK; ++i) { fB[i] = bf16_round(rnd()); hB[i] = __float2bfloat16(fB[i]); }
    __nv_bfloat16 *dP, *dB;
    float* dD;
    CK(cudaMalloc(&dP, sizeof(__nv_bfloat16) * kPLen));
    CK(cudaMalloc(&dB, sizeof(__nv_bfloat16) * kN * kK));
    CK(cudaMalloc(&dD, sizeof(float) * kFrames * kN));
    CK(cudaMemcpy(dP, hP.data(), sizeof(__nv_bfloat16) * kPLen, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), sizeof(__nv_bfloat16) * kN * kK, cudaMemcpyHostToDevice));
    for (int mode = 0; mode < 2; ++mode) {
        CK(cudaMemset(dD, 0, sizeof(float) * kFrames * kN));
        k_check<<<1, 128>>>(dP, dB, dD, mode);
        CK(cudaDeviceSynchronize());
        std::vector<float> hD(kFrames * kN);
        CK(cudaMemcpy(hD.data(), dD, sizeof(float) * kFrames * kN, cudaMemcpyDeviceToHost));
        double maxerr = 0.0, maxref = 0.0;
        for (int f = 0; f < kFrames; ++f)
            for (int n = 0; n < kN; ++n) {
                double ref = 0.0;
                for (int k = 0; k < kK; ++k) ref += (double)fP[8 * f + k] * (double)fB[n * kK + k];
                maxerr = fmax(maxerr, fabs(ref - (double)hD[f * kN + n]));
                maxref = fmax(maxref, fabs(ref));
            }
        printf("check %-28s 128x32x32 bf16: max |D - ref| = %.3e (max |ref| = %.3f)  %s\n",
               mode == 0 ? "Hankel A (lbo 16, sbo 128)" : "standard A (lbo 2048)", maxerr, maxref, maxerr < 1e-4 ? "OK" : "MISMATCH");
    }
    // ---------------- 2. issue rate of small-N MMAs ----------------
    long long* dC;
    CK(cudaMalloc(&dC, sizeof(long long) * prop.multiProcessorCount));
    CK(cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
    const int iters = 4096;
    for (int hankel = 1; hankel >= 0; --hankel)
        for (int nwarps : {1, 2, 4})
            for (int M : {128, 64})
                for (int N : {16, 32, 64, 128}) {
                    if (nwarps * 4 * N > 512) continue;
                    k_rate<<<prop.multiProcessorCount, 128, 16384>>>(M, N, hankel, iters, nwarps, dC);
                    CK(cudaDeviceSynchronize());
                    std::vector<long long> hc(prop.multiProcessorCount);
                    CK(cudaMemcpy(hc.data(), dC, sizeof(long long) * hc.size(), cudaMemcpyDeviceToHost));
                    long long mx = 0;
                    for (auto c : hc) mx = c > mx ? c : mx;
                    const double cyc = (double)mx / (iters * nwarps);   // SM cycles per MMA with nwarps issuing
                    printf("rate  %-8s issuing warps=%d M=%3d N=%3d K=16 : %7.2f cycles/MMA  %8.1f MAC/clk/SM  (floor max(M,128)*N/256 = %5.1f)\n",
                           hankel ? "hankel" : "standard", nwarps, M, N, cyc, (double)M * N * 16 / cyc, (double)(M > 128 ? M : 128) * N / 256.0);
                }
    cudaFree(dP); cudaFree(dB); cudaFree(dD); cudaFree(dC);
    return 0;
}

// How fast can the row kernel's DATA MOVEMENT alone go?  Same staging as de_trial_rows_kernel -- per row four TMA
// bulk copies (the row + three rows at random indices) into a 2-stage shared-memory ring per warp, mbarrier completion
// -- and a minimal consumer: each lane adds its 16-byte vectors of the four rows and stores the sum row (5 x row bytes
// of traffic per row).  No Philox, no objective, no selector / index chase.
//    nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_gather_peak tma_gather_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int S>
__global__ void __launch_bounds__(640, 1) tma_gather(const double2 *__restrict__ A, double2 *__restrict__ B, int nrows, int V, uint32_t seed) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int NC = blockDim.x >> 5, c = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rowb = (uint32_t)V * 16u, stageb = 4u * rowb;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)NC * S * stageb);
    if (threadIdx.x == 0) {
        for (int i = 0; i < NC * S; ++i) mbar_init(smem_u32(bars + i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int G = gridDim.x * NC, first = blockIdx.x * NC + c;
    const uint32_t bar0 = smem_u32(bars + c * S), dst0 = smem_u32(smem) + (uint32_t)(c * S) * stageb;
    const double2 *st0 = reinterpret_cast<const double2 *>(smem + (size_t)(c * S) * stageb);
    auto issue = [&](int r, int s) {
        const uint32_t i1 = hash32((uint32_t)r * 3u + seed) % nrows, i2 = hash32((uint32_t)r * 3u + 1u + seed) % nrows,
                       i3 = hash32((uint32_t)r * 3u + 2u + seed) % nrows;
        const uint32_t full = bar0 + 8u * s, dst = dst0 + (uint32_t)s * stageb;
        mbar_expect(full, 4u * rowb);
        bulk(dst, A + (int64_t)r * V, rowb, full);
        bulk(dst + rowb, A + (int64_t)i1 * V, rowb, full);
        bulk(dst + 2 * rowb, A + (int64_t)i2 * V, rowb, full);
        bulk(dst + 3 * rowb, A + (int64_t)i3 * V, rowb, full);
    };
    int la = first;
    for (int s = 0; s < S; ++s, la += G)
        if (la < nrows && lane == 0) issue(la, s);
    int s = 0;
    uint32_t parity = 0;
    for (int r = first; r < nrows; r += G) {
        while (!mbar_try(bar0 + 8u * s, parity)) {}
        const double2 *st = st0 + (size_t)s * 4 * V;
        double2 *q = B + (int64_t)r * V;
        for (int v = lane; v < V; v += 32) {
            const double2 a = st[v], b = st[V + v], cc = st[2 * V + v], d = st[3 * V + v];
            q[v] = make_double2(a.x + b.x + cc.x + d.x, a.y + b.y + cc.y + d.y);
        }
        __syncwarp();
        if (la < nrows && lane == 0) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); issue(la, s); }
        la += G;
        if (++s == S) { s = 0; parity ^= 1u; }
    }
}

int main() {
    const int cols[] = {100, 120, 128, 200, 256, 504, 1000};
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaFuncSetAttribute(tma_gather<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    for (int D : cols) {
        const int V = D / 2;
        const int nrows = (int)(6.0e9 / (D * 8.0));
        double2 *A, *B;
        cudaMalloc(&A, (size_t)nrows * V * sizeof(double2));
        cudaMalloc(&B, (size_t)nrows * V * sizeof(double2));
        cudaMemset(A, 0, (size_t)nrows * V * sizeof(double2));
        const size_t per_warp = 2 * (4 * (size_t)V * 16 + 8);
        int warps = (int)(232448 / per_warp);
        if (warps > 20) warps = 20;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e30f;
        for (int rep = 0; rep < 6; ++rep) {
            cudaEventRecord(e0);
            tma_gather<2><<<sms, warps * 32, warps * per_warp>>>(A, B, nrows, V, 999u + rep);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        printf("columns %5d (row %5d B)  %2d warps/SM  %8.3f ms  %8.1f GB/s  %6.2f M rows/s/SM  %s\n", D, D * 8, warps, best,
               nrows * 5.0 * D * 8 / best / 1e6, nrows / best / 1e3 / sms, cudaGetErrorString(cudaGetLastError()));
        cudaFree(A); cudaFree(B);
    }
    return 0;
}

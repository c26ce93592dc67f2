// Attainable HBM bandwidth of the DE trial's ACCESS PATTERN with no arithmetic: per row i read the row itself and
// three rows at random indices, write one row (5 x row bytes of traffic per row, 3/5 of it random segments of the
// row width).  This is the roofline a narrow-row generation can be held to: random 800-byte segments do not
// stream from DRAM the way 8 KB ones do.    nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_peak gather_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// one warp per row, 16-byte vectors, V vectors per row
__global__ void __launch_bounds__(1024) gather(const double2 *__restrict__ A, double2 *__restrict__ B, int64_t nrows, int V, uint32_t seed) {
    const int lane = threadIdx.x & 31;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < nrows; r += nw) {
        const int64_t i1 = hash32((uint32_t)r * 3u + seed) % nrows, i2 = hash32((uint32_t)r * 3u + 1u + seed) % nrows,
                      i3 = hash32((uint32_t)r * 3u + 2u + seed) % nrows;
        const double2 *p0 = A + r * V, *p1 = A + i1 * V, *p2 = A + i2 * V, *p3 = A + i3 * V;
        double2 *q = B + r * V;
        for (int v = lane; v < V; v += 32) {
            const double2 a = p0[v], b = p1[v], c = p2[v], d = p3[v];
            q[v] = make_double2(a.x + b.x + c.x + d.x, a.y + b.y + c.y + d.y);
        }
    }
}

int main() {
    const int cols[] = {100, 120, 128, 200, 256, 504, 1000};
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int D : cols) {
        const int V = D / 2;
        const int64_t nrows = (int64_t)(6.0e9 / (D * 8.0));           // ~6 GB per array
        double2 *A, *B;
        cudaMalloc(&A, nrows * V * sizeof(double2));
        cudaMalloc(&B, nrows * V * sizeof(double2));
        cudaMemset(A, 0, nrows * V * sizeof(double2));
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e30f;
        for (int rep = 0; rep < 6; ++rep) {
            cudaEventRecord(e0);
            gather<<<sms * 2, 1024>>>(A, B, nrows, V, 12345u + rep);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        printf("columns %5d (row %5d B)  rows %9lld  %8.3f ms  %8.1f GB/s (5 x row bytes per row: 2 sequential + 3 random)\n", D, D * 8,
               (long long)nrows, best, nrows * 5.0 * D * 8 / best / 1e6);
        cudaFree(A); cudaFree(B);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

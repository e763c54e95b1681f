// Throughput of the legacy tensor path (mma.sync) on sm_100a: m16n8k8 vs m16n8k16, bf16 -> fp32.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 scratch/hmma_bench.cu -o scratch/hmma_bench && ./scratch/hmma_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int K16, int CHAINS>
__global__ void bench(float* out, long long* clk, int iters) {
    float d[CHAINS][4];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { d[c][0] = d[c][1] = d[c][2] = d[c][3] = 0.f; }
    uint32_t a0 = threadIdx.x * 0x10001u, a1 = a0 + 3, a2 = a0 + 5, a3 = a0 + 7, b0 = a0 ^ 0x5555u, b1 = b0 + 11;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (K16)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3]) : "r"(a0), "r"(a1), "r"(b0));
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <int K16, int CHAINS>
void run(int warps) {
    float* out; long long* clk; long long h;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
    const int iters = 2000;
    bench<K16, CHAINS><<<148, warps * 32>>>(out, clk, iters);
    bench<K16, CHAINS><<<148, warps * 32>>>(out, clk, iters);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    const double per = (double)h / (iters * CHAINS);          // clocks per MMA of one warp (all warps run the same count)
    const double per_smsp = per / (warps / 4.0);              // warps / 4 warps share a sub-partition
    printf("m16n8k%-2d chains=%d warps/SM=%2d: %.2f clk per MMA and warp, %.2f clk per MMA and sub-partition -> %.0f flop/clk/SM\n", K16 ? 16 : 8, CHAINS, warps, per,
           per_smsp, (K16 ? 4096.0 : 2048.0) / per_smsp * 4);
    cudaFree(out); cudaFree(clk);
}

int main() {
    run<0, 1>(4); run<0, 8>(4); run<0, 8>(16); run<0, 8>(32);
    run<1, 1>(4); run<1, 8>(4); run<1, 8>(16); run<1, 8>(32);
    return 0;
}

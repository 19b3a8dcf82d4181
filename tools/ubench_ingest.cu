// Micro-benchmark behind the persistent decode kernel's tile shape: how fast can the SMs ingest a weight stream from
// HBM while they also re-read an L2-resident activation tile, and what does a grid-wide barrier cost?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_ingest tools/ubench_ingest.cu && ./ubench_ingest
//
// Each CTA (one per SM) runs a TMA producer thread that fills a ring of stages with (a) `w_bytes` of a stream nobody else reads
// (distinct per CTA and iteration: HBM) and (b) `x_bytes` of a small buffer every CTA reads (L2 hits), optionally multicast across a
// cluster; a consumer thread frees each stage as soon as it has landed.  No math: this is the ceiling of the copy engine path.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_load_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r; }
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar_cluster_addr) { asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory"); }

struct Params {
    const uint8_t* w;      // HBM stream
    size_t w_total;        // bytes available
    const uint8_t* x;      // L2-resident buffer
    size_t x_total;
    int w_bytes, x_bytes;  // per stage
    int stages, iters;
    int csz;               // cluster size; > 1: the x tile is multicast (each CTA issues x_bytes / csz to all members)
    int w_chunk;           // bytes per bulk copy of the weight stream (<= w_bytes)
};

__global__ void __launch_bounds__(128, 1) ingest_kernel(const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t stage_bytes = p.w_bytes + p.x_bytes;
    const uint32_t bar = base + p.stages * stage_bytes;   // full[s] @ +8s, empty[s] @ +128+8s
    const int csz = p.csz;
    const uint32_t rank = csz > 1 ? cluster_rank() : 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(bar + 8 * s, 1);
            mbar_init(bar + 128 + 8 * s, csz);   // every CTA of the cluster frees the stage (its x slice is written by peers)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (csz > 1) cluster_sync();
    const size_t cta = blockIdx.x;
    if (threadIdx.x == 0) {
        // producer
        for (int i = 0; i < p.iters; ++i) {
            const int s = i % p.stages;
            const uint32_t ph = (uint32_t)(i / p.stages) & 1u;
            if (i >= p.stages) mbar_wait(bar + 128 + 8 * s, ph ^ 1u);
            const uint32_t st = base + s * stage_bytes;
            mbar_expect(bar + 8 * s, stage_bytes);
            size_t woff = ((cta * p.iters + i) * (size_t)p.w_bytes) % (p.w_total - p.w_bytes);
            woff &= ~(size_t)127;
            for (int c = 0; c < p.w_bytes; c += p.w_chunk) bulk_load(st + c, p.w + woff + c, p.w_chunk, bar + 8 * s);
            if (p.x_bytes) {
                const size_t xoff = ((size_t)(i % 64) * p.x_bytes) % (p.x_total - p.x_bytes);
                if (csz == 1) {
                    bulk_load(st + p.w_bytes, p.x + xoff, p.x_bytes, bar + 8 * s);
                } else {
                    const uint32_t slice = p.x_bytes / csz;
                    bulk_load_mc(st + p.w_bytes + rank * slice, p.x + xoff + rank * slice, slice, bar + 8 * s, (uint16_t)((1u << csz) - 1));
                }
            }
        }
    } else if (threadIdx.x == 32) {
        // consumer: free the stage as soon as it landed (in every CTA of the cluster: peers write into it)
        for (int i = 0; i < p.iters; ++i) {
            const int s = i % p.stages;
            const uint32_t ph = (uint32_t)(i / p.stages) & 1u;
            mbar_wait(bar + 8 * s, ph);
            if (csz == 1) mbar_arrive(bar + 128 + 8 * s);
            else for (int r = 0; r < csz; ++r) mbar_arrive_remote(mapa(bar + 128 + 8 * s, r));
        }
    }
    __syncthreads();
    if (csz > 1) cluster_sync();
}

// grid barrier: monotonically increasing counter, every CTA adds 1 and spins until it reaches round * gridDim
__global__ void __launch_bounds__(128, 1) barrier_kernel(unsigned int* counter, int rounds, float* sink, int payload) {
    extern __shared__ uint8_t smem_raw[];
    float acc = 0.f;
    for (int r = 1; r <= rounds; ++r) {
        if (payload) {  // every CTA publishes 512 B and later reads everybody's: the data path of a phase hand-off
            sink[(size_t)blockIdx.x * 128 + threadIdx.x] = (float)r;
            __threadfence();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
            const unsigned int target = (unsigned int)r * gridDim.x;
            unsigned int v;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while (v < target);
        }
        __syncthreads();
        if (payload) {
            for (int b = threadIdx.x; b < (int)gridDim.x * 128; b += 128 * 8) acc += __ldcg(sink + b);
        }
    }
    if (acc == -1.f) sink[0] = acc;
}

int main() {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    printf("device %s, %d SMs\n", prop.name, prop.multiProcessorCount);
    const size_t w_total = (size_t)3 << 30;
    const size_t x_total = (size_t)4 << 20;
    uint8_t *w, *x;
    CK(cudaMalloc(&w, w_total));
    CK(cudaMalloc(&x, x_total));
    CK(cudaMemset(w, 1, w_total));
    CK(cudaMemset(x, 2, x_total));
    CK(cudaFuncSetAttribute(ingest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CK(cudaFuncSetAttribute(ingest_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    struct Cfg { int ctas, w_bytes, x_bytes, stages, csz, w_chunk; };
    std::vector<Cfg> cfgs;
    for (int ctas : {128, 148}) {
        for (int x_kb : {0, 8, 16, 32}) {
            const int w_kb = 32;
            const int stages = 192 / (w_kb + x_kb);
            cfgs.push_back({ctas, w_kb * 1024, x_kb * 1024, stages, 1, 16 * 1024});
        }
    }
    // ring depth with the weight stream alone (how many bytes in flight per SM does HBM need?)
    for (int stages : {1, 2, 3, 4, 6}) cfgs.push_back({148, 32 * 1024, 0, stages, 1, 16 * 1024});
    // chunk size of the copies
    for (int chunk_kb : {4, 8, 32}) cfgs.push_back({148, 32 * 1024, 0, 6, 1, chunk_kb * 1024});
    // multicast of the shared tile across clusters of 2 / 4 (grids that are multiples of the cluster size)
    for (int csz : {2, 4}) {
        for (int x_kb : {16, 32}) {
            const int ctas = csz == 2 ? 148 : 144;
            cfgs.push_back({ctas, 32 * 1024, x_kb * 1024, 192 / (32 + x_kb), csz, 16 * 1024});
        }
    }
    // L2-resident tile alone (ceiling of L2 -> SM)
    cfgs.push_back({148, 0, 32 * 1024, 6, 1, 16 * 1024});
    cfgs.push_back({148, 0, 32 * 1024, 6, 2, 16 * 1024});
    printf("%5s %7s %7s %6s %4s %6s | %9s %10s %10s %10s\n", "ctas", "w_KB", "x_KB", "stages", "csz", "chunk", "ms", "W GB/s", "X GB/s", "per-SM B/ns");
    for (const Cfg& c : cfgs) {
        Params p;
        p.w = w; p.w_total = w_total; p.x = x; p.x_total = x_total; p.w_bytes = c.w_bytes; p.x_bytes = c.x_bytes; p.stages = c.stages;
        p.csz = c.csz; p.w_chunk = c.w_chunk > c.w_bytes ? (c.w_bytes ? c.w_bytes : 1) : c.w_chunk;
        p.iters = 2000;
        const size_t smem = (size_t)c.stages * (c.w_bytes + c.x_bytes) + 1024 + 512;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(c.ctas); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = c.csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0));
            CK(cudaLaunchKernelEx(&cfg, ingest_kernel, p));
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        const double wb = (double)c.ctas * p.iters * c.w_bytes, xb = (double)c.ctas * p.iters * c.x_bytes;
        printf("%5d %7d %7d %6d %4d %6d | %9.3f %10.1f %10.1f %10.1f\n", c.ctas, c.w_bytes / 1024, c.x_bytes / 1024, c.stages, c.csz, p.w_chunk / 1024, best,
               wb / best * 1e-6, xb / best * 1e-6, (wb + xb) / best * 1e-6 / c.ctas);
        fflush(stdout);
    }
    // grid barrier cost
    unsigned int* counter;
    float* sink;
    CK(cudaFuncSetAttribute(barrier_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaMalloc(&counter, 4));
    CK(cudaMalloc(&sink, 148 * 128 * 4 * 2));
    for (int payload : {0, 1}) {
        for (int ctas : {128, 148}) {
            CK(cudaMemset(counter, 0, 4));
            const int rounds = 2000;
            CK(cudaEventRecord(e0));
            barrier_kernel<<<ctas, 128, 200 * 1024>>>(counter, rounds, sink, payload);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("grid barrier, %d CTAs, payload %d: %.3f us per round\n", ctas, payload, ms * 1e3 / rounds);
        }
    }
    return 0;
}

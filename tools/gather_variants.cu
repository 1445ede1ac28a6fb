// Which load flavour makes a random 32-byte probe cost ONE 32-byte DRAM sector?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_variants gather_variants.cu
// Run under: ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_requests_srcunit_tex_op_read.sum ./gather_variants
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

template <int V>
__device__ __forceinline__ uint32_t probe(const uint4* p) {
    uint32_t a, b, c, d, e = 0, f = 0, g = 0, h = 0;
    if (V == 0) asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a),"=r"(b),"=r"(c),"=r"(d),"=r"(e),"=r"(f),"=r"(g),"=r"(h) : "l"(p));
    if (V == 1) asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a),"=r"(b),"=r"(c),"=r"(d),"=r"(e),"=r"(f),"=r"(g),"=r"(h) : "l"(p));
    if (V == 2) { asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a),"=r"(b),"=r"(c),"=r"(d) : "l"(p));
                  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e),"=r"(f),"=r"(g),"=r"(h) : "l"(p + 1)); }
    if (V == 3) asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a),"=r"(b),"=r"(c),"=r"(d) : "l"(p));      // 16 B only
    if (V == 4) { asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(a) : "l"(p)); b = c = d = 0; }                          // 4 B only
    if (V == 5) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a),"=r"(b),"=r"(c),"=r"(d) : "l"(p));
    if (V == 6) asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a),"=r"(b),"=r"(c),"=r"(d) : "l"(p));
    if (V == 7) asm volatile("ld.global.cv.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a),"=r"(b),"=r"(c),"=r"(d) : "l"(p));
    if (V == 8) asm volatile("ld.global.lu.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a),"=r"(b),"=r"(c),"=r"(d) : "l"(p));
    if (V == 9) asm volatile("ld.global.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a),"=r"(b),"=r"(c),"=r"(d),"=r"(e),"=r"(f),"=r"(g),"=r"(h) : "l"(p));
    if (V == 10) asm volatile("ld.global.L1::evict_first.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a),"=r"(b),"=r"(c),"=r"(d) : "l"(p));
    if (V == 11) asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a),"=r"(b),"=r"(c),"=r"(d) : "l"(p));
    if (V == 12) asm volatile("ld.global.L2::evict_first.L2::64B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a),"=r"(b),"=r"(c),"=r"(d),"=r"(e),"=r"(f),"=r"(g),"=r"(h) : "l"(p));
    if (V == 13) asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a),"=r"(b),"=r"(c),"=r"(d) : "l"(p));
    if (V == 14) asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a),"=r"(b),"=r"(c),"=r"(d) : "l"(p));
    if (V == 15) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a),"=r"(b),"=r"(c),"=r"(d),"=r"(e),"=r"(f),"=r"(g),"=r"(h) : "l"(p));
    if (V == 16) asm volatile("ld.global.nc.L2::64B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a),"=r"(b),"=r"(c),"=r"(d),"=r"(e),"=r"(f),"=r"(g),"=r"(h) : "l"(p));
    if (V == 17) asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.L2::64B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a),"=r"(b),"=r"(c),"=r"(d),"=r"(e),"=r"(f),"=r"(g),"=r"(h) : "l"(p));
    return a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}

template <int V>
__global__ void __launch_bounds__(256) gather(const uint4* base, uint64_t n_sectors, uint64_t n_probes, uint64_t salt, uint32_t* sink) {
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x, stride = gridDim.x * (uint64_t)blockDim.x;
    uint32_t acc = 0;
    for (uint64_t i = tid; i < n_probes; i += 4 * stride) {
        uint32_t r[4];
#pragma unroll
        for (int j = 0; j < 4; j++) r[j] = probe<V>(base + 2 * __umul64hi(mix64((i + j * stride) ^ salt), n_sectors));
        acc ^= r[0] ^ r[1] ^ r[2] ^ r[3];
    }
    if (acc == 0x12345678u) *sink = acc;
}

template <int V>
void run(const char* name, const uint4* base, uint64_t n_sectors, uint64_t n_probes, uint32_t* sink, int blocks) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather<V><<<blocks, 256>>>(base, n_sectors, n_probes, 1, sink);
    cudaEventRecord(e0);
    for (int it = 0; it < 3; it++) gather<V><<<blocks, 256>>>(base, n_sectors, n_probes, it + 2, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
    printf("variant %2d %-44s %8.3f ms  %7.2f Gprobes/s  err=%s\n", V, name, ms, n_probes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv) {
    const uint64_t gb = argc > 1 ? atoll(argv[1]) : 32;
    const int gran = argc > 2 ? atoi(argv[2]) : 0;
    if (gran) printf("set L2 fetch granularity %d -> %s\n", gran, cudaGetErrorString(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran)));
    size_t lim = 0; cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity limit: %zu\n", lim);
    const uint64_t bytes = gb << 30, n_sectors = bytes / 32;
    uint4* base; uint32_t* sink;
    if (cudaMalloc(&base, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&sink, 4); cudaMemset(base, 0x5a, bytes);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8;
    const uint64_t n_probes = (uint64_t)blocks * 256 * 4 * 96;     // ~116 M
    printf("buffer %llu GB, %llu probes, %d blocks\n", (unsigned long long)gb, (unsigned long long)n_probes, blocks);
    run<0>("nc.L1::no_allocate.v8 (current)", base, n_sectors, n_probes, sink, blocks);
    run<1>("plain ld.global.v8", base, n_sectors, n_probes, sink, blocks);
    run<2>("2 x nc.v4", base, n_sectors, n_probes, sink, blocks);
    run<3>("nc.v4 (16 B)", base, n_sectors, n_probes, sink, blocks);
    run<4>("nc.u32 (4 B)", base, n_sectors, n_probes, sink, blocks);
    run<5>("cg.v4", base, n_sectors, n_probes, sink, blocks);
    run<6>("cs.v4", base, n_sectors, n_probes, sink, blocks);
    run<7>("cv.v4", base, n_sectors, n_probes, sink, blocks);
    run<8>("lu.v4", base, n_sectors, n_probes, sink, blocks);
    run<9>("L1::no_allocate.L2::evict_first.v8", base, n_sectors, n_probes, sink, blocks);
    run<10>("L1::evict_first.L2::evict_first.v4", base, n_sectors, n_probes, sink, blocks);
    run<11>("nc.L2::64B.v4", base, n_sectors, n_probes, sink, blocks);
    run<12>("L2::evict_first.L2::64B.v8", base, n_sectors, n_probes, sink, blocks);
    run<13>("ld.relaxed.gpu.v4", base, n_sectors, n_probes, sink, blocks);
    run<14>("ld.volatile.v4", base, n_sectors, n_probes, sink, blocks);
    run<15>("nc.L1::no_allocate.L2::64B.v8", base, n_sectors, n_probes, sink, blocks);
    run<16>("nc.L2::64B.v8", base, n_sectors, n_probes, sink, blocks);
    run<17>("nc.L1::no_alloc.L2::evict_first.L2::64B.v8", base, n_sectors, n_probes, sink, blocks);
    return 0;
}

// Streaming-load microbenchmark for sm_100a: 148 persistent CTAs, each pulling tiles into a shared-memory ring with bulk copies
// (cp.async.bulk + mbarrier), nothing else -- the load side of tcn2_mac_kernel in isolation.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/stream_probe tools/probe/stream_probe.cu
// Patterns: (a) the conv1 pattern: CTA b walks slot b, b+148, ...; per slot 25 chunks x 7 planes, tile (slot, plane, chunk) at
// ((slot*7 + plane)*ncols + chunk*128) * 32 bytes;  (b) the same bytes read as one contiguous run per CTA.
// Prints GB/s for ring depths 8 and 32 and tile sizes 4 KB / 16 KB.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
#ifdef USE_TEST_WAIT
#define WAIT_OP "mbarrier.test_wait.parity.shared::cta.b64"
#else
#define WAIT_OP "mbarrier.try_wait.parity.shared::cta.b64"
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\t" WAIT_OP " p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// tiles: total tile count of this CTA; addr(i) gives the source of its i-th tile
template <int PATTERN>
__global__ void __launch_bounds__(256, 1) stream(const uint8_t *buf, int tile_bytes, int stages, long nslots, int ncols, int row_bytes, uint32_t *sink, int P, int same_warp) {
    extern __shared__ __align__(1024) uint8_t sm[];
    const uint32_t base = smem_u32(sm);
    const uint32_t bars = base + stages * tile_bytes;
    const int chunks = ncols / 128, planes = 7;
    const long my_slots = (nslots - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const long tiles = my_slots * chunks * planes;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; s++) mbar_init(bars + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (same_warp ? (int)threadIdx.x < P : ((threadIdx.x & 31) == 0 && (int)(threadIdx.x >> 5) < P)) {
        const int t = same_warp ? threadIdx.x : threadIdx.x >> 5;   // producer t owns tiles i with i % P == t (and the stages with the same residue)
        auto addr = [&](long i) -> const uint8_t * {
            if (PATTERN == 1) return buf + ((long)blockIdx.x * tiles + i) * tile_bytes;   // contiguous per CTA
            const long sl = blockIdx.x + (i / (chunks * planes)) * gridDim.x;
            const int r = (int)(i % (chunks * planes)), ch = r / planes, pl = r % planes;
            return buf + ((sl * planes + pl) * ncols + (long)ch * 128) * row_bytes;
        };
        uint32_t acc = 0;
        for (long i = t; i < stages && i < tiles; i += P) { mbar_expect_tx(bars + 8 * (int)i, tile_bytes); bulk_load(base + (int)i * tile_bytes, addr(i), tile_bytes, bars + 8 * (int)i); }
        for (long i = t; i < tiles; i += P) {
            const int s = (int)(i % stages);
            mbar_wait(bars + 8 * s, (uint32_t)((i / stages) & 1));
            acc += sm[s * tile_bytes + (i & 63)];
            if (i + stages < tiles) { mbar_expect_tx(bars + 8 * s, tile_bytes); bulk_load(base + s * tile_bytes, addr(i + stages), tile_bytes, bars + 8 * s); }
        }
        if (acc == 0x7fffffff) sink[0] = acc;
    }
}

template <int PATTERN>
void run(const char *name, const uint8_t *buf, int tile_bytes, int stages, long nslots, int ncols, int row_bytes, uint32_t *sink, int P = 1, int same_warp = 0) {
    const size_t smem = (size_t)stages * tile_bytes + 8 * stages + 64;
    cudaFuncSetAttribute(stream<PATTERN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    stream<PATTERN><<<148, 256, smem>>>(buf, tile_bytes, stages, nslots, ncols, row_bytes, sink, P, same_warp);
    cudaEventRecord(a);
    stream<PATTERN><<<148, 256, smem>>>(buf, tile_bytes, stages, nslots, ncols, row_bytes, sink, P, same_warp);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double bytes = (double)nslots * 7 * ncols * row_bytes;
    printf("%-34s producers %d%s  tile %5d B  ring %2d  %7.1f GB/s  (%.2f ms, %s)\n", name, P, same_warp ? " (lanes of one warp)" : "", tile_bytes, stages, bytes / ms / 1e6, ms, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const long nslots = 8192;
    const int ncols = 3072;           // 24 chunks of 128 columns
    const size_t bytes = (size_t)nslots * 7 * ncols * 128 + (1 << 20);
    uint8_t *buf; uint32_t *sink;
    cudaMalloc(&buf, bytes); cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, bytes);
    for (int stages : {8, 32}) {
        run<0>("conv1 pattern (32 B rows)", buf, 4096, stages, nslots, ncols, 32, sink);
        run<1>("contiguous per CTA", buf, 4096, stages, nslots, ncols, 32, sink);
    }
    for (int P : {2, 4, 8}) run<0>("conv1 pattern (32 B rows)", buf, 4096, 8 * P > 32 ? 32 : 8 * P, nslots, ncols, 32, sink, P);
    for (int P : {2, 4, 8}) run<0>("conv1 pattern (32 B rows)", buf, 4096, 8 * P > 32 ? 32 : 8 * P, nslots, ncols, 32, sink, P, 1);
    for (int P : {2, 4}) run<0>("conv2 pattern (128 B rows)", buf, 16384, 12, nslots / 4, ncols, 128, sink, P);
    for (int stages : {6, 12}) {
        run<0>("conv2 pattern (128 B rows)", buf, 16384, stages, nslots / 4, ncols, 128, sink);
        run<1>("contiguous per CTA", buf, 16384, stages, nslots / 4, ncols, 128, sink);
    }
    return 0;
}

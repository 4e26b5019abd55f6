// store_ceiling.cu -- what a pure store stream can reach on this GPU, in the step kernel's own
// shape (one CTA per 32-env tile = 230 400 contiguous bytes, 128 threads, 16 B per lane per
// instruction), against cudaMemset and against other store flavours / CTA shapes.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o store_ceiling store_ceiling.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

enum { ST_CS = 0, ST_DEFAULT = 1, ST_WT = 2, ST_CG = 3, ST_V8 = 4, ST_NOALLOC = 5 };

template <int MODE>
__device__ __forceinline__ void store16(uint4 *p, uint4 v)
{
    if (MODE == ST_CS) __stcs(p, v);
    else if (MODE == ST_WT) __stwt(p, v);
    else if (MODE == ST_CG) __stcg(p, v);
    else if (MODE == ST_NOALLOC)
        asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else *p = v;
}

// one CTA per tile of `tile_chunks` 16-byte chunks
template <int MODE, int T>
__global__ void __launch_bounds__(T) tile_store(uint4 *out, int tile_chunks, uint32_t seed)
{
    uint4 *o = out + (size_t)blockIdx.x * tile_chunks;
    const uint4 v = make_uint4(seed, threadIdx.x, blockIdx.x, 0x3F800000u);
#pragma unroll 4
    for (int g = threadIdx.x; g < tile_chunks; g += T) store16<MODE>(o + g, v);
}

// 32-byte stores (sm_100 256-bit vector store)
template <int T>
__global__ void __launch_bounds__(T) tile_store_v8(uint4 *out, int tile_chunks, uint32_t seed)
{
    uint4 *o = out + (size_t)blockIdx.x * tile_chunks;
    const uint32_t a = seed, b = threadIdx.x, c = blockIdx.x, d = 0x3F800000u;
#pragma unroll 4
    for (int g = threadIdx.x * 2; g < tile_chunks; g += 2 * T)
        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(o + g), "r"(a), "r"(b), "r"(c), "r"(d),
                     "r"(a), "r"(b), "r"(c), "r"(d)
                     : "memory");
}

// persistent grid-stride over tiles
template <int MODE, int T>
__global__ void __launch_bounds__(T) tile_store_persistent(uint4 *out, int tile_chunks, int ntiles, uint32_t seed)
{
    const uint4 v = make_uint4(seed, threadIdx.x, blockIdx.x, 0x3F800000u);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        uint4 *o = out + (size_t)tile * tile_chunks;
#pragma unroll 4
        for (int g = threadIdx.x; g < tile_chunks; g += T) store16<MODE>(o + g, v);
    }
}

// a "logic phase" stand-in: spin for `spin` clocks on the first warp before storing (all wait at a barrier)
template <int MODE, int T>
__global__ void __launch_bounds__(T) tile_store_with_gap(uint4 *out, int tile_chunks, int spin, uint32_t seed)
{
    if (threadIdx.x < 32) {
        const long long t0 = clock64();
        while (clock64() - t0 < spin) {}
    }
    __syncthreads();
    uint4 *o = out + (size_t)blockIdx.x * tile_chunks;
    const uint4 v = make_uint4(seed, threadIdx.x, blockIdx.x, 0x3F800000u);
#pragma unroll 4
    for (int g = threadIdx.x; g < tile_chunks; g += T) store16<MODE>(o + g, v);
}

template <typename F>
static float time_ms(F f, int reps = 20)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(e));
    return ms / reps;
}

int main()
{
    const int ntiles = 32768;             // 1 048 576 envs / 32
    const int tile_chunks = 32 * 450;     // 32 envs x 7200 B / 16
    const size_t bytes = (size_t)ntiles * tile_chunks * 16;
    uint4 *buf;
    if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    auto report = [&](const char *name, float ms) { printf("%-58s %8.3f ms  %8.1f GB/s\n", name, ms, bytes / ms / 1e6); };

    report("cudaMemsetAsync", time_ms([&] { cudaMemsetAsync(buf, 0, bytes); }));
    report("tile/CTA 128 thr st.cs (the step kernel's stream)", time_ms([&] { tile_store<ST_CS, 128><<<ntiles, 128>>>(buf, tile_chunks, 1); }));
    report("tile/CTA 128 thr st (default)", time_ms([&] { tile_store<ST_DEFAULT, 128><<<ntiles, 128>>>(buf, tile_chunks, 1); }));
    report("tile/CTA 128 thr st.wt", time_ms([&] { tile_store<ST_WT, 128><<<ntiles, 128>>>(buf, tile_chunks, 1); }));
    report("tile/CTA 128 thr st.cg", time_ms([&] { tile_store<ST_CG, 128><<<ntiles, 128>>>(buf, tile_chunks, 1); }));
    report("tile/CTA 128 thr st.L1::no_allocate", time_ms([&] { tile_store<ST_NOALLOC, 128><<<ntiles, 128>>>(buf, tile_chunks, 1); }));
    report("tile/CTA 256 thr st.cs", time_ms([&] { tile_store<ST_CS, 256><<<ntiles, 256>>>(buf, tile_chunks, 1); }));
    report("tile/CTA 256 thr st", time_ms([&] { tile_store<ST_DEFAULT, 256><<<ntiles, 256>>>(buf, tile_chunks, 1); }));
    report("tile/CTA 512 thr st", time_ms([&] { tile_store<ST_DEFAULT, 512><<<ntiles, 512>>>(buf, tile_chunks, 1); }));
    report("tile/CTA 128 thr 32-byte stores (st.v8.b32)", time_ms([&] { tile_store_v8<128><<<ntiles, 128>>>(buf, tile_chunks, 1); }));
    report("tile/CTA 256 thr 32-byte stores (st.v8.b32)", time_ms([&] { tile_store_v8<256><<<ntiles, 256>>>(buf, tile_chunks, 1); }));
    report("half tiles (16 envs) 128 thr st.cs", time_ms([&] { tile_store<ST_CS, 128><<<ntiles * 2, 128>>>(buf, tile_chunks / 2, 1); }));
    report("double tiles (64 envs) 128 thr st.cs", time_ms([&] { tile_store<ST_CS, 128><<<ntiles / 2, 128>>>(buf, tile_chunks * 2, 1); }));
    report("double tiles (64 envs) 256 thr st", time_ms([&] { tile_store<ST_DEFAULT, 256><<<ntiles / 2, 256>>>(buf, tile_chunks * 2, 1); }));
    for (int mult : {2, 4, 8, 16}) {
        char nm[96];
        snprintf(nm, sizeof nm, "persistent %d CTAs/SM x 128 thr st.cs", mult);
        report(nm, time_ms([&] { tile_store_persistent<ST_CS, 128><<<148 * mult, 128>>>(buf, tile_chunks, ntiles, 1); }));
        snprintf(nm, sizeof nm, "persistent %d CTAs/SM x 128 thr st", mult);
        report(nm, time_ms([&] { tile_store_persistent<ST_DEFAULT, 128><<<148 * mult, 128>>>(buf, tile_chunks, ntiles, 1); }));
    }
    for (int spin : {2000, 5000, 10000, 20000}) {
        char nm[96];
        snprintf(nm, sizeof nm, "tile/CTA 128 thr st.cs after a %d-clock single-warp phase", spin);
        report(nm, time_ms([&] { tile_store_with_gap<ST_CS, 128><<<ntiles, 128>>>(buf, tile_chunks, spin, 1); }));
        snprintf(nm, sizeof nm, "tile/CTA 128 thr st after a %d-clock single-warp phase", spin);
        report(nm, time_ms([&] { tile_store_with_gap<ST_DEFAULT, 128><<<ntiles, 128>>>(buf, tile_chunks, spin, 1); }));
    }
    cudaFree(buf);
    return 0;
}

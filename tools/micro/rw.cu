// HBM read/write mix microbenchmark (tools only; not product code).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/micro/rw tools/micro/rw.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int MODE> __device__ __forceinline__ void st16(float4* p, float4 v) {
    if (MODE == 0) *p = v;
    else if (MODE == 1) __stcs(p, v);
    else if (MODE == 2) __stcg(p, v);
    else __stwt(p, v);
}

// write-only: n16 float4 elements
template <int MODE, int U>
__global__ void __launch_bounds__(256) fill_k(float4* __restrict__ out, size_t n16) {
    const size_t stride = (size_t)gridDim.x * 256 * U;
    for (size_t i = (size_t)blockIdx.x * 256 * U + threadIdx.x; i < n16; i += stride) {
#pragma unroll
        for (int u = 0; u < U; ++u) if (i + u * 256 < n16) st16<MODE>(out + i + u * 256, make_float4((float)i, 1.f, 2.f, 3.f));
    }
}
// read-only
template <int U>
__global__ void __launch_bounds__(256) read_k(const int4* __restrict__ in, size_t n16, int* sink) {
    const size_t stride = (size_t)gridDim.x * 256 * U;
    int acc = 0;
    for (size_t i = (size_t)blockIdx.x * 256 * U + threadIdx.x; i < n16; i += stride) {
        int4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = (i + u * 256 < n16) ? __ldg(in + i + u * 256) : make_int4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x12345678) *sink = acc;
}
// expand: read n16 x 16 B, write EXP x 16 B per 16 B read (EXP = 1 copy, 2 = s16->f32 mono, 4 = s16 -> f32 ... )
template <int MODE, int U, int EXP>
__global__ void __launch_bounds__(256) expand_k(const int4* __restrict__ in, float4* __restrict__ out, size_t n16) {
    const size_t stride = (size_t)gridDim.x * 256 * U;
    for (size_t i = (size_t)blockIdx.x * 256 * U + threadIdx.x; i < n16; i += stride) {
        int4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = (i + u * 256 < n16) ? __ldg(in + i + u * 256) : make_int4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (i + u * 256 >= n16) continue;
            const size_t o = (i + u * 256) * EXP;
#pragma unroll
            for (int e = 0; e < EXP; ++e)
                st16<MODE>(out + o + e, make_float4((float)(v[u].x >> e), (float)v[u].y, (float)v[u].z, (float)v[u].w));
        }
    }
}
// expand through shared memory with contiguous warp stores: a CTA tile = 256*U int4 in, EXP x that out
template <int MODE, int U, int EXP>
__global__ void __launch_bounds__(256) expand_smem_k(const int4* __restrict__ in, float4* __restrict__ out, size_t n16) {
    __shared__ int4 s[256 * U];
    const size_t ntile = n16 / (256 * U);
    for (size_t t = blockIdx.x; t < ntile; t += gridDim.x) {
        const int4* ip = in + t * 256 * U;
#pragma unroll
        for (int u = 0; u < U; ++u) s[u * 256 + threadIdx.x] = __ldg(ip + u * 256 + threadIdx.x);
        __syncthreads();
        float4* op = out + t * 256 * U * EXP;
#pragma unroll
        for (int k = 0; k < U * EXP; ++k) {
            const int o = k * 256 + threadIdx.x;          // output float4 index in the tile
            const int4 v = s[o / EXP];
            st16<MODE>(op + o, make_float4((float)(v.x >> (o % EXP)), (float)v.y, (float)v.z, (float)v.w));
        }
        __syncthreads();
    }
}
// phased: chip-wide read phase into shared memory, grid barrier, write phase (cooperative launch)
template <int MODE, int EXP, int SM16>   // SM16 int4 per CTA per phase
__global__ void __launch_bounds__(512) phased_k(const int4* __restrict__ in, float4* __restrict__ out, size_t n16) {
    extern __shared__ int4 sm[];
    cg::grid_group g = cg::this_grid();
    const size_t per_round = (size_t)gridDim.x * SM16;
    for (size_t r0 = 0; r0 + per_round <= n16; r0 += per_round) {
        const int4* ip = in + r0 + (size_t)blockIdx.x * SM16;
        for (int i = threadIdx.x; i < SM16; i += 512) sm[i] = __ldg(ip + i);
        g.sync();
        float4* op = out + (r0 + (size_t)blockIdx.x * SM16) * EXP;
        for (int o = threadIdx.x; o < SM16 * EXP; o += 512) {
            const int4 v = sm[o / EXP];
            st16<MODE>(op + o, make_float4((float)(v.x >> (o % EXP)), (float)v.y, (float)v.z, (float)v.w));
        }
        g.sync();
    }
}
// TMA bulk store variant of expand_smem: convert into shared memory, one thread issues cp.async.bulk stores
template <int U, int EXP>
__global__ void __launch_bounds__(256) expand_bulk_k(const int4* __restrict__ in, float4* __restrict__ out, size_t n16) {
    __shared__ __align__(128) float4 so[2][256 * U * EXP];
    const size_t ntile = n16 / (256 * U);
    int buf = 0;
    for (size_t t = blockIdx.x; t < ntile; t += gridDim.x, buf ^= 1) {
        const int4* ip = in + t * 256 * U;
        int4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldg(ip + u * 256 + threadIdx.x);
        // the bulk store issued two tiles ago from this buffer must have finished reading shared memory
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncthreads();
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int e = 0; e < EXP; ++e)
                so[buf][(u * 256 + threadIdx.x) * EXP + e] = make_float4((float)(v[u].x >> e), (float)v[u].y, (float)v[u].z, (float)v[u].w);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned saddr = (unsigned)__cvta_generic_to_shared(&so[buf][0]);
            float4* op = out + t * 256 * U * EXP;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(op), "r"(saddr), "r"(256 * U * EXP * 16) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <class F> float time_ms(F&& f, int iters = 5) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < iters; ++i) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    const size_t in_bytes = 1ull << 30;                 // 1 GiB read
    const size_t n16 = in_bytes / 16;
    const size_t out_bytes = in_bytes * 4;
    int4* in; float4* out; int* sink;
    CK(cudaMalloc(&in, in_bytes)); CK(cudaMalloc(&out, out_bytes)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(in, 1, in_bytes)); CK(cudaMemset(out, 0, out_bytes));
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto rep = [&](const char* name, double bytes, float ms) { printf("%-44s %8.3f ms  %7.1f GB/s\n", name, ms, bytes / ms / 1e6); fflush(stdout); };
    const size_t o16 = out_bytes / 16;
    for (int g : {8, 16}) {
        const int grid = sms * g;
        printf("-- grid = %d x SMs\n", g);
        rep("fill st      (4 GiB)", out_bytes, time_ms([&] { fill_k<0, 4><<<grid, 256>>>(out, o16); }));
        rep("fill stcs", out_bytes, time_ms([&] { fill_k<1, 4><<<grid, 256>>>(out, o16); }));
        rep("fill stcg", out_bytes, time_ms([&] { fill_k<2, 4><<<grid, 256>>>(out, o16); }));
        rep("fill stwt", out_bytes, time_ms([&] { fill_k<3, 4><<<grid, 256>>>(out, o16); }));
        rep("read ldg U4  (1 GiB)", in_bytes, time_ms([&] { read_k<4><<<grid, 256>>>(in, n16, sink); }));
        rep("read ldg U8", in_bytes, time_ms([&] { read_k<8><<<grid, 256>>>(in, n16, sink); }));
        rep("copy 1:1 st", in_bytes * 2.0, time_ms([&] { expand_k<0, 4, 1><<<grid, 256>>>(in, out, n16); }));
        rep("copy 1:1 stcs", in_bytes * 2.0, time_ms([&] { expand_k<1, 4, 1><<<grid, 256>>>(in, out, n16); }));
        rep("expand 1:2 st (strided lanes)", in_bytes * 3.0, time_ms([&] { expand_k<0, 4, 2><<<grid, 256>>>(in, out, n16); }));
        rep("expand 1:2 stcs (strided lanes)", in_bytes * 3.0, time_ms([&] { expand_k<1, 4, 2><<<grid, 256>>>(in, out, n16); }));
        rep("expand 1:2 smem st (contiguous)", in_bytes * 3.0, time_ms([&] { expand_smem_k<0, 4, 2><<<grid, 256>>>(in, out, n16); }));
        rep("expand 1:2 smem stcs", in_bytes * 3.0, time_ms([&] { expand_smem_k<1, 4, 2><<<grid, 256>>>(in, out, n16); }));
        rep("expand 1:4 smem st", in_bytes * 5.0, time_ms([&] { expand_smem_k<0, 4, 4><<<grid, 256>>>(in, out, n16); }));
        rep("expand 1:2 bulk store", in_bytes * 3.0, time_ms([&] { expand_bulk_k<2, 2><<<grid, 256>>>(in, out, n16); }));
        rep("expand 1:4 bulk store", in_bytes * 5.0, time_ms([&] { expand_bulk_k<1, 4><<<grid, 256>>>(in, out, n16); }));
    }
    {   // phased (cooperative): one 512-thread CTA per SM, 64 KiB / 96 KiB of input per CTA per round
        auto run = [&](auto kern, int sm16, const char* name, int exp) {
            const size_t smem = (size_t)sm16 * 16;
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const int4* ip = in; float4* op = out; size_t n = n16;
            void* args[] = {&ip, &op, &n};
            rep(name, in_bytes * (1.0 + exp), time_ms([&] { CK(cudaLaunchCooperativeKernel((void*)kern, dim3(sms), dim3(512), args, smem, 0)); }));
        };
        run(phased_k<0, 2, 4096>, 4096, "phased 1:2 st, 64 KiB rounds", 2);
        run(phased_k<0, 2, 12288>, 12288, "phased 1:2 st, 192 KiB rounds", 2);
        run(phased_k<1, 2, 12288>, 12288, "phased 1:2 stcs, 192 KiB rounds", 2);
    }
    return 0;
}

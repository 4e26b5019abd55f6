// policy_kernels.cu -- fused LayerNorm(+residual)+ReLU for the policy's channels-last bf16 path.
//
// The policy CNN (inversus_rl/policies.py:27-45, :94-100) applies LayerNorm over the whole
// [C,H,W] feature map after every convolution, then ReLU (and a residual add before the fourth).
// The contractions stay in cuDNN/cuBLAS; this file fuses the memory-bound glue between them:
//
//   forward   y = relu( LN(x [+ res]) * gamma + beta )          one read of x (and res), one write of y
//   backward  dx = LN'(relu'(dy)) ; dgamma, dbeta               one read of dy, x (and res), one write of dx
//
// Layout: a sample is D = H*W*C contiguous bf16 values in HWC order (the memory of a
// channels-last tensor); gamma/beta are [D] bf16 in the same order. Statistics and all arithmetic
// are fp32. Both kernels are persistent: a CTA keeps its slice of gamma/beta in registers and
// walks over samples, so the affine parameters are read once per CTA instead of once per sample.
// The backward keeps per-CTA dgamma/dbeta accumulators in shared memory (2*D floats, up to 150 KB)
// and writes them as partials that a second small kernel reduces -- deterministic, no atomics.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/inversus_b200.h"

namespace {

constexpr int kLnThreads = 512;

__device__ __forceinline__ void unpack8(const uint4 &v, float (&f)[8])
{
    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8])
{
    uint4 v;
    __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return v;
}

// sum of (a, b) over the CTA; result broadcast to every thread
__device__ __forceinline__ float2 block_sum2(float a, float b, float2 *scratch)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads(); // scratch may still be read from the previous call
    if (lane == 0) scratch[warp] = make_float2(a, b);
    __syncthreads();
    float2 t = lane < (kLnThreads / 32) ? scratch[lane] : make_float2(0.f, 0.f);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        t.x += __shfl_xor_sync(0xffffffffu, t.x, o);
        t.y += __shfl_xor_sync(0xffffffffu, t.y, o);
    }
    return make_float2(__shfl_sync(0xffffffffu, t.x, 0), __shfl_sync(0xffffffffu, t.y, 0));
}

// Two CTAs per SM (<= 64 registers at 512 threads): while one CTA sits in its block reduction the
// other streams. gamma/beta are re-read per sample from L1/L2 (caching them in registers would not
// fit the 64-register budget together with the packed inputs).
template <int MAXV, bool HAS_RES>
__global__ void __launch_bounds__(kLnThreads, 2)
ln_relu_fwd_kernel(const uint4 *__restrict__ x, const uint4 *__restrict__ res, const uint4 *__restrict__ cbias,
                   int cvecs, const uint4 *__restrict__ gamma, const uint4 *__restrict__ beta, int64_t B, int nvec,
                   float eps, uint4 *__restrict__ y, float *__restrict__ mean_out, float *__restrict__ rstd_out)
{
    __shared__ float2 scratch[kLnThreads / 32];
    const int tid = threadIdx.x;
    // per-channel convolution bias, folded in here so cuDNN runs bias-free: the element (i*8 + j) of
    // an HWC sample belongs to channel ((i mod C/8)*8 + j), and since kLnThreads is a multiple of
    // C/8 every vector of a thread sees the same 8 channels -> one register vector per thread
    const uint4 cbv = cbias ? cbias[tid % cvecs] : make_uint4(0u, 0u, 0u, 0u); // kept packed: 4 registers
    constexpr bool kCacheGB = MAXV <= 1;
    uint4 g[kCacheGB ? MAXV : 1], bt[kCacheGB ? MAXV : 1];
    if (kCacheGB) {
#pragma unroll
        for (int k = 0; k < MAXV; ++k) {
            const int i = k * kLnThreads + tid;
            if (i < nvec) { g[k] = gamma[i]; bt[k] = beta[i]; }
        }
    }
    const float inv_d = 1.0f / (float)(nvec * 8);
    for (int64_t s = blockIdx.x; s < B; s += gridDim.x) {
        const uint4 *xs = x + s * nvec;
        uint4 xv[MAXV], rv[HAS_RES ? MAXV : 1]; // inputs stay packed (bf16) between the two passes
#pragma unroll
        for (int k = 0; k < MAXV; ++k) {
            const int i = k * kLnThreads + tid;
            if (i < nvec) {
                xv[k] = xs[i];
                if (HAS_RES) rv[k] = res[s * nvec + i];
            }
        }
        float sum = 0.f, sq = 0.f;
#pragma unroll
        for (int k = 0; k < MAXV; ++k) {
            const int i = k * kLnThreads + tid;
            if (i < nvec) {
                float z[8];
                unpack8(xv[k], z);
                {
                    float cb[8];
                    unpack8(cbv, cb);
#pragma unroll
                    for (int j = 0; j < 8; ++j) z[j] += cb[j];
                }
                if (HAS_RES) {
                    float r[8];
                    unpack8(rv[k], r);
#pragma unroll
                    for (int j = 0; j < 8; ++j) z[j] += r[j];
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) { sum += z[j]; sq += z[j] * z[j]; }
            }
        }
        const float2 t = block_sum2(sum, sq, scratch);
        const float mean = t.x * inv_d;
        const float var = fmaxf(t.y * inv_d - mean * mean, 0.f);
        const float rstd = rsqrtf(var + eps);
        if (tid == 0) { mean_out[s] = mean; rstd_out[s] = rstd; }
#pragma unroll
        for (int k = 0; k < MAXV; ++k) {
            const int i = k * kLnThreads + tid;
            if (i < nvec) {
                float z[8], gf[8], bf[8], o[8];
                unpack8(xv[k], z);
                {
                    float cb[8];
                    unpack8(cbv, cb);
#pragma unroll
                    for (int j = 0; j < 8; ++j) z[j] += cb[j];
                }
                if (HAS_RES) {
                    float r[8];
                    unpack8(rv[k], r);
#pragma unroll
                    for (int j = 0; j < 8; ++j) z[j] += r[j];
                }
                unpack8(kCacheGB ? g[k] : gamma[i], gf);
                unpack8(kCacheGB ? bt[k] : beta[i], bf);
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = fmaxf((z[j] - mean) * rstd * gf[j] + bf[j], 0.f);
                y[s * nvec + i] = pack8(o);
            }
        }
    }
}

// dynamic smem: float acc[2][nvec*8]  (dgamma, dbeta of this CTA)
template <int MAXV, bool HAS_RES>
__global__ void __launch_bounds__(kLnThreads)
ln_relu_bwd_kernel(const uint4 *__restrict__ dy, const uint4 *__restrict__ x, const uint4 *__restrict__ res,
                   const uint4 *__restrict__ cbias, int cvecs, const uint4 *__restrict__ gamma,
                   const uint4 *__restrict__ beta, const float *__restrict__ mean_in,
                   const float *__restrict__ rstd_in, int64_t B, int nvec, uint4 *__restrict__ dx,
                   float *__restrict__ partials)
{
    extern __shared__ float acc[];
    __shared__ float2 scratch[kLnThreads / 32];
    const int tid = threadIdx.x;
    const int D = nvec * 8;
    float cb[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (cbias) unpack8(cbias[tid % cvecs], cb);
    float db[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}; // d(conv bias) of this thread's 8 channels
    for (int i = tid; i < 2 * D; i += kLnThreads) acc[i] = 0.f;
    __syncthreads();
    // gamma/beta live in registers across samples when the slice is small; for the two 19200-wide
    // layers (MAXV >= 4) that would spill, so they are re-read per sample (they stay L1/L2-resident)
    constexpr bool kCacheGB = MAXV <= 3;
    uint4 g[kCacheGB ? MAXV : 1], bt[kCacheGB ? MAXV : 1];
    if (kCacheGB) {
#pragma unroll
        for (int k = 0; k < MAXV; ++k) {
            const int i = k * kLnThreads + tid;
            if (i < nvec) { g[k] = gamma[i]; bt[k] = beta[i]; }
        }
    }
    const float inv_d = 1.0f / (float)D;
    // from here on each thread only touches its own accumulator slots: no barrier needed around acc
    for (int64_t s = blockIdx.x; s < B; s += gridDim.x) {
        const float mean = mean_in[s], rstd = rstd_in[s];
        // issue every global load of this sample before touching the data (memory-level
        // parallelism); the packed inputs stay in registers for both passes
        uint4 xv[MAXV], dv[MAXV], rv[HAS_RES ? MAXV : 1];
#pragma unroll
        for (int k = 0; k < MAXV; ++k) {
            const int i = k * kLnThreads + tid;
            if (i < nvec) {
                xv[k] = x[s * nvec + i];
                dv[k] = dy[s * nvec + i];
                if (HAS_RES) rv[k] = res[s * nvec + i];
            }
        }
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < MAXV; ++k) {
            const int i = k * kLnThreads + tid;
            if (i < nvec) {
                float z[8], d[8], gf[8], bf[8];
                unpack8(xv[k], z);
#pragma unroll
                for (int j = 0; j < 8; ++j) z[j] += cb[j];
                if (HAS_RES) {
                    float r[8];
                    unpack8(rv[k], r);
#pragma unroll
                    for (int j = 0; j < 8; ++j) z[j] += r[j];
                }
                unpack8(dv[k], d);
                unpack8(kCacheGB ? g[k] : gamma[i], gf);
                unpack8(kCacheGB ? bt[k] : beta[i], bf);
                // accumulators are element-major (acc[j*nvec + i]) so that the 32 lanes of a warp
                // hit 32 different banks; i*8 + j would be an 8-way conflict on every access
                float *ag = acc + i, *ab = acc + D + i;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float h = (z[j] - mean) * rstd;
                    const float gy = (h * gf[j] + bf[j] > 0.f) ? d[j] : 0.f; // relu'
                    ag[j * nvec] += gy * h;
                    ab[j * nvec] += gy;
                    const float w = gy * gf[j];
                    s1 += w;
                    s2 += w * h;
                }
            }
        }
        const float2 t = block_sum2(s1, s2, scratch);
        const float m1 = t.x * inv_d, m2 = t.y * inv_d;
#pragma unroll
        for (int k = 0; k < MAXV; ++k) {
            const int i = k * kLnThreads + tid;
            if (i < nvec) {
                float z[8], d[8], gf[8], bf[8], o[8];
                unpack8(xv[k], z);
#pragma unroll
                for (int j = 0; j < 8; ++j) z[j] += cb[j];
                if (HAS_RES) {
                    float r[8];
                    unpack8(rv[k], r);
#pragma unroll
                    for (int j = 0; j < 8; ++j) z[j] += r[j];
                }
                unpack8(dv[k], d);
                unpack8(kCacheGB ? g[k] : gamma[i], gf);
                unpack8(kCacheGB ? bt[k] : beta[i], bf);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float h = (z[j] - mean) * rstd;
                    const float w = (h * gf[j] + bf[j] > 0.f) ? d[j] * gf[j] : 0.f;
                    o[j] = rstd * (w - m1 - h * m2);
                    db[j] += o[j];
                }
                dx[s * nvec + i] = pack8(o);
            }
        }
    }
    const int C = cvecs * 8;
    float *out = partials + (size_t)blockIdx.x * (2 * D + C);
    for (int k = 0; k < MAXV; ++k) {
        const int i = k * kLnThreads + tid;
        if (i < nvec) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                out[i * 8 + j] = acc[j * nvec + i];
                out[D + i * 8 + j] = acc[D + j * nvec + i];
            }
        }
    }
    // d(conv bias): every thread holds sums for channels (tid % cvecs)*8 .. +7; fold the
    // kLnThreads / cvecs threads of each channel group through shared memory (acc is free now)
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[tid * 8 + j] = db[j];
    __syncthreads();
    if (tid < C) {
        const int grp = tid >> 3, j = tid & 7;
        float sum = 0.f;
        for (int t = grp; t < kLnThreads; t += cvecs) sum += acc[t * 8 + j];
        out[2 * D + tid] = sum;
    }
}

__global__ void reduce_partials_kernel(const float *__restrict__ partials, int nparts, int width,
                                       float *__restrict__ dgamma, float *__restrict__ dbeta,
                                       float *__restrict__ dcbias, int D)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= width) return;
    float s = 0.f;
    for (int p = 0; p < nparts; ++p) s += partials[(size_t)p * width + i];
    if (i < D) dgamma[i] = s;
    else if (i < 2 * D) dbeta[i - D] = s;
    else if (dcbias) dcbias[i - 2 * D] = s;
}

// Per-row transpose with dtype conversion: dst[r][b*A + a] = src[r][a*B + b] for a < A, b < B.
// Used to bring the 19200 trunk columns of the fused head weight from the checkpoint's CHW order
// (fp32 master) to the HWC order of the channels-last activations (bf16 working copy), and to
// carry the gradient back. 32x32 shared-memory tiles, coalesced on both sides.
template <typename TS, typename TD>
__global__ void transpose_cast_kernel(const TS *__restrict__ src, int64_t ld_src, TD *__restrict__ dst,
                                      int64_t ld_dst, int A, int B)
{
    __shared__ float tile[32][33];
    const int64_t r = blockIdx.z;
    const int a0 = blockIdx.y * 32, b0 = blockIdx.x * 32;
    const TS *s = src + r * ld_src;
    TD *d = dst + r * ld_dst;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int a = a0 + i, b = b0 + threadIdx.x;
        if (a < A && b < B) tile[i][threadIdx.x] = (float)s[(int64_t)a * B + b];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int b = b0 + i, a = a0 + threadIdx.x;
        if (a < A && b < B) d[(int64_t)b * A + a] = (TD)tile[threadIdx.x][i];
    }
}

int sm_count_cached()
{
    static int n[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 148;
    if (n[dev] == 0) {
        cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
        if (n[dev] <= 0) n[dev] = 148;
    }
    return n[dev];
}

} // namespace

extern "C" {

int inv_transpose_cast(const void *src, int32_t src_is_f32, int64_t ld_src, void *dst, int32_t dst_is_f32,
                       int64_t ld_dst, int64_t rows, int32_t A, int32_t B, void *stream)
{
    if (!src || !dst || rows < 0 || A <= 0 || B <= 0 || rows > 65535 || src_is_f32 == dst_is_f32)
        return INV_ERR_INVALID_ARG;
    if (rows == 0) return INV_OK;
    const dim3 grid((B + 31) / 32, (A + 31) / 32, (unsigned)rows), block(32, 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (src_is_f32)
        transpose_cast_kernel<float, __nv_bfloat16><<<grid, block, 0, st>>>(
            (const float *)src, ld_src, (__nv_bfloat16 *)dst, ld_dst, A, B);
    else
        transpose_cast_kernel<__nv_bfloat16, float><<<grid, block, 0, st>>>(
            (const __nv_bfloat16 *)src, ld_src, (float *)dst, ld_dst, A, B);
    return cudaGetLastError() == cudaSuccess ? INV_OK : INV_ERR_CUDA;
}

int inv_ln_relu_partials(int32_t D) { (void)D; return sm_count_cached(); }

int inv_ln_relu_fwd(const void *x, const void *res, const void *cbias, const void *gamma, const void *beta,
                    int64_t B, int32_t D, int32_t C, float eps, void *y, float *mean, float *rstd, void *stream)
{
    if (!x || !gamma || !beta || !y || !mean || !rstd || B < 0 || D <= 0 || D % 8) return INV_ERR_INVALID_ARG;
    if (C <= 0 || C % 8 || D % C || kLnThreads % (C / 8)) return INV_ERR_INVALID_ARG;
    const int cvecs = C / 8;
    if (B == 0) return INV_OK;
    const int nvec = D / 8;
    const int maxv = (nvec + kLnThreads - 1) / kLnThreads;
    if (maxv > 5) return INV_ERR_INVALID_ARG; // D <= 20480
    const int64_t cap = (int64_t)sm_count_cached() * 4;
    const unsigned grid = (unsigned)(B < cap ? B : cap);
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(MV)                                                                                              \
    if (res)                                                                                                    \
        ln_relu_fwd_kernel<MV, true><<<grid, kLnThreads, 0, st>>>((const uint4 *)x, (const uint4 *)res,         \
            (const uint4 *)cbias, cvecs, (const uint4 *)gamma, (const uint4 *)beta, B, nvec, eps, (uint4 *)y,    \
            mean, rstd);                                                                                        \
    else                                                                                                        \
        ln_relu_fwd_kernel<MV, false><<<grid, kLnThreads, 0, st>>>((const uint4 *)x, nullptr,                   \
            (const uint4 *)cbias, cvecs, (const uint4 *)gamma, (const uint4 *)beta, B, nvec, eps, (uint4 *)y,    \
            mean, rstd);
    switch (maxv) {
    case 1: LAUNCH(1) break;
    case 2: LAUNCH(2) break;
    case 3: LAUNCH(3) break;
    case 4: LAUNCH(4) break;
    default: LAUNCH(5) break;
    }
#undef LAUNCH
    return cudaGetLastError() == cudaSuccess ? INV_OK : INV_ERR_CUDA;
}

int inv_ln_relu_bwd(const void *dy, const void *x, const void *res, const void *cbias, const void *gamma,
                    const void *beta, const float *mean, const float *rstd, int64_t B, int32_t D, int32_t C,
                    void *dx, float *dgamma, float *dbeta, float *dcbias, float *partials, void *stream)
{
    if (!dy || !x || !gamma || !beta || !mean || !rstd || !dx || !dgamma || !dbeta || !partials || B <= 0 ||
        D <= 0 || D % 8)
        return INV_ERR_INVALID_ARG;
    if (C <= 0 || C % 8 || D % C || kLnThreads % (C / 8) || C > kLnThreads) return INV_ERR_INVALID_ARG;
    const int cvecs = C / 8;
    const int nvec = D / 8;
    const int maxv = (nvec + kLnThreads - 1) / kLnThreads;
    if (maxv > 5) return INV_ERR_INVALID_ARG;
    const int nsm = sm_count_cached();
    const unsigned grid = (unsigned)(B < nsm ? B : nsm); // one CTA per SM: 2*D floats of shared memory each
    size_t smem = (size_t)2 * D * sizeof(float);
    if (smem < (size_t)kLnThreads * 8 * sizeof(float)) smem = (size_t)kLnThreads * 8 * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(MV)                                                                                              \
    if (res) {                                                                                                  \
        cudaFuncSetAttribute(ln_relu_bwd_kernel<MV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        ln_relu_bwd_kernel<MV, true><<<grid, kLnThreads, smem, st>>>((const uint4 *)dy, (const uint4 *)x,       \
            (const uint4 *)res, (const uint4 *)cbias, cvecs, (const uint4 *)gamma, (const uint4 *)beta, mean,    \
            rstd, B, nvec, (uint4 *)dx, partials);                                                              \
    } else {                                                                                                    \
        cudaFuncSetAttribute(ln_relu_bwd_kernel<MV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        ln_relu_bwd_kernel<MV, false><<<grid, kLnThreads, smem, st>>>((const uint4 *)dy, (const uint4 *)x,      \
            nullptr, (const uint4 *)cbias, cvecs, (const uint4 *)gamma, (const uint4 *)beta, mean, rstd, B,      \
            nvec, (uint4 *)dx, partials);                                                                       \
    }
    switch (maxv) {
    case 1: LAUNCH(1) break;
    case 2: LAUNCH(2) break;
    case 3: LAUNCH(3) break;
    case 4: LAUNCH(4) break;
    default: LAUNCH(5) break;
    }
#undef LAUNCH
    if (cudaGetLastError() != cudaSuccess) return INV_ERR_CUDA;
    const int width = 2 * D + C;
    reduce_partials_kernel<<<(width + 255) / 256, 256, 0, st>>>(partials, (int)grid, width, dgamma, dbeta, dcbias, D);
    return cudaGetLastError() == cudaSuccess ? INV_OK : INV_ERR_CUDA;
}

} // extern "C"

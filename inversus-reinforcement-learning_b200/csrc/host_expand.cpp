// host_expand.cpp -- host half of inv_step_host's "packed over PCIe, expanded on the host" path.
//
// A float32 observation is 7200 B per env but carries 1800 bits. When the caller wants it in HOST
// memory (the reference's numpy contract, env_wrappers.py:522-528), PCIe (about 52 GB/s measured)
// caps the direct copy at about 7e6 env-steps/s. The step kernel therefore also emits each env's
// observation as the packed 1800-bit row it already holds in shared memory (256 B/env with
// padding); this file turns those rows back into the float32 layout with non-temporal AVX2 stores
// on the host threads, while the copy engine moves the remaining envs' float32 data directly.
// Pure format conversion: no game logic runs on the CPU.
#include <immintrin.h>
#include <stdint.h>

#include <algorithm>
#include <thread>
#include <vector>

namespace inv_host {

static void expand_scalar(const uint32_t *bits, float *dst, int64_t lo, int64_t hi)
{
    for (int64_t e = lo; e < hi; ++e) {
        const uint32_t *row = bits + e * 64;
        float *o = dst + e * 1800;
        for (int i = 0; i < 1800; ++i) o[i] = ((row[i >> 5] >> (i & 31)) & 1u) ? 1.0f : 0.0f;
    }
}

__attribute__((target("avx2"))) static void expand_avx2(const uint32_t *bits, float *dst, int64_t lo, int64_t hi)
{
    const __m256i sel = _mm256_setr_epi32(1, 2, 4, 8, 16, 32, 64, 128);
    const __m256 one = _mm256_set1_ps(1.0f);
    const bool aligned = (reinterpret_cast<uintptr_t>(dst) & 31u) == 0; // 1800 floats = 225 x 32 B per env
    for (int64_t e = lo; e < hi; ++e) {
        const uint8_t *row = reinterpret_cast<const uint8_t *>(bits + e * 64);
        float *o = dst + e * 1800;
        if (aligned) {
            for (int b = 0; b < 225; ++b) {
                const __m256i v = _mm256_and_si256(_mm256_set1_epi32(row[b]), sel);
                const __m256 f = _mm256_and_ps(_mm256_castsi256_ps(_mm256_cmpeq_epi32(v, sel)), one);
                _mm256_stream_ps(o + 8 * b, f);
            }
        } else {
            for (int b = 0; b < 225; ++b) {
                const __m256i v = _mm256_and_si256(_mm256_set1_epi32(row[b]), sel);
                const __m256 f = _mm256_and_ps(_mm256_castsi256_ps(_mm256_cmpeq_epi32(v, sel)), one);
                _mm256_storeu_ps(o + 8 * b, f);
            }
        }
    }
    _mm_sfence();
}

// Expand envs [lo, hi) of `bits` ([n][64] u32, little-endian bit i = observation element i) into
// `dst` ([n][1800] f32) using up to `nthreads` threads.
void expand_f32(const uint32_t *bits, float *dst, int64_t lo, int64_t hi, int nthreads)
{
    if (hi <= lo) return;
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    auto work = [&](int64_t a, int64_t b) {
        if (have_avx2) expand_avx2(bits, dst, a, b);
        else expand_scalar(bits, dst, a, b);
    };
    const int64_t n = hi - lo;
    int nt = (int)std::max<int64_t>(1, std::min<int64_t>(nthreads, n / 256));
    if (nt == 1) { work(lo, hi); return; }
    std::vector<std::thread> th;
    th.reserve(nt);
    for (int t = 0; t < nt; ++t) th.emplace_back(work, lo + n * t / nt, lo + n * (t + 1) / nt);
    for (auto &x : th) x.join();
}

int hardware_threads()
{
    unsigned n = std::thread::hardware_concurrency();
    return n ? (int)n : 1;
}

} // namespace inv_host

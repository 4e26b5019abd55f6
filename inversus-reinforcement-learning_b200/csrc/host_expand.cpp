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
#include <stdlib.h>

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

// AVX-512: envs are handled in pairs. One env is 1800 floats = 112.5 cache lines, so only every
// second env starts on a 64-byte boundary; a pair is exactly 225 lines and every store is a full
// non-temporal cache line. Line g of a pair holds elements [16g, 16g+16) of the concatenated
// 3600-bit stream: lines 0..111 come from the first row, line 112 straddles the two rows, lines
// 113..224 from the second row at a byte-shifted offset.
__attribute__((target("avx512f,avx512bw,avx512vl"))) static void expand_avx512(const uint32_t *bits, float *dst,
                                                                                int64_t lo, int64_t hi)
{
    const __m512 one = _mm512_set1_ps(1.0f);
    int64_t e = lo;
    if ((e & 1) && e < hi) { expand_avx2(bits, dst, e, e + 1); ++e; }              // odd head
    for (; e + 1 < hi; e += 2) {
        const uint8_t *r0 = reinterpret_cast<const uint8_t *>(bits + e * 64);
        const uint8_t *r1 = reinterpret_cast<const uint8_t *>(bits + (e + 1) * 64);
        float *o = dst + e * 1800;
        for (int g = 0; g < 112; ++g) {
            uint16_t m;
            __builtin_memcpy(&m, r0 + 2 * g, 2);
            _mm512_stream_ps(o + 16 * g, _mm512_maskz_mov_ps((__mmask16)m, one));
        }
        _mm512_stream_ps(o + 16 * 112, _mm512_maskz_mov_ps((__mmask16)(r0[224] | (r1[0] << 8)), one));
        for (int g = 113; g < 225; ++g) {
            uint16_t m;
            __builtin_memcpy(&m, r1 + 2 * (g - 113) + 1, 2);
            _mm512_stream_ps(o + 16 * g, _mm512_maskz_mov_ps((__mmask16)m, one));
        }
    }
    if (e < hi) expand_avx2(bits, dst, e, hi);                                      // odd tail
    _mm_sfence();
}

// Expand envs [lo, hi) of `bits` ([n][64] u32, little-endian bit i = observation element i) into
// `dst` ([n][1800] f32) using up to `nthreads` threads.
void expand_f32(const uint32_t *bits, float *dst, int64_t lo, int64_t hi, int nthreads)
{
    if (hi <= lo) return;
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    static const bool have_avx512 = have_avx2 && __builtin_cpu_supports("avx512f") &&
                                    __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") &&
                                    !getenv("INV_NO_AVX512");
    const bool aligned64 = (reinterpret_cast<uintptr_t>(dst) & 63u) == 0;
    auto work = [&](int64_t a, int64_t b) {
        if (have_avx512 && aligned64) expand_avx512(bits, dst, a, b);
        else if (have_avx2) expand_avx2(bits, dst, a, b);
        else expand_scalar(bits, dst, a, b);
    };
    const int64_t n = hi - lo;
    int nt = (int)std::max<int64_t>(1, std::min<int64_t>(nthreads, n / 256));
    if (nt == 1) { work(lo, hi); return; }
    std::vector<std::thread> th;
    th.reserve(nt);
    auto cut = [&](int t) { // even boundaries keep the AVX-512 pairs aligned
        if (t == 0) return lo;
        if (t == nt) return hi;
        return (lo + n * t / nt) & ~(int64_t)1;
    };
    for (int t = 0; t < nt; ++t) th.emplace_back(work, cut(t), cut(t + 1));
    for (auto &x : th) x.join();
}

// Action ids on their way to the device: one pass copies them into the pinned staging buffer and
// checks 0 <= id <= 12 (discrete_to_action's ValueError, env_wrappers.py:66). Returns true if an id
// is out of range. AVX2: max_epu8(v, 12) differs from 12 exactly for the bytes above 12 (negative
// int8 ids are large unsigned bytes).
__attribute__((target("avx2"))) static bool stage_ids_avx2(int8_t *dst, const int8_t *src, int64_t n)
{
    const __m256i twelve = _mm256_set1_epi8(12);
    __m256i bad0 = _mm256_setzero_si256(), bad1 = _mm256_setzero_si256();
    int64_t i = 0;
    for (; i + 64 <= n; i += 64) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 32));
        _mm256_storeu_si256(reinterpret_cast<__m256i *>(dst + i), a);
        _mm256_storeu_si256(reinterpret_cast<__m256i *>(dst + i + 32), b);
        bad0 = _mm256_or_si256(bad0, _mm256_xor_si256(_mm256_max_epu8(a, twelve), twelve));
        bad1 = _mm256_or_si256(bad1, _mm256_xor_si256(_mm256_max_epu8(b, twelve), twelve));
    }
    unsigned bad = _mm256_testz_si256(_mm256_or_si256(bad0, bad1), _mm256_set1_epi8(-1)) ? 0u : 1u;
    for (; i < n; ++i) {
        const int8_t v = src[i];
        dst[i] = v;
        bad |= (unsigned)((uint8_t)v > 12);
    }
    return bad != 0;
}

// the same check without the copy (ids that already sit in page-locked memory go to the device from there)
__attribute__((target("avx2"))) static bool check_ids_avx2(const int8_t *src, int64_t n)
{
    const __m256i twelve = _mm256_set1_epi8(12);
    __m256i bad0 = _mm256_setzero_si256(), bad1 = _mm256_setzero_si256();
    int64_t i = 0;
    for (; i + 64 <= n; i += 64) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 32));
        bad0 = _mm256_or_si256(bad0, _mm256_xor_si256(_mm256_max_epu8(a, twelve), twelve));
        bad1 = _mm256_or_si256(bad1, _mm256_xor_si256(_mm256_max_epu8(b, twelve), twelve));
    }
    unsigned bad = _mm256_testz_si256(_mm256_or_si256(bad0, bad1), _mm256_set1_epi8(-1)) ? 0u : 1u;
    for (; i < n; ++i) bad |= (unsigned)((uint8_t)src[i] > 12);
    return bad != 0;
}

bool check_action_ids(const int8_t *src, int64_t n)
{
    static const bool have_avx2 = __builtin_cpu_supports("avx2") && !getenv("INV_NO_AVX2");
    if (have_avx2) return check_ids_avx2(src, n);
    unsigned bad = 0;
    for (int64_t i = 0; i < n; ++i) bad |= (unsigned)((uint8_t)src[i] > 12);
    return bad != 0;
}

bool stage_action_ids(int8_t *dst, const int8_t *src, int64_t n)
{
    static const bool have_avx2 = __builtin_cpu_supports("avx2") && !getenv("INV_NO_AVX2");
    if (have_avx2) return stage_ids_avx2(dst, src, n);
    unsigned bad = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int8_t v = src[i];
        dst[i] = v;
        bad |= (unsigned)((uint8_t)v > 12);
    }
    return bad != 0;
}

int hardware_threads()
{
    unsigned n = std::thread::hardware_concurrency();
    return n ? (int)n : 1;
}

} // namespace inv_host

// C ABI (include/inversus_b200.h): host-side expansion of packed observation rows, usable on its own.
extern "C" int inv_host_expand_f32(const uint32_t *bits, float *dst, int64_t first, int64_t count, int nthreads)
{
    if (!bits || !dst || first < 0 || count < 0) return -1;
    inv_host::expand_f32(bits, dst, first, first + count, nthreads > 0 ? nthreads : 1);
    return 0;
}

// C ABI: the id check of the *_host calls, usable on its own (no GPU involved). Copies ids[n] to
// staged[n] (staged may equal ids) and returns 0 if every id is in 0..12, INV_ERR_INVALID_ACTION
// (-3) otherwise -- discrete_to_action's ValueError, env_wrappers.py:66.
extern "C" int inv_host_stage_action_ids(const int8_t *ids, int8_t *staged, int64_t n)
{
    if (!ids || !staged || n < 0) return -1;
    if (staged == ids) return inv_host::check_action_ids(ids, n) ? -3 : 0; // in place: check only
    return inv_host::stage_action_ids(staged, ids, n) ? -3 : 0;
}

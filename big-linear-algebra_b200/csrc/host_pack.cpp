// host_pack.cpp -- the inner loop of the host-side byte packing of whole-number float batches (mlp.cu: step_packed).  Plain C++ with
// SSE2 intrinsics (the x86-64 baseline: no target flags), compiled by the host compiler alone.
#include <cstdint>
#include <cstring>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace bla {

// dst[i] = (unsigned char)src[i] for i < n; returns true when EVERY src[i] is bit for bit one of the floats 0.0f .. 255.0f -- no
// fraction, no negative zero, no NaN, nothing out of range -- i.e. when widening the bytes again reproduces the input exactly.
bool pack_row_u8(const float* src, unsigned char* dst, int n) {
    int i = 0;
    uint32_t bad = 0;
#if defined(__SSE2__)
    __m128i acc = _mm_setzero_si128();
    const __m128i low8 = _mm_set1_epi32(255);
    for (; i + 16 <= n; i += 16) {
        const __m128 v0 = _mm_loadu_ps(src + i), v1 = _mm_loadu_ps(src + i + 4), v2 = _mm_loadu_ps(src + i + 8), v3 = _mm_loadu_ps(src + i + 12);
        // truncating conversion: NaN and |v| >= 2^31 give 0x80000000, which fails both checks below
        const __m128i i0 = _mm_cvttps_epi32(v0), i1 = _mm_cvttps_epi32(v1), i2 = _mm_cvttps_epi32(v2), i3 = _mm_cvttps_epi32(v3);
        // (a) converting back must give the same BITS (fractions, -0.0f, NaN differ), (b) nothing outside 0..255
        acc = _mm_or_si128(acc, _mm_xor_si128(_mm_castps_si128(v0), _mm_castps_si128(_mm_cvtepi32_ps(i0))));
        acc = _mm_or_si128(acc, _mm_xor_si128(_mm_castps_si128(v1), _mm_castps_si128(_mm_cvtepi32_ps(i1))));
        acc = _mm_or_si128(acc, _mm_xor_si128(_mm_castps_si128(v2), _mm_castps_si128(_mm_cvtepi32_ps(i2))));
        acc = _mm_or_si128(acc, _mm_xor_si128(_mm_castps_si128(v3), _mm_castps_si128(_mm_cvtepi32_ps(i3))));
        acc = _mm_or_si128(acc, _mm_andnot_si128(low8, _mm_or_si128(_mm_or_si128(i0, i1), _mm_or_si128(i2, i3))));
        const __m128i b = _mm_packus_epi16(_mm_packs_epi32(i0, i1), _mm_packs_epi32(i2, i3));   // saturating: exact for 0..255
        _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), b);
    }
    alignas(16) uint32_t lanes[4];
    _mm_store_si128(reinterpret_cast<__m128i*>(lanes), acc);
    bad = lanes[0] | lanes[1] | lanes[2] | lanes[3];
#endif
    for (; i < n; ++i) {
        const float v = src[i];
        const int iv = (v >= 0.f && v < 256.f) ? (int)v : 256;
        const float back = (float)(iv & 255);
        uint32_t a, b2;
        memcpy(&a, &v, 4);
        memcpy(&b2, &back, 4);
        bad |= a ^ b2;
        dst[i] = (unsigned char)iv;
    }
    return bad == 0;
}

}  // namespace bla

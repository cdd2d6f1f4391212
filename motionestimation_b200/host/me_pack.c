/* me_pack.c -- the int -> u8 narrowing at the reference seam.
 *
 * The reference keeps a frame as one `int` per pixel (src/common/utils.c:49-53 widens the
 * bytes of the .yuv file; main.c:132-139 hands those arrays to the search).  The device
 * layout is the file's own 8 bits per pixel, so me_b200_search narrows the caller's frames
 * on the way to the pinned staging buffer.  This is the host-side hot loop of the drop-in
 * call (16.6 MB read, 4.1 MB written per 1080p pair): 16 pixels per step with SSE2 /
 * 32 with AVX2 (chosen once at run time), and me_api.cu spreads row chunks over a few
 * worker threads so the narrowing of chunk k+1 overlaps the upload of chunk k.
 *
 * Returns the OR of all input values: the caller rejects the frame when any bit above
 * bit 7 is set (a pixel outside 0..255 cannot be represented; negative ints have bit 31).
 */
#include <emmintrin.h>
#include <immintrin.h>
#include <stddef.h>
#include <stdint.h>

unsigned me_pack_int_to_u8(uint8_t *dst, const int *src, size_t n);  /* used by csrc/me_api.cu */

static unsigned pack_sse2(uint8_t *dst, const int *src, size_t n) {
  size_t i = 0;
  __m128i bad = _mm_setzero_si128();
  for (; i + 16 <= n; i += 16) {
    const __m128i a = _mm_loadu_si128((const __m128i *)(src + i));
    const __m128i b = _mm_loadu_si128((const __m128i *)(src + i + 4));
    const __m128i c = _mm_loadu_si128((const __m128i *)(src + i + 8));
    const __m128i d = _mm_loadu_si128((const __m128i *)(src + i + 12));
    bad = _mm_or_si128(bad, _mm_or_si128(_mm_or_si128(a, b), _mm_or_si128(c, d)));
    /* values that pass the range check are 0..255, so the saturating packs are plain narrowing */
    _mm_storeu_si128((__m128i *)(dst + i), _mm_packus_epi16(_mm_packs_epi32(a, b), _mm_packs_epi32(c, d)));
  }
  unsigned acc = 0;
  {
    uint32_t t[4];
    _mm_storeu_si128((__m128i *)t, bad);
    acc = t[0] | t[1] | t[2] | t[3];
  }
  for (; i < n; i++) {
    acc |= (unsigned)src[i];
    dst[i] = (uint8_t)src[i];
  }
  return acc;
}

__attribute__((target("avx2"))) static unsigned pack_avx2(uint8_t *dst, const int *src, size_t n) {
  size_t i = 0;
  __m256i bad = _mm256_setzero_si256();
  /* the 256-bit packs work per 128-bit lane; one cross-lane permute of the dwords restores the order */
  const __m256i order = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
  for (; i + 32 <= n; i += 32) {
    const __m256i a = _mm256_loadu_si256((const __m256i *)(src + i));
    const __m256i b = _mm256_loadu_si256((const __m256i *)(src + i + 8));
    const __m256i c = _mm256_loadu_si256((const __m256i *)(src + i + 16));
    const __m256i d = _mm256_loadu_si256((const __m256i *)(src + i + 24));
    bad = _mm256_or_si256(bad, _mm256_or_si256(_mm256_or_si256(a, b), _mm256_or_si256(c, d)));
    const __m256i ab = _mm256_packs_epi32(a, b), cd = _mm256_packs_epi32(c, d);
    const __m256i v = _mm256_permutevar8x32_epi32(_mm256_packus_epi16(ab, cd), order);
    _mm256_storeu_si256((__m256i *)(dst + i), v);
  }
  unsigned acc = 0;
  {
    uint32_t t[8];
    _mm256_storeu_si256((__m256i *)t, bad);
    for (int k = 0; k < 8; k++) acc |= t[k];
  }
  if (i < n) acc |= pack_sse2(dst + i, src + i, n - i);
  return acc;
}

unsigned me_pack_int_to_u8(uint8_t *dst, const int *src, size_t n) {
  static int use_avx2 = -1;
  if (use_avx2 < 0) use_avx2 = __builtin_cpu_supports("avx2") ? 1 : 0;
  return use_avx2 ? pack_avx2(dst, src, n) : pack_sse2(dst, src, n);
}

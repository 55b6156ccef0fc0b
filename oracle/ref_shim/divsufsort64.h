/* Hand-written stand-in for the public header cmake would generate from
 * libdivsufsort/include/divsufsort.h.cmake with W64BIT=64 (types + prototypes only).
 * Test infrastructure only: used by oracle/Makefile to compile the reference's C files in place.
 * The prototypes match what src/divsufsort.rs:8-33 binds. */
#ifndef _DIVSUFSORT64_H
#define _DIVSUFSORT64_H 1
#include <inttypes.h>
#ifdef __cplusplus
extern "C" {
#endif
#ifndef DIVSUFSORT_API
# define DIVSUFSORT_API __attribute__((visibility("default")))
#endif
#ifndef SAUCHAR_T
#define SAUCHAR_T
typedef uint8_t sauchar_t;
#endif
#ifndef SAINT_T
#define SAINT_T
typedef int32_t saint_t;
#endif
#ifndef SAIDX64_T
#define SAIDX64_T
typedef int64_t saidx64_t;
#endif
#ifndef PRIdSAINT_T
#define PRIdSAINT_T PRId32
#endif
#ifndef PRIdSAIDX64_T
#define PRIdSAIDX64_T PRId64
#endif
DIVSUFSORT_API saint_t divsufsort64(const sauchar_t *T, saidx64_t *SA, saidx64_t n);
DIVSUFSORT_API saidx64_t divbwt64(const sauchar_t *T, sauchar_t *U, saidx64_t *A, saidx64_t n);
DIVSUFSORT_API const char *divsufsort64_version(void);
DIVSUFSORT_API saint_t bw_transform64(const sauchar_t *T, sauchar_t *U, saidx64_t *SA, saidx64_t n, saidx64_t *idx);
DIVSUFSORT_API saint_t inverse_bw_transform64(const sauchar_t *T, sauchar_t *U, saidx64_t *A, saidx64_t n, saidx64_t idx);
DIVSUFSORT_API saint_t sufcheck64(const sauchar_t *T, const saidx64_t *SA, saidx64_t n, saint_t verbose);
DIVSUFSORT_API saidx64_t sa_search64(const sauchar_t *T, saidx64_t Tsize, const sauchar_t *P, saidx64_t Psize,
                                     const saidx64_t *SA, saidx64_t SAsize, saidx64_t *left);
DIVSUFSORT_API saidx64_t sa_searchb64(const sauchar_t *T, saidx64_t Tsize, const sauchar_t *P, saidx64_t Psize,
                                      const saidx64_t *SA, saidx64_t SAsize, saidx64_t *left,
                                      saidx64_t init_left, saidx64_t init_right);
DIVSUFSORT_API saidx64_t sa_simplesearch64(const sauchar_t *T, saidx64_t Tsize, const saidx64_t *SA, saidx64_t SAsize,
                                           saint_t c, saidx64_t *left);
#ifdef __cplusplus
}
#endif
#endif

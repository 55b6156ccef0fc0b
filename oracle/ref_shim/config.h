/* Hand-written stand-in for the header the reference's cmake would generate from
 * libdivsufsort/include/config.h.cmake.  Test infrastructure only (see oracle/README.md):
 * it lets oracle/Makefile compile the reference's four C files where they lie under
 * /root/reference/libdivsufsort/lib with plain gcc, without running cmake. */
#ifndef _CONFIG_H
#define _CONFIG_H 1
#define HAVE_INTTYPES_H 1
#define HAVE_STDDEF_H 1
#define HAVE_STDINT_H 1
#define HAVE_STDLIB_H 1
#define HAVE_STRING_H 1
#define HAVE_STRINGS_H 1
#define HAVE_MEMORY_H 1
#define HAVE_SYS_TYPES_H 1
#define PROJECT_VERSION_FULL "2.0.2-asgart-fork"
#ifndef INLINE
# define INLINE inline
#endif
#endif

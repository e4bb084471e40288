/* Build configuration used ONLY to compile the unmodified reference sources
 * (femto-dev/femto, mounted read-only at /root/reference) into oracle/_ref/.
 * Mirrors the knobs of the reference's config.h.cmake.in; test infrastructure. */
#ifndef FEMTO_B200_ORACLE_REF_CONFIG_H
#define FEMTO_B200_ORACLE_REF_CONFIG_H
#define PACKAGE_STRING "femto-reference (oracle/_ref build)"
#define HAVE_CLOCK_GETTIME_C 1
#define HAVE_CLOCK_GETTIME 1
#define HAVE_STATVFS 1
#define HAVE_DECL___SYNC_FETCH_AND_ADD 1
#define EXTRA_CHECKS 0
#ifndef restrict
#define restrict __restrict__
#endif
#endif

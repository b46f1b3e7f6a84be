/* COMPILE-CHECK ONLY: the reference generates src/version.h at build time */
#define GFS_VERSION "check"
#define GFS_BUILD_VERSION "check"

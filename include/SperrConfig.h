/* Version macros of the SPERR release whose bitstream and API libsperr_b200 implements (v0.8.5:
 * /root/reference/CMakeLists.txt, SperrConfig.h.in:1-13 -- the reference generates this file with
 * cmake). SPERR_VERSION_MAJOR is byte 0 of every container and is checked on decode
 * (src/SPERR3D_OMP_D.cpp:33-34). */
#ifndef SPERR_CONFIG
#define SPERR_CONFIG

#define SPERR_VERSION_MAJOR 0
#define SPERR_VERSION_MINOR 8
#define SPERR_VERSION_PATCH 5

#define SPERR_B200 1 /* the chunk loop runs on B200 GPUs, not on OpenMP threads */

#endif

/* codes_config.h for the reference-source oracle build: every optional subsystem off
 * (stands in for the file CMake generates from codes_config.h.cmake.in) */
#ifndef CODES_CONFIG_H
#define CODES_CONFIG_H
#define CODES_HAVE_DUMPI 0
#define CODES_HAVE_ONLINE 0
#define CODES_HAVE_SWM 0
#define CODES_HAVE_UNION 0
#define CODES_HAVE_RECORDER 0
#define CODES_HAVE_TORCH 0
#define CODES_HAVE_ZEROMQ 0
#define CODES_HAVE_DARSHAN 0
#define SWM_DATAROOTDIR ""
#define UNION_DATADIR ""
#endif

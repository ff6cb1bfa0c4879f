#ifndef __CONFIG_H__
#define __CONFIG_H__
#define SCIP_BUILD_TYPE "Release"
#define SCIP_VERSION_MAJOR 11
#define SCIP_VERSION_MINOR 0
#define SCIP_VERSION_PATCH 0
#define SCIP_VERSION_API 167
#define TPI_NONE
#define SCIP_THREADSAFE
#define WITH_SCIPDEF
#define SCIP_ROUNDING_FE
#endif

#define SCIP_BUILDFLAGS " gcc -O3 -DNDEBUG (oracle/Makefile.ref)"
#define SCIP_LPS "none"
#define SCIP_IPOPT "false"
#define SCIP_CONOPT "false"

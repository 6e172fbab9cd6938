/* TEST INFRASTRUCTURE ONLY: what CMake would generate from CMCConfig.h.in. */
#define CMC_VERSION_MAJOR 1
#define CMC_VERSION_MINOR 0
#define CMCPRETTYNAME "ClusterMonteCarlo (oracle build)"
#define CMCVERSION "oracle"
#define CMCNICK "oracle"
#define CMCDATE "n/a"

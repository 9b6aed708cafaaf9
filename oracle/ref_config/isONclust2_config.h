/* Stand-in for the header the reference's CMake generates from
 * src/isONclust2_config.h.in (CMakeLists.txt:35-38). Oracle build only. */
#define isONclust2_VERSION_MAJOR 2
#define isONclust2_VERSION_MINOR 4
#define isONclust2_VERSION "v2.4-oracle"

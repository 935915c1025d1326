/* lets unmodified reference sources (#include "zstd-mt.h", programs/zstd-mt.c:16) build against libzstdmt_b200 */
#include "zstdmt_b200_zstd.h"

/* lets unmodified reference sources (#include "lz4-mt.h", programs/lz4-mt.c:16) build against libzstdmt_b200 */
#include "zstdmt_b200_lz4.h"

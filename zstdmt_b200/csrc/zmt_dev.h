/* zmt_dev.h — internal alias of the public device-level header. */
#pragma once
#include "../../include/zstdmt_b200_dev.h"

// TEST INFRASTRUCTURE ONLY — see ../hwloc.h (single-node stand-in for libhwloc).
#pragma once
#include "../hwloc.h"

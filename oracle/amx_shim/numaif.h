// TEST INFRASTRUCTURE ONLY — single-NUMA-node stand-in for <numaif.h>, which this image lacks: lets the UNMODIFIED
// kt-kernel AMX backend (kt-kernel/operators/amx, cpu_backend/worker_pool.cpp) compile from /root/reference as a CPU baseline.
// One node (0) owning every CPU; binding calls succeed without binding.  Results measured through it are labelled "shimmed".
#pragma once

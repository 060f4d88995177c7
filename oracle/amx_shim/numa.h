// TEST INFRASTRUCTURE ONLY — single-NUMA-node stand-in for <numa.h> / libnuma, which this image lacks: lets the UNMODIFIED
// kt-kernel AMX backend (kt-kernel/operators/amx, cpu_backend/worker_pool.cpp) compile from /root/reference as a CPU baseline.
// One node (0) owning every CPU; binding calls succeed without binding.  Results measured through it are labelled "shimmed".
// single-node stand-in for <numa.h> (libnuma is not installed in this image): one NUMA node, node 0
#pragma once
#include <cstdlib>
struct bitmask { unsigned long size; unsigned long* maskp; };
static inline int numa_available(void) { return 0; }
static inline int numa_num_configured_nodes(void) { return 1; }
static inline int numa_max_node(void) { return 0; }
static inline int numa_node_of_cpu(int) { return 0; }
static inline struct bitmask* numa_bitmask_alloc(unsigned int n) { auto* b = (bitmask*)calloc(1, sizeof(bitmask)); b->size = n; b->maskp = (unsigned long*)calloc(1, sizeof(unsigned long)); return b; }
static inline struct bitmask* numa_bitmask_setbit(struct bitmask* b, unsigned int i) { b->maskp[0] |= 1ul << i; return b; }
static inline void numa_bitmask_free(struct bitmask* b) { free(b->maskp); free(b); }
static inline void numa_bind(struct bitmask*) {}
static inline void* numa_alloc_onnode(size_t sz, int) { void* p = nullptr; if (posix_memalign(&p, 64, sz)) return nullptr; return p; }
static inline void numa_free(void* p, size_t) { free(p); }
static inline int numa_run_on_node(int) { return 0; }
static inline void numa_set_preferred(int) {}

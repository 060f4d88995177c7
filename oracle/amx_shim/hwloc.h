// TEST INFRASTRUCTURE ONLY — single-NUMA-node stand-in for <hwloc.h> / libhwloc, which this image lacks: lets the UNMODIFIED
// kt-kernel AMX backend (kt-kernel/operators/amx, cpu_backend/worker_pool.cpp) compile from /root/reference as a CPU baseline.
// One node (0) owning every CPU; binding calls succeed without binding.  Results measured through it are labelled "shimmed".
// single-node stand-in for <hwloc.h>: one NUMA node holding every CPU; binding calls succeed without doing anything
#pragma once
#include <errno.h>
#include <pthread.h>
#include <string.h>
typedef struct hwloc_bitmap_s { int dummy; }* hwloc_bitmap_t;
typedef hwloc_bitmap_t hwloc_cpuset_t;
typedef hwloc_bitmap_t hwloc_nodeset_t;
typedef struct hwloc_obj { hwloc_cpuset_t cpuset; hwloc_nodeset_t nodeset; unsigned os_index; }* hwloc_obj_t;
typedef struct hwloc_topology { struct hwloc_obj node; struct hwloc_bitmap_s bm; }* hwloc_topology_t;
enum { HWLOC_OBJ_NUMANODE = 1, HWLOC_OBJ_CORE = 2 };
enum { HWLOC_MEMBIND_BIND = 1 };
enum { HWLOC_MEMBIND_THREAD = 1, HWLOC_MEMBIND_STRICT = 2, HWLOC_MEMBIND_BYNODESET = 4 };
enum { HWLOC_CPUBIND_STRICT = 1, HWLOC_CPUBIND_THREAD = 2 };
static inline int hwloc_topology_init(hwloc_topology_t* t) { *t = new hwloc_topology(); (*t)->node.cpuset = &(*t)->bm; (*t)->node.nodeset = &(*t)->bm; (*t)->node.os_index = 0; return 0; }
static inline int hwloc_topology_load(hwloc_topology_t) { return 0; }
static inline void hwloc_topology_destroy(hwloc_topology_t t) { delete t; }
static inline hwloc_obj_t hwloc_get_obj_by_type(hwloc_topology_t t, int, unsigned idx) { return idx == 0 ? &t->node : nullptr; }
static inline hwloc_obj_t hwloc_get_obj_inside_cpuset_by_type(hwloc_topology_t t, hwloc_cpuset_t, int, unsigned) { return &t->node; }
static inline int hwloc_set_membind(hwloc_topology_t, hwloc_nodeset_t, int, int) { return 0; }
static inline int hwloc_set_membind_nodeset(hwloc_topology_t, hwloc_nodeset_t, int, int) { return 0; }
static inline int hwloc_set_thread_cpubind(hwloc_topology_t, pthread_t, hwloc_cpuset_t, int) { return 0; }
static inline int hwloc_get_thread_cpubind(hwloc_topology_t, pthread_t, hwloc_cpuset_t, int) { return 0; }
static inline hwloc_bitmap_t hwloc_bitmap_alloc(void) { return new hwloc_bitmap_s(); }
static inline void hwloc_bitmap_free(hwloc_bitmap_t b) { delete b; }
static inline int hwloc_bitmap_copy(hwloc_bitmap_t, hwloc_bitmap_t) { return 0; }
static inline int hwloc_bitmap_singlify(hwloc_bitmap_t) { return 0; }
static inline int hwloc_bitmap_first(hwloc_bitmap_t) { return 0; }
#define hwloc_bitmap_foreach_begin(id, bm) do { id = 0; {
#define hwloc_bitmap_foreach_end() } } while (0)

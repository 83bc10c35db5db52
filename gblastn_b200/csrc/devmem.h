// devmem.h — device allocations of per-search data (query tables, uploaded volumes, chunk tables).
//
// Default: the stream-ordered pool (cudaMallocAsync / cudaFreeAsync).  A job pipeline (bn_prelim_search_jobs) loads a
// batch and a volume per job; dozens of pool allocations per job cost milliseconds of host time each whenever the pool
// has to grow or re-map memory (worse with several processes on one host), so a pipeline job draws from its lane's
// ARENA instead: one grow-only device buffer, bump-allocated, reset when the lane's next job starts.  The arena is
// thread-local state of the preparing thread; memory inside an arena is never freed individually.
#ifndef GBLASTN_B200_DEVMEM_H
#define GBLASTN_B200_DEVMEM_H
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace bn {

struct DevArena {
    uint8_t *base = nullptr;
    size_t cap = 0;
    size_t used = 0;        // bytes asked for since the last reset (also counted when they did not fit: the size to grow to)
};

// arena of the calling thread (nullptr: pool)
DevArena *&current_arena();
// ranges of all live arenas, so that dev_free recognises their memory whichever thread frees
void arena_register(const DevArena &a);
void arena_unregister(const DevArena &a);

cudaError_t dev_malloc(void **p, size_t bytes, cudaStream_t st);
cudaError_t dev_free(void *p, cudaStream_t st);

}  // namespace bn
#endif

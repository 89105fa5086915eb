// Process-wide caching allocator for device and pinned-host buffers.  cudaMalloc / cudaFree of the multi-GB
// factor slab and predict workspace cost far more than the kernels of a whole fit+predict step, so handles
// return their buffers here and the next handle of similar shape reuses them (mogp_trim releases the cache).
#pragma once
#include <cstddef>

namespace mogp {

// device == -1: pinned host memory.  Returns nullptr on failure (after releasing the cache and retrying once).
void* pool_alloc(size_t bytes, int device);
void pool_free(void* p);
// true if pool_alloc(bytes, device) would be served from the cache
bool pool_has_block(size_t bytes, int device);
void pool_trim();
size_t pool_cached_bytes();

}  // namespace mogp

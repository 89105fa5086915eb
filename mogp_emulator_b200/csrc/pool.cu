#include "pool.h"

#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <vector>

namespace mogp {

namespace {
struct Block {
    void* p;
    size_t bytes;
    int device;
};
std::mutex g_mu;
std::map<void*, Block> g_live;
std::vector<Block> g_free;

void release_all_locked() {
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto& b : g_free) {
        if (b.device < 0) {
            cudaFreeHost(b.p);
        } else {
            cudaSetDevice(b.device);
            cudaFree(b.p);
        }
    }
    g_free.clear();
    cudaSetDevice(cur);
    cudaGetLastError();
}

void* raw_alloc(size_t bytes, int device) {
    void* p = nullptr;
    cudaError_t e;
    if (device < 0) {
        e = cudaMallocHost(&p, bytes);
    } else {
        int cur = 0;
        cudaGetDevice(&cur);
        if (cur != device) cudaSetDevice(device);
        e = cudaMalloc(&p, bytes);
        if (cur != device) cudaSetDevice(cur);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
}  // namespace

void* pool_alloc(size_t bytes, int device) {
    if (bytes == 0) bytes = 1;
    std::lock_guard<std::mutex> lk(g_mu);
    // best fit among cached blocks of the same kind that waste at most 25%
    int best = -1;
    for (int i = 0; i < (int)g_free.size(); i++) {
        const Block& b = g_free[i];
        if (b.device != device || b.bytes < bytes || b.bytes > bytes + bytes / 4 + 4096) continue;
        if (best < 0 || b.bytes < g_free[best].bytes) best = i;
    }
    Block blk;
    if (best >= 0) {
        blk = g_free[best];
        g_free.erase(g_free.begin() + best);
    } else {
        void* p = raw_alloc(bytes, device);
        if (!p) {
            release_all_locked();
            p = raw_alloc(bytes, device);
            if (!p) return nullptr;
        }
        blk = Block{p, bytes, device};
    }
    g_live[blk.p] = blk;
    return blk.p;
}

void pool_free(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_live.find(p);
    if (it == g_live.end()) return;
    g_free.push_back(it->second);
    g_live.erase(it);
}

bool pool_has_block(size_t bytes, int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    for (const Block& b : g_free)
        if (b.device == device && b.bytes >= bytes && b.bytes <= bytes + bytes / 4 + 4096) return true;
    return false;
}

void pool_trim() {
    std::lock_guard<std::mutex> lk(g_mu);
    release_all_locked();
}

size_t pool_cached_bytes() {
    std::lock_guard<std::mutex> lk(g_mu);
    size_t s = 0;
    for (auto& b : g_free) s += b.bytes;
    return s;
}

}  // namespace mogp

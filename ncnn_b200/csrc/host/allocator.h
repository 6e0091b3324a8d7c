// allocator.h -- host and device allocators of the CUDA backend.
//
// Mirrors the reference's src/allocator.h: Allocator / PoolAllocator (:142-206) for host Mats and the
// VkAllocator family (:267-400) for device blobs:
//   CudaAllocator        <-> VkAllocator        (fastMalloc/fastFree of device memory)
//   CudaBlobAllocator    <-> VkBlobAllocator    (pooled, reused across layers; safe because every command of
//                                                one extract is ordered on one stream, allocator.cpp:695-760)
//   CudaWeightAllocator  <-> VkWeightAllocator  (layer weights, freed at Net::clear)
//   CudaStagingAllocator <-> VkStagingAllocator (pinned host memory for upload/download)
#ifndef NCNN_B200_ALLOCATOR_H
#define NCNN_B200_ALLOCATOR_H

#include <list>
#include <mutex>
#include <stdlib.h>
#include <utility>
#include <vector>

#include "platform.h"

namespace ncnn {

#define NCNN_MALLOC_ALIGN 64
#define NCNN_MALLOC_OVERREAD 64

static inline size_t alignSize(size_t sz, int n)
{
    return (sz + n - 1) & -(size_t)n;
}

static inline void* fastMalloc(size_t size)
{
    void* ptr = 0;
    if (posix_memalign(&ptr, NCNN_MALLOC_ALIGN, size + NCNN_MALLOC_OVERREAD)) ptr = 0;
    return ptr;
}

static inline void fastFree(void* ptr)
{
    if (ptr) free(ptr);
}

static inline int NCNN_XADD(int* addr, int delta)
{
    return __atomic_fetch_add(addr, delta, __ATOMIC_ACQ_REL);
}

class NCNN_EXPORT Allocator
{
public:
    virtual ~Allocator();
    virtual void* fastMalloc(size_t size) = 0;
    virtual void fastFree(void* ptr) = 0;
};

// best-fit free list with a size-compare ratio (reference: src/allocator.cpp:98-160)
class NCNN_EXPORT PoolAllocator : public Allocator
{
public:
    PoolAllocator();
    ~PoolAllocator();
    void set_size_compare_ratio(float scr);
    void clear();
    virtual void* fastMalloc(size_t size);
    virtual void fastFree(void* ptr);

private:
    std::mutex lock_;
    unsigned int size_compare_ratio_; // 0~256
    std::list<std::pair<size_t, void*> > budgets_;
    std::list<std::pair<size_t, void*> > payouts_;
};

// ------------------------------------------------------------------ device side
class NCNN_EXPORT CudaAllocator
{
public:
    explicit CudaAllocator(int device_index);
    virtual ~CudaAllocator();
    virtual void clear();
    virtual void* fastMalloc(size_t size) = 0;
    virtual void fastFree(void* ptr) = 0;
    // Placement hook of CudaMat::create_dims: an allocator may hand out a VIEW into memory it manages (own pitch, shared
    // refcount) for the blob being created -- how a producer writes straight into the channel range of its Concat's buffer
    // (CudaPlacementAllocator, net.cpp).  Default: no placement.
    virtual bool place(class CudaMat& /*m*/)
    {
        return false;
    }
    // the allocator a blob allocated through this one must be freed with (a placement allocator lives on the executor's stack)
    virtual CudaAllocator* real()
    {
        return this;
    }
    int device_index;
};

// Pooled device blobs.  Freed buffers go back to a best-fit free list and are handed out again without touching
// the driver: reuse is stream-ordered (one stream per extract), exactly the argument the reference makes for
// VkBlobAllocator.  180 GB of HBM means the pool can simply keep what a forward walk needs.
class NCNN_EXPORT CudaBlobAllocator : public CudaAllocator
{
public:
    explicit CudaBlobAllocator(int device_index);
    virtual ~CudaBlobAllocator();
    virtual void clear();
    virtual void* fastMalloc(size_t size);
    virtual void fastFree(void* ptr);
    size_t bytes_reserved() const
    {
        return reserved_;
    }

private:
    std::mutex lock_;
    std::list<std::pair<size_t, void*> > budgets_;
    std::list<std::pair<size_t, void*> > payouts_;
    size_t reserved_;
};

class NCNN_EXPORT CudaWeightAllocator : public CudaAllocator
{
public:
    explicit CudaWeightAllocator(int device_index);
    virtual ~CudaWeightAllocator();
    virtual void clear();
    virtual void* fastMalloc(size_t size);
    virtual void fastFree(void* ptr);

private:
    std::mutex lock_;
    std::vector<void*> blocks_;
};

// pinned host memory, pooled
class NCNN_EXPORT CudaStagingAllocator : public Allocator
{
public:
    CudaStagingAllocator();
    ~CudaStagingAllocator();
    void clear();
    virtual void* fastMalloc(size_t size);
    virtual void fastFree(void* ptr);

private:
    std::mutex lock_;
    std::list<std::pair<size_t, void*> > budgets_;
    std::list<std::pair<size_t, void*> > payouts_;
};

} // namespace ncnn

#endif // NCNN_B200_ALLOCATOR_H

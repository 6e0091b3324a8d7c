// allocator.cpp -- see allocator.h.  Device memory is reached only through the C ABI (include/ncnn_cuda.h).
#include "allocator.h"

#include "ncnn_cuda.h"

namespace ncnn {

Allocator::~Allocator()
{
}

// ------------------------------------------------------------------ generic best-fit pool over (size, ptr) lists
namespace {

typedef std::list<std::pair<size_t, void*> > BudgetList;

// smallest budget with size >= want and want >= size * ratio/256 (do not burn a huge block on a tiny request)
static void* take_budget(BudgetList& budgets, BudgetList& payouts, size_t size, unsigned int ratio)
{
    BudgetList::iterator best = budgets.end();
    for (BudgetList::iterator it = budgets.begin(); it != budgets.end(); ++it)
    {
        size_t bs = it->first;
        if (bs >= size && ((bs * ratio) >> 8) <= size)
        {
            if (best == budgets.end() || bs < best->first) best = it;
        }
    }
    if (best == budgets.end()) return 0;
    void* ptr = best->second;
    payouts.push_back(*best);
    budgets.erase(best);
    return ptr;
}

static bool give_back(BudgetList& budgets, BudgetList& payouts, void* ptr)
{
    for (BudgetList::iterator it = payouts.begin(); it != payouts.end(); ++it)
    {
        if (it->second == ptr)
        {
            budgets.push_back(*it);
            payouts.erase(it);
            return true;
        }
    }
    return false;
}

} // namespace

PoolAllocator::PoolAllocator()
    : size_compare_ratio_(192)
{
}

PoolAllocator::~PoolAllocator()
{
    clear();
    if (!payouts_.empty())
    {
        NCNN_LOGE("FATAL ERROR! pool allocator destroyed too early");
        for (BudgetList::iterator it = payouts_.begin(); it != payouts_.end(); ++it) NCNN_LOGE("%p still in use", it->second);
    }
}

void PoolAllocator::set_size_compare_ratio(float scr)
{
    if (scr < 0.f || scr > 1.f)
    {
        NCNN_LOGE("invalid size compare ratio %f", scr);
        return;
    }
    size_compare_ratio_ = (unsigned int)(scr * 256);
}

void PoolAllocator::clear()
{
    std::lock_guard<std::mutex> lk(lock_);
    for (BudgetList::iterator it = budgets_.begin(); it != budgets_.end(); ++it) ncnn::fastFree(it->second);
    budgets_.clear();
}

void* PoolAllocator::fastMalloc(size_t size)
{
    {
        std::lock_guard<std::mutex> lk(lock_);
        void* p = take_budget(budgets_, payouts_, size, size_compare_ratio_);
        if (p) return p;
    }
    void* ptr = ncnn::fastMalloc(size);
    if (!ptr) return 0;
    std::lock_guard<std::mutex> lk(lock_);
    payouts_.push_back(std::make_pair(size, ptr));
    return ptr;
}

void PoolAllocator::fastFree(void* ptr)
{
    if (!ptr) return;
    std::lock_guard<std::mutex> lk(lock_);
    if (!give_back(budgets_, payouts_, ptr))
    {
        NCNN_LOGE("FATAL ERROR! pool allocator get wild %p", ptr);
        ncnn::fastFree(ptr);
    }
}

// ------------------------------------------------------------------ device
CudaAllocator::CudaAllocator(int _device_index)
    : device_index(_device_index)
{
}

CudaAllocator::~CudaAllocator()
{
}

void CudaAllocator::clear()
{
}

CudaBlobAllocator::CudaBlobAllocator(int _device_index)
    : CudaAllocator(_device_index), reserved_(0)
{
}

CudaBlobAllocator::~CudaBlobAllocator()
{
    clear();
    for (BudgetList::iterator it = payouts_.begin(); it != payouts_.end(); ++it) ncnn_cuda_free(it->second);
}

void CudaBlobAllocator::clear()
{
    std::lock_guard<std::mutex> lk(lock_);
    for (BudgetList::iterator it = budgets_.begin(); it != budgets_.end(); ++it)
    {
        ncnn_cuda_free(it->second);
        reserved_ -= it->first;
    }
    budgets_.clear();
}

void* CudaBlobAllocator::fastMalloc(size_t size)
{
    size = alignSize(size, 512);
    {
        std::lock_guard<std::mutex> lk(lock_);
        void* p = take_budget(budgets_, payouts_, size, 128);
        if (p) return p;
    }
    void* ptr = 0;
    if (ncnn_cuda_malloc(&ptr, size) != 0 || !ptr)
    {
        // out of device memory: drop the idle budgets and retry once
        clear();
        if (ncnn_cuda_malloc(&ptr, size) != 0 || !ptr) return 0;
    }
    std::lock_guard<std::mutex> lk(lock_);
    payouts_.push_back(std::make_pair(size, ptr));
    reserved_ += size;
    return ptr;
}

void CudaBlobAllocator::fastFree(void* ptr)
{
    if (!ptr) return;
    std::lock_guard<std::mutex> lk(lock_);
    if (!give_back(budgets_, payouts_, ptr))
    {
        NCNN_LOGE("FATAL ERROR! cuda blob allocator get wild %p", ptr);
    }
}

CudaWeightAllocator::CudaWeightAllocator(int _device_index)
    : CudaAllocator(_device_index)
{
}

CudaWeightAllocator::~CudaWeightAllocator()
{
    clear();
}

void CudaWeightAllocator::clear()
{
    std::lock_guard<std::mutex> lk(lock_);
    for (size_t i = 0; i < blocks_.size(); i++) ncnn_cuda_free(blocks_[i]);
    blocks_.clear();
}

void* CudaWeightAllocator::fastMalloc(size_t size)
{
    void* ptr = 0;
    if (ncnn_cuda_malloc(&ptr, alignSize(size, 512)) != 0) return 0;
    std::lock_guard<std::mutex> lk(lock_);
    blocks_.push_back(ptr);
    return ptr;
}

void CudaWeightAllocator::fastFree(void* ptr)
{
    if (!ptr) return;
    std::lock_guard<std::mutex> lk(lock_);
    for (size_t i = 0; i < blocks_.size(); i++)
    {
        if (blocks_[i] == ptr)
        {
            ncnn_cuda_free(ptr);
            blocks_.erase(blocks_.begin() + i);
            return;
        }
    }
}

CudaStagingAllocator::CudaStagingAllocator()
{
}

CudaStagingAllocator::~CudaStagingAllocator()
{
    clear();
    for (BudgetList::iterator it = payouts_.begin(); it != payouts_.end(); ++it) ncnn_cuda_free_host(it->second);
}

void CudaStagingAllocator::clear()
{
    std::lock_guard<std::mutex> lk(lock_);
    for (BudgetList::iterator it = budgets_.begin(); it != budgets_.end(); ++it) ncnn_cuda_free_host(it->second);
    budgets_.clear();
}

void* CudaStagingAllocator::fastMalloc(size_t size)
{
    size = alignSize(size, 4096);
    {
        std::lock_guard<std::mutex> lk(lock_);
        void* p = take_budget(budgets_, payouts_, size, 128);
        if (p) return p;
    }
    void* ptr = 0;
    if (ncnn_cuda_malloc_host(&ptr, size) != 0 || !ptr) return 0;
    std::lock_guard<std::mutex> lk(lock_);
    payouts_.push_back(std::make_pair(size, ptr));
    return ptr;
}

void CudaStagingAllocator::fastFree(void* ptr)
{
    if (!ptr) return;
    std::lock_guard<std::mutex> lk(lock_);
    if (!give_back(budgets_, payouts_, ptr)) NCNN_LOGE("FATAL ERROR! cuda staging allocator get wild %p", ptr);
}

} // namespace ncnn

// C-ABI shim over the REFERENCE's own descriptor-slot bookkeeping, compiled from the sources where they lie:
//   /root/reference/include/DescriptorPool.h   FreeList (:25-44), DeviceDescriptors (:13-20), DescriptorPool::make (:62-76)
//   /root/reference/src/DescriptorPool.cc      constructor / destructor / slot_ptr (cudaMalloc fails without a GPU:
//                                              the slots are then null pointers, the bookkeeping is unaffected)
// Built by oracle/Makefile into oracle/_ref/libref_pool.so.  TEST INFRASTRUCTURE: tests/test_oracle_ref_pool.py uses
// it to pin oracle/frontend.py::FreeList and the slot semantics the product's ssb_sp_slot_* calls mirror.
#include <map>
#include <memory>

#include "DescriptorPool.h"

using superslam::DescriptorPool;
using superslam::DeviceDescriptors;
using superslam::FreeList;

namespace {
std::map<int, DeviceDescriptors> g_handles;
int g_next = 1;
}  // namespace

extern "C" {
void* ref_freelist_new(int n) { return new FreeList(n); }
void ref_freelist_delete(void* f) { delete static_cast<FreeList*>(f); }
int ref_freelist_acquire(void* f) { return static_cast<FreeList*>(f)->acquire(); }
void ref_freelist_release(void* f, int slot) { static_cast<FreeList*>(f)->release(slot); }
int ref_freelist_in_use(void* f) { return static_cast<FreeList*>(f)->in_use(); }

void* ref_pool_new(int num_slots, int max_keypoints, int dim) { return new DescriptorPool(num_slots, max_keypoints, dim); }
void ref_pool_delete(void* p) { delete static_cast<DescriptorPool*>(p); }
int ref_pool_in_use(void* p) { return static_cast<DescriptorPool*>(p)->in_use(); }
// DescriptorPool::make(count): returns a handle id; *slot receives the slot (-1 = exhausted), *count / *dim the shape
int ref_pool_make(void* p, int count, int* slot, int* out_count, int* out_dim) {
  DeviceDescriptors d = static_cast<DescriptorPool*>(p)->make(count);
  *slot = d.slot;
  *out_count = d.count;
  *out_dim = d.dim;
  const int id = g_next++;
  g_handles.emplace(id, std::move(d));
  return id;
}
int ref_handle_copy(int id) {   // copying the struct shares the slot (refcount)
  const int n = g_next++;
  g_handles.emplace(n, g_handles.at(id));
  return n;
}
void ref_handle_drop(int id) { g_handles.erase(id); }
int ref_handle_slot(int id) { return g_handles.at(id).slot; }
long ref_handle_use_count(int id) { return g_handles.at(id).slot_ref.use_count(); }
}

"""oracle/_ref: the reference's OWN descriptor-slot bookkeeping (include/DescriptorPool.h, src/DescriptorPool.cc),
compiled in place by oracle/Makefile, pins the restated FreeList of oracle/frontend.py and documents the handle
semantics the product's ssb_sp_slot_retain / release mirror."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle.frontend import FreeList

LIB = os.path.join(ROOT, "oracle", "_ref", "libref_pool.so")
pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref not built (needs /root/reference: build())")


@pytest.fixture(scope="module")
def ref():
    lib = C.CDLL(LIB)
    lib.ref_freelist_new.restype = C.c_void_p
    lib.ref_pool_new.restype = C.c_void_p
    lib.ref_handle_use_count.restype = C.c_long
    for f in ("ref_freelist_delete", "ref_freelist_acquire", "ref_freelist_in_use", "ref_pool_delete", "ref_pool_in_use"):
        getattr(lib, f).argtypes = [C.c_void_p]
    lib.ref_freelist_release.argtypes = [C.c_void_p, C.c_int]
    lib.ref_pool_make.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    return lib


def test_restated_freelist_equals_the_reference_on_random_sequences(ref):
    rng = np.random.default_rng(0)
    for n in (1, 3, 8):
        theirs, ours, held = ref.ref_freelist_new(n), FreeList(n), []
        for _ in range(400):
            if held and rng.random() < 0.45:
                s = held.pop(int(rng.integers(len(held))))
                ref.ref_freelist_release(theirs, s)
                ours.release(s)
            else:
                a, b = ref.ref_freelist_acquire(theirs), ours.acquire()
                assert a == b
                if a >= 0:
                    held.append(a)
            assert ref.ref_freelist_in_use(theirs) == ours.in_use() == len(held)
        ref.ref_freelist_delete(theirs)


def test_reference_pool_handle_semantics(ref):
    """DescriptorPool::make (include/DescriptorPool.h:62-76): 8 slots; a 9th live handle is empty (slot -1) and the
    frame simply has no descriptors (src/SuperPoint.cc:724-727); copies share the slot; the last copy to die returns
    it; the freed slot is the next one handed out; handles may outlive the pool."""
    pool = ref.ref_pool_new(8, 1024, 256)

    def make(count):
        slot, cnt, dim = C.c_int(), C.c_int(), C.c_int()
        h = ref.ref_pool_make(pool, count, C.byref(slot), C.byref(cnt), C.byref(dim))
        return h, slot.value, cnt.value, dim.value

    hs = [make(100 + i) for i in range(8)]
    assert [s for _, s, _, _ in hs] == list(range(8)) and hs[3][2:] == (103, 256)
    assert ref.ref_pool_in_use(pool) == 8
    h9, s9, c9, _ = make(50)
    assert s9 == -1 and c9 == 50 and ref.ref_pool_in_use(pool) == 8          # exhausted: empty handle, no error
    ref.ref_handle_drop(h9)
    keep = ref.ref_handle_copy(hs[5][0])                                       # e.g. VoEstimator's last_keyframe_
    assert ref.ref_handle_slot(keep) == 5 and ref.ref_handle_use_count(keep) == 2
    ref.ref_handle_drop(hs[5][0])
    assert ref.ref_pool_in_use(pool) == 8                                      # still referenced by the copy
    ref.ref_handle_drop(keep)
    assert ref.ref_pool_in_use(pool) == 7
    assert make(1)[1] == 5                                                     # LIFO: the freed slot comes back first
    survivor = hs[0][0]
    ref.ref_pool_delete(pool)                                                  # "a handle may outlive the pool"
    assert ref.ref_handle_slot(survivor) == 0
    ref.ref_handle_drop(survivor)                                              # releases into the shared free list: no crash

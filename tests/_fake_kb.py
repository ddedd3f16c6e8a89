"""TEST INFRASTRUCTURE: a stand-in for the `kvazaar_b200` module whose motion-search entry points run the HOST build of the
device code (tests/hostsim/libkvzme_hostsim.so) on CPU torch tensors.  It exists so that the bodies of the `-m gpu` tests
(views, strides, argument order, record layouts) are exercised on boxes without a GPU; it validates the TEST code, not the
device, and nothing outside tests/ may use it."""
import ctypes as C
import os

import numpy as np
import torch

import kvazaar_b200.api as api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOSTSIM = os.path.join(ROOT, "tests", "hostsim", "libkvzme_hostsim.so")


class FakeKB:
    LIB_PATH = HOSTSIM

    def __init__(self):
        self.lib = C.CDLL(HOSTSIM)
        self.launches = 0

    def init(self, device=0):
        pass

    def to_dev(self, a):
        a = np.ascontiguousarray(a)
        if a.dtype.fields is not None:
            return torch.from_numpy(a.view(np.uint8).copy())
        if a.dtype == np.uint16:
            a = a.view(np.int16)
        return torch.from_numpy(a.copy())

    def launch_count(self):
        return self.launches

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr())

    def me_search_batch(self, params, cur, ref, pus, out=None):
        count = pus.numel() // api.ME_PU.itemsize
        if out is None:
            out = torch.empty(count * api.ME_RESULT.itemsize, dtype=torch.uint8)
        rc = self.lib.kvz_cuda_me_search_batch(C.byref(params), self._p(cur), C.c_int(cur.stride(0)), self._p(ref), C.c_int(ref.stride(0)), self._p(pus),
                                               C.c_int(count), self._p(out), None)
        assert rc == 0
        self.launches += 1
        return out

    def me_frac_search_batch(self, params, fme_level, cur, ref, pus, out=None):
        count = pus.numel() // api.ME_PU.itemsize
        if out is None:
            out = torch.empty(count * api.ME_RESULT.itemsize, dtype=torch.uint8)
        rc = self.lib.kvz_cuda_me_frac_search_batch(C.byref(params), C.c_int(fme_level), self._p(cur), C.c_int(cur.stride(0)), self._p(ref),
                                                    C.c_int(ref.stride(0)), self._p(pus), C.c_int(count), self._p(out), None)
        assert rc == 0
        self.launches += 1
        return out

    def me_candidates_batch(self, frame, cus, col_cus, pus, out=None):
        count = pus.numel() // api.ME_CAND_PU.itemsize
        if out is None:
            out = torch.empty(count * api.ME_CAND_OUT.itemsize, dtype=torch.uint8)
        rc = self.lib.kvz_cuda_me_candidates_batch(C.byref(frame), self._p(cus), C.c_int(cus.stride(0) // api.ME_CU.itemsize), self._p(col_cus),
                                                   C.c_int(col_cus.stride(0) // api.ME_CU.itemsize), self._p(pus), C.c_int(count), self._p(out), None)
        assert rc == 0
        self.launches += 1
        return out

    def me_merge_cost_batch(self, params, refs, cur, pus, out=None):
        count = pus.numel() // api.ME_PU.itemsize
        if out is None:
            out = torch.empty(count * api.ME_MERGE_COST.itemsize, dtype=torch.uint8)
        rc = self.lib.kvz_cuda_me_merge_cost_batch(C.byref(params), C.byref(refs), self._p(cur), C.c_int(cur.stride(0)), self._p(pus), C.c_int(count),
                                                   self._p(out), None)
        assert rc == 0
        self.launches += 1
        return out

    def me_bipred_batch(self, params, refs, cur, pus, out=None):
        count = pus.numel() // api.ME_BIPRED_PU.itemsize
        if out is None:
            out = torch.empty(count * api.ME_BIPRED_RESULT.itemsize, dtype=torch.uint8)
        rc = self.lib.kvz_cuda_me_bipred_batch(C.byref(params), C.byref(refs), self._p(cur), C.c_int(cur.stride(0)), self._p(pus), C.c_int(count),
                                               self._p(out), None)
        assert rc == 0
        self.launches += 1
        return out

    def me_predict_batch(self, params, refs, pus, pred_y, pred_u, pred_v):
        count = pus.numel() // api.ME_MC_PU.itemsize
        rc = self.lib.kvz_cuda_me_predict_batch(C.byref(params), C.byref(refs), self._p(pus), C.c_int(count), self._p(pred_y), self._p(pred_u),
                                                self._p(pred_v), None)
        assert rc == 0
        self.launches += 1

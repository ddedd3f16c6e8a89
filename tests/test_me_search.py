"""Integer motion estimation (include/kvz_cuda.h: kvz_cuda_me_search_batch).

Checker: the UNMODIFIED reference's own search functions -- oracle/ref_me.c includes src/search_inter.c where it lies, so
select_starting_point / early_terminate / hexagon_search / diamond_search / check_mv_cost / calc_mvd_cost /
fracmv_within_tile run as compiled from the reference -- and tests/golden/me_search.npz, the same outputs committed
(tools/make_golden_me.py), for boxes without the reference build.
  * CPU tests: the host build of the device code (tests/hostsim/me_hostsim.cpp, one lane per PU) -- TEST INFRASTRUCTURE;
  * GPU tests: kvazaar_b200/libkvzcuda.so, through the C ABI (device-pointer entry and host-buffer entry).
Bar: best MV, bits and cost (IEEE double) identical for every PU.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from _me_cases import (CAND_CASES, CAND_OUT, CASES, GPU_FIRST_RUN_DONE, PU, RESULT, Params, grid_case, make_cand_case, make_case, run_cand_host_api, run_cand_reference,
                       run_host_api, run_reference, same)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOSTSIM = os.path.join(ROOT, "tests", "hostsim", "libkvzme_hostsim.so")


def _hostsim():
    if not os.path.exists(HOSTSIM):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unused-function", "-Wno-unknown-pragmas",
                               "-o", HOSTSIM, os.path.join(ROOT, "tests", "hostsim", "me_hostsim.cpp")])
    return C.CDLL(HOSTSIM)


def _golden(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", "me_search.npz"))
    out = np.zeros(len(g[name + "/bits"]), RESULT)
    out["mv"], out["bits"], out["cost"] = g[name + "/mv"], g[name + "/bits"], g[name + "/cost"]
    return out


def _explain(got, want, pus):
    bad = np.nonzero((got["mv"] != want["mv"]).any(1) | (got["bits"] != want["bits"]) | (got["cost"] != want["cost"]))[0]
    i = bad[0]
    return f"{len(bad)} of {len(pus)} PUs differ; first: PU {i} {pus[i]} got {got[i]} want {want[i]}"


# ------------------------------------------------------------------------------------------------ CPU
@pytest.mark.parametrize("name", sorted(CASES))
def test_reference_matches_golden(name, ref, ref10):
    """the committed golden outputs are what the compiled reference returns (pins the fixture to the reference)"""
    p, cur, rf, pus = make_case(name)
    want = run_reference(ref if p.bitdepth == 8 else ref10, p, cur, rf, pus)
    assert same(want, _golden(name)), _explain(want, _golden(name), pus)


@pytest.mark.parametrize("name", sorted(CASES))
def test_hostbuild_matches_golden(name):
    p, cur, rf, pus = make_case(name)
    got = run_host_api(_hostsim(), p, cur, rf, pus)
    assert same(got, _golden(name)), _explain(got, _golden(name), pus)
    assert got["points"].max() > 8                     # the searches do travel


def test_hostbuild_grid_matches_reference(ref):
    """every 16x16 PU of a 416x240 picture with the --preset slow settings (hexbs, early termination on, WPP + SAO limits)"""
    p, cur, rf, pus = grid_case(416, 240, 8)
    got = run_host_api(_hostsim(), p, cur, rf, pus)
    want = run_reference(ref, p, cur, rf, pus)
    assert same(got, want), _explain(got, want, pus)


def test_params_outside_scope_are_refused():
    lib = _hostsim()
    lib.kvz_cuda_me_params_supported.argtypes = [C.POINTER(Params)]
    p, _, _, _ = make_case("hexbs_et_sensitive")
    assert lib.kvz_cuda_me_params_supported(C.byref(p)) == 0
    for field, value in (("ime_algorithm", 8), ("ime_algorithm", -1), ("bitdepth", 12), ("mv_constraint", 5), ("me_early_termination", 3)):
        q = Params.from_buffer_copy(bytes(p))
        setattr(q, field, value)
        assert lib.kvz_cuda_me_params_supported(C.byref(q)) != 0, field


# ---- AMVP / merge candidates
def _golden_cand(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", "me_search.npz"))
    return g["cand/" + name].view(CAND_OUT)


def _explain_cand(got, want, pus):
    bad = [i for i in range(len(pus)) if got[i].tobytes() != want[i].tobytes()]
    i = bad[0]
    return f"{len(bad)} of {len(pus)} PUs differ; first: PU {i} {pus[i]}\n got  {got[i]}\n want {want[i]}"


@pytest.mark.parametrize("name", sorted(CAND_CASES))
def test_candidates_reference_matches_golden(name, ref):
    f, crp, clx, cus, col, pus = make_cand_case(name)
    want = run_cand_reference(ref, f, crp, clx, cus, col, pus)
    assert want.tobytes() == _golden_cand(name).tobytes(), _explain_cand(want, _golden_cand(name), pus)


@pytest.mark.parametrize("name", sorted(CAND_CASES))
def test_candidates_hostbuild_matches_golden(name):
    f, _, _, cus, col, pus = make_cand_case(name)
    got = run_cand_host_api(_hostsim(), f, cus, col, pus)
    want = _golden_cand(name)
    assert got.tobytes() == want.tobytes(), _explain_cand(got, want, pus)
    assert (np.abs(got["mv_cand"]).sum((1, 2, 3)) > 0).mean() > 0.5 and got["num_merge"].min() == f.max_merge


def _search_pus_from_candidates(cand_pus, cand, r):
    """the PU records of the search, fed with the derived candidates (list 0) -- what search_pu_inter hands to search_pu_inter_ref"""
    pus = np.zeros(len(cand_pus), PU)
    for k in ("x", "y", "w", "h"):
        pus[k] = cand_pus[k]
    pus["mv_cand"] = cand["mv_cand"][:, 0]
    pus["num_merge"] = cand["num_merge"]
    pus["merge"] = cand["merge"]
    pus["start_mv"] = r.integers(-24, 25, (len(pus), 2))
    return pus


def _chain_case():
    f, crp, clx, cus, col, cand_pus = make_cand_case("p_four_refs")
    p, cur, rf, _ = grid_case(f.width, f.height, 8)
    return f, crp, clx, cus, col, cand_pus, p, cur, rf


def test_hostbuild_candidates_feed_the_search(ref):
    """candidate derivation -> integer search, chained, against the reference doing the same chain"""
    f, crp, clx, cus, col, cand_pus, p, cur, rf = _chain_case()
    lib = _hostsim()
    cand = run_cand_host_api(lib, f, cus, col, cand_pus)
    want_cand = run_cand_reference(ref, f, crp, clx, cus, col, cand_pus)
    assert cand.tobytes() == want_cand.tobytes()
    pus = _search_pus_from_candidates(cand_pus, cand, np.random.default_rng(11))
    got, want = run_host_api(lib, p, cur, rf, pus), run_reference(ref, p, cur, rf, pus)
    assert same(got, want), _explain(got, want, pus)


# ------------------------------------------------------------------------------------------------ GPU (the product)
def _dev_api(kb, p, cur, rf, pus):
    """kvz_cuda_me_search_batch through the Python host layer: pictures, PU records and results in device memory"""
    import torch
    d_cur, d_ref, d_pus = kb.to_dev(cur), kb.to_dev(rf), kb.to_dev(pus)
    before = kb.launch_count()
    d_out = kb.me_search_batch(p, d_cur, d_ref, d_pus)
    torch.cuda.synchronize()
    assert kb.launch_count() == before + 1
    return d_out.cpu().numpy().view(RESULT).copy()


def check_cuda_case(kb, name, ref, ref10):
    p, cur, rf, pus = make_case(name)
    got = _dev_api(kb, p, cur, rf, pus)
    assert same(got, _golden(name)), _explain(got, _golden(name), pus)
    want = run_reference(ref if p.bitdepth == 8 else ref10, p, cur, rf, pus)
    assert same(got, want), _explain(got, want, pus)
    got_host = run_host_api(C.CDLL(kb.LIB_PATH), p, cur, rf, pus)          # kvz_cuda_call_me_search: host buffers
    assert same(got_host, want) and np.array_equal(got_host["points"], got["points"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GPU_FIRST_RUN_DONE))
def test_cuda_matches_golden_and_reference(cuda_lib, name, ref, ref10):
    check_cuda_case(cuda_lib, name, ref, ref10)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,bd,size", [(1920, 1080, 8, 16), (1920, 1080, 10, 32), (3840, 2160, 8, 64), (832, 480, 8, 8)])
def test_cuda_full_picture_matches_reference(cuda_lib, ref, ref10, w, h, bd, size):
    """every size x size PU of a full picture (BASELINE config 4's frame size and settings) vs the reference's functions"""
    p, cur, rf, pus = grid_case(w, h, bd, size)
    got = _dev_api(cuda_lib, p, cur, rf, pus)
    want = run_reference(ref if bd == 8 else ref10, p, cur, rf, pus)
    assert same(got, want), _explain(got, want, pus)
    assert (np.abs(got["mv"]).sum(1) > 0).mean() > 0.5


def _cand_dev_api(kb, f, cus, col, pus):
    import torch
    d_cus = kb.to_dev(cus.view(np.uint8).reshape(cus.shape[0], -1))
    d_col = kb.to_dev(col.view(np.uint8).reshape(col.shape[0], -1))
    d_out = kb.me_candidates_batch(f, d_cus, d_col, kb.to_dev(pus))
    torch.cuda.synchronize()
    return d_out.cpu().numpy().view(CAND_OUT).copy()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CAND_CASES))
def test_cuda_candidates_match_golden_and_reference(cuda_lib, ref, name):
    kb = cuda_lib
    f, crp, clx, cus, col, pus = make_cand_case(name)
    got = _cand_dev_api(kb, f, cus, col, pus)
    assert got.tobytes() == _golden_cand(name).tobytes(), _explain_cand(got, _golden_cand(name), pus)
    want = run_cand_reference(ref, f, crp, clx, cus, col, pus)
    assert got.tobytes() == want.tobytes(), _explain_cand(got, want, pus)
    got_host = run_cand_host_api(C.CDLL(kb.LIB_PATH), f, cus, col, pus)       # kvz_cuda_call_me_candidates: host buffers
    assert got_host.tobytes() == want.tobytes()


@pytest.mark.gpu
def test_cuda_candidates_feed_the_search(cuda_lib, ref):
    """candidate derivation -> integer search on the device, the candidates staying in device memory in between"""
    import torch
    kb = cuda_lib
    f, crp, clx, cus, col, cand_pus, p, cur, rf = _chain_case()
    cand = _cand_dev_api(kb, f, cus, col, cand_pus)
    assert cand.tobytes() == run_cand_reference(ref, f, crp, clx, cus, col, cand_pus).tobytes()
    pus = _search_pus_from_candidates(cand_pus, cand, np.random.default_rng(11))
    got, want = _dev_api(kb, p, cur, rf, pus), run_reference(ref, p, cur, rf, pus)
    assert same(got, want), _explain(got, want, pus)
    torch.cuda.synchronize()


def test_bad_pu_records_do_not_fault():
    """records outside the picture / with impossible sizes come back as "nothing found" instead of reading out of bounds"""
    from _me_cases import MERGE_COST, make_merge_case, run_frac_host_api, run_merge_host_api
    lib = _hostsim()
    p, cur, rf, pus = make_case("hexbs_et_sensitive")
    bad = pus[:6].copy()
    bad[0]["x"] = p.width - 4            # sticks out on the right
    bad[1]["y"] = -8
    bad[2]["w"] = 0
    bad[3]["h"] = 128
    bad[4]["w"] = 10                     # not a multiple of 4
    bad[5]["num_merge"] = 9
    for got in (run_host_api(lib, p, cur, rf, bad), run_frac_host_api(lib, p, 4, cur, rf, bad)):
        assert (got["bits"] == 0x7FFFFFFF).all() and (got["cost"] == 1.7e308).all() and (got["mv"] == 0).all()
    mp, mc, mcur, mplanes, mpus, _ = make_merge_case("merge_b_two_refs")
    mbad = mpus[:3].copy()
    mbad[0]["x"], mbad[1]["h"], mbad[2]["num_merge"] = mp.width, 3, 6
    m = run_merge_host_api(lib, mp, mc, mcur, mplanes, mbad, (1.0, 1.0, 1.0))
    assert (m["size"] == 0).all()
    assert m.dtype == MERGE_COST

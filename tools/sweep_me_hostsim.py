#!/usr/bin/env python
"""Randomised sweep of the motion-search host build (the single-source device code) against the reference's own functions:
random picture sizes, algorithms, limits, QPs, bit depths and PU lists for the integer search, the fractional search from
its results, the candidate derivation and the merge analysis.  CPU only (needs oracle/_ref).   python tools/sweep_me_hostsim.py <seed> <count>"""
import ctypes as C
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import me_cases as M  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
count = int(sys.argv[2]) if len(sys.argv) > 2 else 50
rnd = random.Random(seed)
host = C.CDLL(os.path.join(ROOT, "tests", "hostsim", "libkvzme_hostsim.so"))
refs = {8: M.RefShim(8), 10: M.RefShim(10)}
bad = 0
for it in range(count):
    w, h = rnd.choice([64, 72, 136, 200, 264, 320]), rnd.choice([64, 72, 136, 192])
    bd = rnd.choice([8, 8, 10])
    algo = rnd.choice([0, 0, 0, 7, 1, 3])
    name = f"sweep{seed}_{it}"
    M.CASES[name] = dict(w=w, h=h, bd=bd, algo=algo, steps=rnd.choice([-1, -1, 0, 1, 3]), et=rnd.choice([0, 1, 2]), mvc=rnd.choice([0, 0, 1, 4]),
                         wpp=rnd.choice([0, 1]), delay=rnd.choice([0, 8, 10]), qp=rnd.randint(10, 45), seed=1000 * seed + it, n=60 if algo == 3 else 150,
                         noisy=rnd.random() < 0.3, satd_final=int(rnd.random() < 0.2))
    p, cur, rf, pus = M.make_case(name)
    got, want = M.run_host_api(host, p, cur, rf, pus), M.run_reference(refs[bd], p, cur, rf, pus)
    ok = M.same(got, want)
    # fractional search from the integer results
    level = rnd.randint(1, 4)
    pus2 = pus.copy()
    pus2["start_mv"] = (want["mv"] >> 2) << 2
    p.satd_final = 0
    ok2 = M.same(M.run_frac_host_api(host, p, level, cur, rf, pus2), M.run_frac_reference(refs[bd], p, level, cur, rf, pus2))
    # candidates
    cname = f"csweep{seed}_{it}"
    nrefs = rnd.randint(1, 4)
    slice_b = rnd.random() < 0.5
    poc = rnd.randint(1, 40)
    pocs = rnd.sample([x for x in range(max(0, poc - 12), poc + 12) if x != poc], nrefs)
    l0 = [i for i in range(nrefs)][:rnd.randint(1, nrefs)]
    l1 = (rnd.sample(range(nrefs), rnd.randint(1, nrefs)) if slice_b else [])
    M.CAND_CASES[cname] = dict(w=w, h=h, poc=poc, slice_b=int(slice_b), tmvp=rnd.choice([0, 1, 1]), max_merge=rnd.randint(1, 5), pocs=pocs, l0=l0, l1=l1,
                               seed=2000 * seed + it, n=200)
    f, crp, clx, cus, col, cpus = M.make_cand_case(cname)
    ok3 = M.run_cand_host_api(host, f, cus, col, cpus).tobytes() == M.run_cand_reference(refs[8], f, crp, clx, cus, col, cpus).tobytes()
    # merge analysis
    mname = f"msweep{seed}_{it}"
    npic = rnd.randint(1, 4)
    ml0 = [rnd.randrange(npic) for _ in range(rnd.randint(1, 3))]
    ml1 = [rnd.randrange(npic) for _ in range(rnd.randint(0, 3))]
    M.MERGE_CASES[mname] = dict(w=w, h=h, bd=bd, pics=npic, l0=ml0, l1=ml1, bipred=int(bool(ml1) and rnd.random() < 0.8), mvc=rnd.choice([0, 0, 1, 4]),
                                wpp=rnd.choice([0, 1]), delay=rnd.choice([0, 8, 10]), qp=rnd.randint(10, 45), ctx=(rnd.randint(0, 125), rnd.randint(0, 125)),
                                seed=3000 * seed + it, n=120)
    mp, mc, mcur, mplanes, mpus, mcu = M.make_merge_case(mname)
    mwant, mbits = M.run_merge_reference(refs[bd], mp, mc, mcur, mplanes, mpus, mcu)
    ok4 = M.run_merge_host_api(host, mp, mc, mcur, mplanes, mpus, mbits).tobytes() == mwant.tobytes()
    ok3 = ok3 and ok4
    if ml1:                                                   # bi-prediction of two uni-predictions on the same pictures / lists
        M.BIPRED_CASES.append(mname)
        bp, bc, bcur, bplanes, bpus = M.make_bipred_case(mname)
        ok3 = ok3 and M.run_bipred_host_api(host, bp, bc, bcur, bplanes, bpus).tobytes() == M.run_bipred_reference(refs[bd], bp, bc, bcur, bplanes, bpus).tobytes()
    # motion compensation on the same pictures / lists
    M.MC_CASES.append(mname)
    cp, cc, cplanes, cus_, cvs_, cpus_, ccu_ = M.make_mc_case(mname)
    gw = M.run_mc_reference(refs[bd], cp, cc, cplanes, cus_, cvs_, cpus_, ccu_)
    gg = M.run_mc_host_api(host, cp, cc, cplanes, cus_, cvs_, cpus_)
    ok3 = ok3 and all(np.array_equal(a, b) for a, b in zip(gg, gw))
    if not (ok and ok2 and ok3):
        bad += 1
        print("DIFF", M.CASES[name], "integer", ok, "frac level", level, ok2, "cand+merge", M.CAND_CASES[cname], ok3, "merge", M.MERGE_CASES[mname], ok4, flush=True)
print("done", count, "bad", bad, flush=True)

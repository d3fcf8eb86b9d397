#!/usr/bin/env python3
"""Stage-by-stage GPU-vs-oracle report for one frame per config (debugging aid; the pytest -m gpu suite is the gate)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monoorbslam3_b200 import ORBExtractor, synth
from oracle import orb_oracle as orc


def check(w, h, nf, profile, use_tma, seed=1000):
    img = synth.frame(h, w, seed, profile)
    ex = ORBExtractor(nf, 1.2, 8, 20, 7, use_tma=use_tma, keep_stages=True)
    t0 = time.time(); kps, desc = ex(img); t1 = time.time()
    oc = orc.Extractor(nf, 1.2, 8, 20, 7)
    okps, odesc = oc(img)
    print("== %dx%d nf=%d %s tma=%s: gpu %d kps in %.1f ms, oracle %d" % (w, h, nf, profile, use_tma, len(kps), 1e3 * (t1 - t0), len(okps)))
    ok = True
    for l in range(8):
        a = ex.level_image(l); b = oc.level_image(l)
        same_img = a.shape == b.shape and np.array_equal(a, b)
        ab = ex.level_image(l, blurred=True); bb = oc.level_blurred(l)
        same_blur = bb is None or (ab.shape == bb.shape and np.array_equal(ab, bb))
        c = ex.level_candidates(l); oc_c = oc.level_candidates(l)
        oc_arr = np.stack([oc_c['x'], oc_c['y'], oc_c['score']], 1) if len(oc_c) else np.zeros((0, 3), np.int32)
        same_c = c.shape == oc_arr.shape and np.array_equal(c, oc_arr)
        k = ex.level_keypoints(l); ok_l = oc.level_keypoints(l)
        ok_arr = np.stack([ok_l['x'], ok_l['y'], ok_l['response']], 1).astype(np.int32) if len(ok_l) else np.zeros((0, 3), np.int32)
        same_k = k.shape == ok_arr.shape and np.array_equal(k, ok_arr)
        print("  L%d img %s (%s) blur %s cand %s (%d vs %d) kp %s (%d vs %d)" % (l, same_img, a.shape, same_blur, same_c, len(c), len(oc_arr), same_k, len(k), len(ok_arr)))
        if not same_img:
            d = np.argwhere(a != b) if a.shape == b.shape else None
            print("     img diffs:", None if d is None else (len(d), d[:5].tolist()))
        if not same_blur and ab.shape == bb.shape:
            d = np.argwhere(ab != bb); print("     blur diffs:", len(d), d[:5].tolist())
        if not same_c and len(c) and len(oc_arr):
            m = min(len(c), len(oc_arr)); d = np.argwhere((c[:m] != oc_arr[:m]).any(1)); print("     cand first diff:", d[:3].tolist(), c[:3].tolist(), oc_arr[:3].tolist())
        ok &= same_img and same_blur and same_c and same_k
    same = len(kps) == len(okps)
    if same:
        for f in ("x", "y", "size", "response", "octave", "class_id"):
            if not np.array_equal(kps[f], okps[f]): print("  field", f, "differs"); same = False
        da = np.abs(kps["angle"] - okps["angle"]).max() if len(kps) else 0
        nd = (desc != odesc).any(1).sum() if len(kps) else 0
        print("  final: max angle diff %.3g, angle bit-equal %s, descriptor rows differing %d / %d" % (da, np.array_equal(kps["angle"], okps["angle"]), nd, len(kps)))
        same &= da <= 1e-3 and nd <= 0.001 * len(kps)
    print("  RESULT", "OK" if ok and same else "MISMATCH")
    ex.close()
    return ok and same


if __name__ == "__main__":
    res = []
    for tma in (False, True):
        res.append(check(752, 480, 1000, "dense", tma))
        res.append(check(752, 480, 1000, "natural", tma))
        res.append(check(1241, 376, 2000, "dense", tma))
    res.append(check(1920, 1080, 4000, "dense", True))
    res.append(check(1920, 1080, 8000, "natural", True))
    print("ALL OK" if all(res) else "SOME MISMATCH")
    sys.exit(0 if all(res) else 1)

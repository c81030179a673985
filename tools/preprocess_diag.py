"""Diagnostic for the preprocessing parity tests: prints, per case and per leg (cuda vs reference kernels, cuda vs oracle,
oracle vs reference kernels), the numbers tests/test_gpu_preprocess.py::compare asserts on -- instead of stopping at the
first failed assertion.

    gpurun -- python tools/preprocess_diag.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))


def s8_pair(n16):
    return (n16 & 0xff).astype(np.int8).astype(np.int32), (n16 >> 8).astype(np.int8).astype(np.int32)


def metrics(got, want, rtf):
    gd, gn, gr, gc, gmin, gmax = got
    wd, wn, wr, wc, wmin, wmax = want
    gv, valid = (gd & 0x8000) == 0, (wd & 0x8000) == 0
    out = {"valid": float(valid.mean()), "mask_diff": int((gv != valid).sum())}
    both = gv & valid
    dd = np.abs(gd[both].astype(np.int32) - wd[both].astype(np.int32))
    out["depth_max"] = int(dd.max()) if dd.size else 0
    out["depth_frac"] = float(np.mean(dd != 0)) if dd.size else 0.0
    same = both & (gd == wd)
    nb = same.copy()
    nb[1:] &= same[:-1]; nb[:-1] &= same[1:]; nb[:, 1:] &= same[:, :-1]; nb[:, :-1] &= same[:, 1:]
    out["nb_frac_of_valid"] = float(nb.sum() / max(valid.sum(), 1))
    ax, ay = s8_pair(gn[nb]); bx, by = s8_pair(wn[nb])
    if ax.size:
        out["normal_max"] = int(max(np.abs(ax - bx).max(), np.abs(ay - by).max()))
        out["normal_frac"] = float(np.mean((ax != bx) | (ay != by)))
        ra, rb = gr[nb].view(np.float16).astype(np.float64), wr[nb].view(np.float16).astype(np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            rel = np.where(rb != 0, np.abs(ra - rb) / np.abs(rb), np.where(ra != 0, np.inf, 0.0))
        out["radius_rel_max"] = float(rel.max())
        out["radius_frac"] = float(np.mean(ra != rb))
        out["radius_viol"] = int(np.sum(np.abs(ra - rb) > 2.0 ** -9 * rb))
    out["border_normals_zero"] = bool(np.all(gn[0] == 0) and np.all(gn[:, 0] == 0))
    if wc is not None and gc is not None:
        out["rgba_diff"] = int((gc != wc).sum())
    out["min"] = (gmin, wmin, abs(gmin - wmin) / rtf if np.isfinite(gmin) and np.isfinite(wmin) else None)
    out["max"] = (gmax, wmax, abs(gmax - wmax) / rtf)
    return out


def main():
    import torch
    from badslam_b200 import scene as S
    from badslam_b200.direct_ba import DirectBA
    from oracle import cpu_oracle as O, ref_cuda as R
    from test_gpu_preprocess import run_cuda

    def three(sc, raw, rgb, tag):
        ba, ref = DirectBA.from_scene(sc), R.RefDirectBA(sc)
        got = run_cuda(ba, raw, rgb)
        want_ref = ref.preprocess_frame(raw, rgb)
        want_orc = O.Oracle(sc).preprocess_frame(raw, rgb)
        rtf = sc.cfg.raw_to_float_depth
        print(f"== {tag}")
        print("  cuda vs ref   :", metrics(got, want_ref, rtf))
        print("  cuda vs oracle:", metrics(got, want_orc, rtf))
        print("  oracle vs ref :", metrics(want_orc, want_ref, rtf), flush=True)

    for name, kf in (("small", 0), ("small", 1)):
        sc = S.make_scene(S.config_by_name(name))
        rng = np.random.default_rng(5)
        sc.depth_a = 0.02
        sc.cfactor = (2e-3 * rng.random(sc.cfactor.shape)).astype(np.float32)
        raw, rgb = S.raw_frame(sc, kf)
        raw[100:103, :] = 0
        three(sc, raw, rgb, f"{name}/{kf}")
    for (w, h) in ((70, 45), (641, 479)):
        sc = S.blank_scene(w, h)
        raw, rgb = S.random_raw_frame(w, h, seed=w * 100 + h)
        three(sc, raw, rgb, f"ragged {w}x{h}")


if __name__ == "__main__":
    main()

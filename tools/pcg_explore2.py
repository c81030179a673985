"""GPU exploration: PCG building blocks (r, M, p0, g, alpha_n, alpha_d), ours vs reference kernels vs CPU oracle."""
import dataclasses, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from badslam_b200 import scene as S
from badslam_b200.direct_ba import DirectBA
from oracle import cpu_oracle as O, ref_cuda


def cmp(name, a, b, K, n, stride, extra):
    segs = [("pose", 0, 6 * (K - 1)), ("surfel", 6 * (K - 1), 6 * (K - 1) + stride * n)]
    if extra:
        segs.append(("intr", 6 * (K - 1) + stride * n, len(a)))
    out = []
    for nm, lo, hi in segs:
        d = np.abs(a[lo:hi].astype(np.float64) - b[lo:hi])
        sc = np.abs(b[lo:hi]).max() + 1e-30
        out.append(f"{nm} max|d|/max {d.max()/sc:.2e} (at {int(d.argmax())})")
    return f"{name}: " + "; ".join(out)


def run(name, distort, intr, use_depth=True, use_desc=True, a_init=0.0):
    cfg = S.config_by_name(name)
    if distort:
        cfg = dataclasses.replace(cfg, depth_a=0.03, cfactor=0.005)
    sc = S.make_scene(cfg)
    K, n = cfg.num_keyframes, sc.num_surfels
    ba = DirectBA.from_scene(sc, use_depth_residuals=use_depth, use_descriptor_residuals=use_desc)
    ref = ref_cuda.RefDirectBA(sc, use_depth, use_desc)
    orc = O.Oracle(sc, use_depth, use_desc)
    if a_init:
        cf = (np.random.default_rng(5).standard_normal(sc.cfactor.shape) * 0.003).astype(np.float32)
        ba.SetA(a_init); ba.SetCFactorBuffer(cf); ref.set_depth_params(a_init, cf); orc.model.a = a_init; orc.cfactor[:] = cf
    kw = dict(optimize_depth_intrinsics=intr, optimize_color_intrinsics=intr, gauge_keyframe=1)
    A, B, Cc = ba.PCGDebug(**kw), ref.pcg_debug(**kw), orc.pcg_debug(**kw)
    stride = 3 if use_desc else 1
    print(f"=== {name} distort={distort} intr={intr} depth={use_depth} desc={use_desc} a={a_init}: unknowns {len(A[0])}")
    for idx, nm in enumerate(("r", "M", "p", "g")):
        print("  ours-ref", cmp(nm, A[idx], B[idx], K, n, stride, intr))
        print("  orc -ref", cmp(nm, Cc[idx], B[idx], K, n, stride, intr))
    print("  scalars ours", A[4], "ref", B[4], "orc", Cc[4])


if __name__ == "__main__":
    run("tiny", False, False)
    run("tiny", False, False, use_desc=False)
    run("small", True, True)
    run("small", True, True, a_init=0.02)
    run("cfg2", False, False)

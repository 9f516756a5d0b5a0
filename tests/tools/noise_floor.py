"""Parity noise floor (run on the GPU box): per 10 steps, the relative position error between
  refA/refB : two runs of the REFERENCE's own CUDA kernels (float atomics reorder run to run)
  eng0/eng1 : this engine, Newton-polar and always-SVD rotation paths
  orc       : the CPU oracle (float), orc64: its double-precision twin
for the parity scenes.  Writes gpurun_out/noise_floor.json and prints a table."""
import importlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import meshes  # noqa: E402
import oracle as O  # noqa: E402
import ref  # noqa: E402
from test_gpu_parity import _fixed_arrays  # noqa: E402

if __name__ == "__main__":
    pd = importlib.import_module("soft-body-simulation-cuda_b200")
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        assets = meshes.write_assets(tmp)
        for ctx, steps, with_f64 in [("C1 cube", 100, True), ("C5 house&sphere", 100, True), ("Armadillo&house", 100, False),
                                     ("C2 armadillo&bunny", 100, False)]:
            sc = pd.Scene.from_json(assets["json"], ctx)
            p = sc.params
            if ctx == "C1 cube":
                p["dt"] = 1 / 60
                sc.params = p
            a = sc.arrays()
            nA = int(a["body_vert_start"][1]) if (ctx.startswith("C2") and len(a["body_vert_start"]) > 1) else a["X"].shape[0]
            scale = float(np.linalg.norm(a["X"][:nA].max(0) - a["X"][:nA].min(0)))
            planes, spheres, cyls = _fixed_arrays(pd, a["fixed"])
            kw = dict(dt=p["dt"], gravity=p["gravity"], rho=p["rho"], muN=p["muN"], muT=p["muT"], num_iterations=p["num_iterations"])
            refA = ref.RefScene(a["X"], a["Tet"], a["mass"], a["mu"], planes=planes, spheres=spheres, cylinders=cyls)
            refB = ref.RefScene(a["X"], a["Tet"], a["mass"], a["mu"], planes=planes, spheres=spheres, cylinders=cyls)
            e0 = pd.PdSolver(sc, rot_mode=0); e1 = pd.PdSolver(sc, rot_mode=1, reorder=0)   # eng1 = faithful mode: reference SVD, input tet order
            osc, _ = meshes.oracle_scene(O, assets, ctx)
            osc64 = meshes.oracle_scene(O, assets, ctx)[0] if with_f64 else None
            op = O.make_params(dt=p["dt"], gravity=p["gravity"], muN=p["muN"], muT=p["muT"], rho=p["rho"],
                               num_iterations=p["num_iterations"], threads=os.cpu_count() or 1)
            rows = []
            print(f"== {ctx}: nV {a['X'].shape[0]} (compared: first {nA}), scale {scale:.3f}")
            print(f"{'step':>5} {'refA-refB':>10} {'eng0-refA':>10} {'eng1-refA':>10} {'orc-refA':>10} {'eng1-orc':>10} {'orc64-refA':>10} {'min y':>9}")
            for s in range(steps // 10):
                refA.step(10, **kw); refB.step(10, **kw); e0.Update(10); e1.Update(10); osc.step(op, 10)
                if osc64 is not None:
                    osc64.step(op, 10, f64=True)
                XA = refA.get()[2][:nA]; XB = refB.get()[2][:nA]; X0 = e0.download()[2][:nA]; X1 = e1.download()[2][:nA]
                Xo = osc.get()[2][:nA]
                X64 = osc64.get()[2][:nA] if osc64 is not None else None
                r = dict(step=10 * (s + 1), ref_ref=meshes.rel_err(XB, XA, scale), eng0_ref=meshes.rel_err(X0, XA, scale),
                         eng1_ref=meshes.rel_err(X1, XA, scale), orc_ref=meshes.rel_err(Xo, XA, scale), eng1_orc=meshes.rel_err(X1, Xo, scale),
                         orc64_ref=meshes.rel_err(X64, XA, scale) if X64 is not None else None, min_y=float(XA[:, 1].min()))
                rows.append(r)
                f = lambda v: "      n/a " if v is None else f"{v:10.2e}"
                print(f"{r['step']:5d} {f(r['ref_ref'])} {f(r['eng0_ref'])} {f(r['eng1_ref'])} {f(r['orc_ref'])} {f(r['eng1_orc'])} {f(r['orc64_ref'])} {r['min_y']:9.3f}")
            out[ctx] = dict(scale=scale, rows=rows)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "noise_floor.json"), "w"), indent=1)

"""Test scenes: re-emits the mesh fixture (tests/golden/meshes.npz) as TetGen text files laid
out like the reference's assets/ directory, plus a context.json in the reference's schema
(context.json:1-246) whose float contexts are the bench/parity configs of SURVEY.md section 8d."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_NPZ = None


def npz():
    global _NPZ
    if _NPZ is None:
        _NPZ = np.load(os.path.join(HERE, "golden", "meshes.npz"))
    return _NPZ


def raw_mesh(name):
    z = npz()
    return z[name + "_X"], z[name + "_ele"], int(z[name + "_node_idx0"])


SOFT_DEFS = [  # the shipped definitions, context.json:20-108
    dict(name="house1", nodeFile="../assets/house2/house2.node", eleFile="../assets/house2/house2.ele", centralize=True,
         **{"start index": 1}, pos=[0.0, 40.0, 0.0], rot=[0.0, 0.0, 0.0], scale=[4, 4, 4], mass=10.0, mu=200000.0, DBC=[], **{"lambda": 5000.0}),
    dict(name="armadillo0", nodeFile="../assets/armadillo0/armadillo0.1.node", eleFile="../assets/armadillo0/armadillo0.1.ele",
         centralize=False, **{"start index": 0}, pos=[0.0, 150.0, 0.0], rot=[0.0, 0.0, 0.0], scale=[0.2, 0.2, 0.2], mass=1.0,
         mu=2000000.0, DBC=[], **{"lambda": 5000.0}),
    dict(name="bunny", nodeFile="../assets/bunny/bunny.1.node", eleFile="../assets/bunny/bunny.1.ele", mass=10.0, mu=2000000.0,
         DBC=[], pos=[0.0, 25.0, 25.0], rot=[0.0, 180.0, 0.0], scale=[35, 35, 35], centralize=False, **{"start index": 1, "lambda": 5000.0}),
    dict(name="softSphere", nodeFile="../assets/sphere/sphere.1.node", eleFile="../assets/sphere/sphere.1.ele", mass=10.0, mu=200000.0,
         DBC=[], pos=[0.0, 10.0, 25.0], rot=[0.0, 0.0, 0.0], scale=[20, 20, 20], centralize=False, **{"start index": 1, "lambda": 5000.0}),
    dict(name="cube", nodeFile="../assets/cube/cube.1.node", eleFile="../assets/cube/cube.1.ele", mass=10.0, mu=200000.0, DBC=[],
         pos=[0.0, 10.0, 25.0], rot=[0.0, 0.0, 0.0], scale=[5, 5, 5], centralize=False, **{"start index": 1, "lambda": 5000.0}),
    dict(name="tet", nodeFile="../assets/tet/tet.node", eleFile="../assets/tet/tet.ele", mass=10.0, mu=20000.0, DBC=[],
         pos=[0.0, 10.0, 25.0], rot=[0.0, 0.0, 0.0], scale=[4, 4, 4], centralize=False, **{"start index": 1, "lambda": 5000.0}),
]
FIXED_DEFS = [  # context.json:109-118
    dict(type="sphere", name="sphere1", pos=[0.0, -50.0, 0.0], radius=10.0),
    dict(type="cylinder", name="cylinder1", pos=[0.0, 10.0, 0.0], rot=[90.0, 0.0, 0.0], scale=[10.0, 75.0, 10.0]),
    dict(type="plane", name="bottom plane", pos=[0.0, 0.0, 0.0], scale=[450.0, 450.0, 450.0]),
    dict(type="plane", name="upper plane", pos=[0.0, 450.0, 0.0], scale=[450.0, 450.0, 450.0], rot=[0.0, 0.0, 180.0]),
    dict(type="plane", name="right plane", pos=[450.0, 0.0, 0.0], scale=[450.0, 450.0, 450.0], rot=[0.0, 0.0, 90.0]),
    dict(type="plane", name="left plane", pos=[-450.0, 0.0, 0.0], scale=[450.0, 450.0, 450.0], rot=[0.0, 0.0, -90.0]),
    dict(type="plane", name="front plane", pos=[0.0, 0.0, -300.0], scale=[450.0, 450.0, 450.0], rot=[90.0, 0.0, 0.0]),
    dict(type="plane", name="back plane", pos=[0.0, 0.0, 450.0], scale=[450.0, 450.0, 450.0], rot=[-90.0, 0.0, 0.0]),
]
WALLS = [{"name": n} for n in ("upper plane", "right plane", "left plane", "front plane", "back plane")]
CONTEXTS = [
    # C1: SURVEY 8d -- cube dropped on the floor, dt is overridden to float(1/60) by the tests
    dict(name="C1 cube", precision="float", load=True, dt=0.016666668, gravity=98, damp=0.999, muN=0.5, muT=0.5, tolerance=1e-2,
         softBodies=[dict(name="cube", pos=[0.0, 30.0, 0.0], rot=[0.0, 0.0, 0.0])],
         fixedBodies=[dict(name="bottom plane", pos=[0.0, 0.0, 0.0])]),
    # C2: armadillo0 + bunny, parameters of the shipped "Armadillo&house" context (context.json:121-129)
    dict(name="C2 armadillo&bunny", precision="float", load=True, dt=0.01, gravity=98, damp=0.999, muN=0.5, muT=0.5, tolerance=1e-2,
         softBodies=[dict(name="armadillo0", pos=[2.0, 80.0, 0.0]), dict(name="bunny", pos=[60.0, 40.0, 0.0])],
         fixedBodies=[dict(name="bottom plane", pos=[0.0, 0.0, 0.0])] + WALLS),
    dict(name="C2 armadillo", precision="float", load=True, dt=0.01, gravity=98, damp=0.999, muN=0.5, muT=0.5, tolerance=1e-2,
         softBodies=[dict(name="armadillo0", pos=[2.0, 80.0, 0.0])], fixedBodies=[dict(name="bottom plane", pos=[0.0, 0.0, 0.0])] + WALLS),
    # the shipped float context, all three fixed-body kinds (context.json:120-151), mesh collision off
    dict(name="Armadillo&house", precision="float", load=True, dt=0.01, gravity=98, damp=0.999, muN=0.5, muT=0.5, tolerance=1e-2,
         softBodies=[dict(name="armadillo0", pos=[2.0, 80.0, 0.0]), dict(name="house1", pos=[0.0, 180.0, 0.0], rot=[0.0, 0.0, 75.0]),
                     dict(name="softSphere", pos=[35, 250, 0], rot=[0, 0, 0]), dict(name="cube", pos=[60, 300, 0], rot=[0, 0, 0], scale=[5, 5, 5])],
         fixedBodies=[dict(name="cylinder1", pos=[-25.0, 80.0, 0.0]), dict(name="cylinder1", pos=[25.0, 80.0, 0.0]),
                      dict(name="cylinder1", pos=[0.0, 50.0, 0.0]), dict(name="cylinder1", pos=[45.0, 50.0, 0.0]),
                      dict(name="cylinder1", pos=[-45.0, 50.0, 0.0]), dict(name="bottom plane", pos=[0.0, 0.0, 0.0])] + WALLS),
    # C5 unit: house1 + softSphere, one context of the 512-context batch
    dict(name="C5 house&sphere", precision="float", load=True, dt=0.01, gravity=98, damp=0.999, muN=0.5, muT=0.5, tolerance=1e-2,
         softBodies=[dict(name="house1", pos=[0.0, 25.0, 0.0], rot=[0.0, 0.0, 0.0]), dict(name="softSphere", pos=[0.0, 70.0, 0.0], rot=[30.0, 30.0, 30.0])],
         fixedBodies=[dict(name="bottom plane", pos=[0.0, 0.0, 0.0]), dict(name="sphere1", pos=[0.0, 5.0, 0.0], radius=6.0)] + WALLS),
    dict(name="a double context", precision="double", load=True, dt=0.05, gravity=9.8,
         softBodies=[dict(name="cube", pos=[0.0, 20.0, 0.0])], fixedBodies=[dict(name="bottom plane")]),
    dict(name="not loaded", precision="float", load=False, dt=0.02, gravity=200,
         softBodies=[dict(name="tet", pos=[0, 5, 0])], fixedBodies=[dict(name="bottom plane")]),
]


def write_tetgen(path_node, path_ele, X, E, idx0):
    with open(path_node, "w") as f:
        f.write(f"{X.shape[0]}  3  0  0\n")
        for i in range(X.shape[0]):
            f.write(f"   {idx0 + i}    {X[i, 0]:.9g}  {X[i, 1]:.9g}  {X[i, 2]:.9g}\n")
        f.write("# regenerated from tests/golden/meshes.npz\n")
    with open(path_ele, "w") as f:
        f.write(f"{E.shape[0]}  4  0\n")
        for r in E:
            f.write(f"    {r[0]}    {r[1]} {r[2]} {r[3]} {r[4]}\n")


def write_assets(root):
    """-> dict(root=..., json=path, assets=dir).  Layout mirrors the reference: <root>/context.json,
    <root>/assets/<mesh>/..., and "../assets/..." resolves from <root>/build."""
    z = npz()
    names = sorted({k[:-2] for k in z.files if k.endswith("_X")})
    for name in names:
        node, ele = [str(s) for s in z[name + "_files"]]
        os.makedirs(os.path.join(root, "assets", os.path.dirname(node)), exist_ok=True)
        X, E, idx0 = raw_mesh(name)
        write_tetgen(os.path.join(root, "assets", node), os.path.join(root, "assets", ele), X, E, idx0)
    os.makedirs(os.path.join(root, "build"), exist_ok=True)
    cfg = {"pause": False, "threads per block": 128, "threads per block(bvh)": 128, "num of iterations": 100,
           "softBodies": SOFT_DEFS, "fixedBodies": FIXED_DEFS, "contexts": CONTEXTS}
    jp = os.path.join(root, "context.json")
    with open(jp, "w") as f:
        json.dump(cfg, f, indent=2)
    return dict(root=root, json=jp, assets=os.path.join(root, "assets"))


def oracle_scene(O, assets, ctx_name):
    """Build the same scene with the oracle's own loaders/transforms (independent of the product)."""
    cfg = json.load(open(assets["json"]))
    ctx = [c for c in cfg["contexts"] if c["name"] == ctx_name][0]
    sdefs = {d["name"]: d for d in cfg["softBodies"]}
    fdefs = {d["name"]: d for d in cfg["fixedBodies"]}
    Xs, Ts, masses, mus = [], [], [], []
    off = 0
    for sb in ctx["softBodies"]:
        d = sdefs[sb["name"]]
        g = lambda k, dflt=None: sb.get(k, d.get(k, dflt))
        X = O.load_node(os.path.join(assets["root"], d["nodeFile"].replace("../", "")), d.get("centralize", False))
        T = O.load_ele(os.path.join(assets["root"], d["eleFile"].replace("../", "")), d.get("start index", 0))
        M = O.model_matrix(g("pos", [0, 0, 0]), g("rot", [0, 0, 0]), g("scale", [0, 0, 0]), 1)
        X = O.transform_vertices(X, M)
        Xs.append(X); Ts.append(T + np.uint32(off))
        masses.append(np.full(X.shape[0], g("mass", 1.0), np.float32)); mus.append(np.full(T.shape[0], g("mu", 1000.0), np.float32))
        off += X.shape[0]
    planes, spheres, cyls = [], [], []
    for fb in ctx.get("fixedBodies", []):
        d = fdefs[fb["name"]]
        g = lambda k, dflt: fb.get(k, d.get(k, dflt))
        pos, rot, scale = g("pos", [0, 0, 0]), g("rot", [0, 0, 0]), g("scale", [1, 1, 1])
        if d["type"] == "plane":
            M = O.model_matrix(pos, rot, scale, 0)
            planes.append((M[12:15].copy(), O.plane_up(M)))
        elif d["type"] == "sphere":
            r = g("radius", 1.0)
            M = O.model_matrix(pos, rot, [r, r, r], 0)
            spheres.append((M[12:15].copy(), r))
        else:
            M = O.model_matrix(pos, rot, [scale[0], scale[1], scale[0]], 0)
            cyls.append((M[12:15].copy(), O.cylinder_axis(M), scale[0]))
    X = np.concatenate(Xs); T = np.concatenate(Ts)
    scene = O.Scene(X, T, np.concatenate(masses), np.concatenate(mus), planes=planes, spheres=spheres, cylinders=cyls)
    params = dict(dt=ctx.get("dt", 0.001), gravity=ctx.get("gravity", 9.8), muN=ctx.get("muN", 0.5), muT=ctx.get("muT", 0.5),
                  num_iterations=cfg["num of iterations"])
    return scene, params


def rel_err(x, ref, scale=None):
    """SURVEY 8c tolerance plan: max_v |x_v - ref_v| / max(|ref_v|, scene_scale)."""
    x = np.asarray(x, np.float64); ref = np.asarray(ref, np.float64)
    if scale is None:
        scale = np.linalg.norm(ref.max(0) - ref.min(0))
    den = np.maximum(np.linalg.norm(ref, axis=1), scale)
    return float((np.linalg.norm(x - ref, axis=1) / den).max())

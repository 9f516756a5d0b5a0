"""Generates tests/golden/meshes.npz from the reference's TetGen assets.

Run in the authoring container (needs /root/reference):  python tests/golden/make_mesh_fixtures.py
The GPU box has no /root/reference, so the parsed meshes travel as this fixture: raw float32
coordinates exactly as the reference's loader parses them (`>> float`, dataLoader.cu:150-152, before
centralize/transform) and the five integer columns of every .ele row (dataLoader.cu:56-64).
tests/meshes.py re-emits them as .node/.ele text ("%.9g" round-trips float32 exactly).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/assets"
MESHES = {  # name: (node, ele)
    "cube": ("cube/cube.1.node", "cube/cube.1.ele"),
    "tet": ("tet/tet.node", "tet/tet.ele"),
    "house2": ("house2/house2.node", "house2/house2.ele"),
    "sphere": ("sphere/sphere.1.node", "sphere/sphere.1.ele"),
    "bunny": ("bunny/bunny.1.node", "bunny/bunny.1.ele"),
    "armadillo0": ("armadillo0/armadillo0.1.node", "armadillo0/armadillo0.1.ele"),
}


def parse_node(path):
    with open(path) as f:
        n = int(f.readline().split()[0])
        idx = np.zeros(n, np.int32); X = np.zeros((n, 3), np.float32)
        for i in range(n):
            t = f.readline().split()
            idx[i] = int(t[0]); X[i] = [np.float32(t[1]), np.float32(t[2]), np.float32(t[3])]
    return idx, X


def parse_ele(path):
    with open(path) as f:
        n = int(f.readline().split()[0])
        E = np.zeros((n, 5), np.int32)
        for i in range(n):
            E[i] = [int(x) for x in f.readline().split()[:5]]
    return E


if __name__ == "__main__":
    out = {}
    for name, (node, ele) in MESHES.items():
        idx, X = parse_node(os.path.join(REF, node))
        E = parse_ele(os.path.join(REF, ele))
        out[name + "_node_idx0"] = np.int32(idx[0])
        out[name + "_X"] = X
        out[name + "_ele"] = E
        out[name + "_files"] = np.array([node, ele])
        print(name, X.shape, E.shape)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "meshes.npz"), **out)

"""pd_run, the headless C++ host driver over the C ABI (soft-body-simulation-cuda_b200/csrc/pd_run.cpp): the reference's
frame loop (main.cpp -> Context::Update -> SimulationCUDAContext::Update, simulationContext.cpp:79-86) without the window.
CPU part: argument handling, scene loading and host-side layout (--info), loud failures.  GPU part: frames written as
TetGen .node files are exactly what the C ABI returns."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "soft-body-simulation-cuda_b200", "pd_run")


def _run(*args):
    if not os.path.exists(EXE):           # built next to the library by soft-body-simulation-cuda_b200/build.py
        import importlib
        importlib.import_module("soft-body-simulation-cuda_b200.build").build(force=True)
    return subprocess.run([EXE] + [str(a) for a in args], capture_output=True, text=True, timeout=600)


def test_info_reports_the_scene_and_the_host_side_layout(pd, assets):
    r = _run("--json", assets["json"], "--context", "C5 house&sphere", "--iterations", 40, "--solver", "pcg", "--info")
    assert r.returncode == 0, r.stderr
    info = json.loads(r.stdout)
    sc = pd.Scene.from_json(assets["json"], "C5 house&sphere")
    nV, nT, nF, nB = sc.counts()
    assert (info["num_verts"], info["num_tets"], info["num_fixed"], info["num_bodies"]) == (nV, nT, nF, nB)
    assert info["num_iterations"] == 40 and info["global_solver"] == pd.PD_PCG_JACOBI
    assert info["num_tiles"] == sc.layout().num_tiles and info["dt"] == pytest.approx(0.01, rel=1e-6)
    # first loadable context when none is named (context.cpp:319-387 keeps every context with "load": true; the first one is shown)
    r = _run("--json", assets["json"], "--info")
    assert r.returncode == 0 and json.loads(r.stdout)["num_tets"] == 6


def test_failures_are_loud(assets):
    assert _run("--steps", 3).returncode == 64                                     # --json missing
    assert _run("--json", assets["json"], "--solver", "lu", "--info").returncode == 64
    r = _run("--json", assets["json"], "--context", "no such context", "--info")
    assert r.returncode == 1 and "context not found" in r.stderr
    r = _run("--json", os.path.join(assets["root"], "missing.json"), "--info")
    assert r.returncode == 1 and r.stderr.strip()
    r = _run("--json", assets["json"], "--context", "a double context", "--steps", 1)      # IPC contexts are not PD's
    assert r.returncode == 2 and "not a float" in r.stderr


@pytest.mark.gpu
def test_frames_are_what_the_c_abi_returns(pd, assets, tmp_path):
    out = str(tmp_path / "cube")
    r = _run("--json", assets["json"], "--context", "C1 cube", "--steps", 12, "--iterations", 30, "--out", out, "--every", 6, "--perf",
             "--drag", "3,0.5,31,0.25,4,9")
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout)
    assert line["finite"] and line["steps"] == 12 and line["pd_iterations"] == 12 * 30 and line["perf_ms"]["local step"] > 0
    frames = sorted(f for f in os.listdir(tmp_path) if f.endswith(".node"))
    assert frames == ["cube.00000.node", "cube.00006.node", "cube.00012.node"]
    sc = pd.Scene.from_json(assets["json"], "C1 cube")
    p = sc.params
    p["num_iterations"] = 30
    sc.params = p
    eng = pd.PdSolver(sc)
    eng.SetPerf(True)
    target = np.float32([0.5, 31.0, 0.25])

    def frame(name):
        return np.loadtxt(os.path.join(tmp_path, name), skiprows=1, dtype=np.float64)[:, 1:].astype(np.float32)

    assert np.array_equal(frame("cube.00000.node"), eng.download()[0])
    eng.Update(4)
    eng.drag_select(3, target)
    eng.Update(2)
    assert np.array_equal(frame("cube.00006.node"), eng.download()[0]) and np.array_equal(eng.download()[0][3], target)
    eng.Update(3)
    eng.set_drag(None)
    eng.Update(3)
    assert np.array_equal(frame("cube.00012.node"), eng.download()[0])

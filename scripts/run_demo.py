#!/usr/bin/env python
"""Run one of the reference's demos UNCHANGED on this engine:

    python scripts/run_demo.py /root/reference/python-implementation/visualization.py
    python scripts/run_demo.py /root/reference/python-implementation/robot-visualization.py --headless --frames 60

Both demos do ``from gicp import gicp, apply_transformation`` (visualization.py:7, robot-visualization.py:6) and
Python would resolve that to the gicp.py lying beside the script.  The launcher leaves the script's bytes alone and
controls the resolution instead: it imports THIS repo's drop-in module, registers it as ``sys.modules['gicp']`` and
executes the script with ``runpy.run_path`` (which does not put the script's directory on sys.path).  The robot demo
forks its GICP worker after that (robot-visualization.py:199), so the worker inherits the module and makes the
first CUDA call itself - the engine initialises lazily inside gicp().  ``--headless`` puts a no-op pygame
(scripts/headless_pygame) first on sys.path: no display, a scripted event queue that quits after ``--frames`` frames.
"""
import argparse
import multiprocessing as mp
import os
import runpy
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(script, frames=30, headless=True, gicp_module=None, seed=None, tick=0.05):
    """Executes `script` as __main__ with `gicp_module` (default: this repo's drop-in) standing in for the
    reference's gicp.py.  Returns {"gicp_calls": completed registrations (counted across forked workers),
    "exit": the script's exit code, "pygame_calls": {...} when headless}."""
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    if gicp_module is None:
        import gicp as gicp_module                      # the repo-root shim: nothing CUDA-related runs at import
    calls = mp.get_context("fork").Value("i", 0)        # shared with the demo's forked worker
    real = gicp_module.gicp

    def counted(*a, **k):
        out = real(*a, **k)
        with calls.get_lock():
            calls.value += 1
        return out

    shim = types.ModuleType("gicp")
    shim.gicp = counted
    shim.apply_transformation = gicp_module.apply_transformation
    saved = {k: sys.modules.get(k) for k in ("gicp", "pygame", "pygame.locals")}
    sys.modules["gicp"] = shim
    stub_dir = os.path.join(ROOT, "scripts", "headless_pygame")
    if headless:
        os.environ["PYGAME_STUB_FRAMES"] = str(frames)
        os.environ["PYGAME_STUB_TICK"] = str(tick)
        for k in ("pygame", "pygame.locals"):
            sys.modules.pop(k, None)
        sys.path.insert(0, stub_dir)
    if seed is not None:
        import random
        import numpy as np
        random.seed(seed)
        np.random.seed(seed)
    code = 0
    try:
        runpy.run_path(script, run_name="__main__")
    except SystemExit as ex:                            # robot-visualization.py ends with sys.exit()
        code = ex.code or 0
    finally:
        pg = sys.modules.get("pygame")
        stats = dict(getattr(pg, "_state", {}).get("calls", {})) if headless and pg else None
        if headless and stub_dir in sys.path:
            sys.path.remove(stub_dir)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return {"gicp_calls": calls.value, "exit": code, "pygame_calls": stats}


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("script")
    ap.add_argument("--headless", action="store_true", help="no display: use the no-op pygame of scripts/headless_pygame")
    ap.add_argument("--frames", type=int, default=60, help="headless: frames before the scripted QUIT")
    ap.add_argument("--seed", type=int, default=None, help="seed random / numpy.random first (the demos are unseeded)")
    a = ap.parse_args()
    r = run(a.script, frames=a.frames, headless=a.headless, seed=a.seed)
    print(f"[run_demo] exit {r['exit']}, {r['gicp_calls']} registration(s) completed")
    sys.exit(r["exit"])


if __name__ == "__main__":
    main()

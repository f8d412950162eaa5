"""Loader for the committed golden fixtures (tests/golden/, made by make_golden.py
from the unmodified reference).  Pure NumPy; no access to /root/reference."""
from __future__ import annotations

import functools
import hashlib
import json
import os
from typing import Dict, Tuple

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@functools.lru_cache(None)
def manifest() -> dict:
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        return json.load(f)


@functools.lru_cache(None)
def _npz(name: str):
    return np.load(os.path.join(GOLDEN, name))


def stage(key: str) -> np.ndarray:
    return _npz("stage_vectors.npz")[key]


def bundle(name: str) -> Tuple[np.ndarray, Dict[int, np.ndarray]]:
    z = _npz("bundles.npz")
    objs = {}
    for k in z.files:
        if k.startswith(name + "/obj"):
            objs[int(k.split("obj")[1])] = z[k]
    return z[name + "/background"], objs


def case_names():
    return [c["name"] for c in manifest()["cases"]]


def case(name: str):
    """-> (background array, {id: cutout array}, placements, expected output array)."""
    c = next(c for c in manifest()["cases"] if c["name"] == name)
    z = _npz("composites.npz")
    expected = z["case/" + name]
    assert sha(expected) == c["sha256"]
    W, H = c["canvas"]
    if c.get("synthetic") == "known_answer":
        bg = np.zeros((10, 10, 4), np.uint8)
        bg[...] = (255, 0, 0, 255)
        obj = np.zeros((2, 2, 4), np.uint8)
        obj[...] = (0, 255, 0, 255)
        return bg, {1: obj}, c["placements"], expected
    if c.get("synthetic") == "stored_inputs":
        bg = z[f"in/{name}/bg"]
        objs = {}
        for k in z.files:
            if k.startswith(f"in/{name}/obj"):
                objs[int(k.split("obj")[1])] = z[k]
        return bg, objs, c["placements"], expected
    bundle_bg, objs = bundle(c["bundle"])
    if c["bg"] == "fill_solid":
        col = manifest()["bundles"][c["bundle"]]["median_color"]
        bg = np.zeros((H, W, 4), np.uint8)
        bg[...] = tuple(col) + (255,)
    else:
        bg = bundle_bg
    return bg, objs, c["placements"], expected

"""Reader for the RFW1 weight blob (weights/rover_fe.rfw) -- numpy side.

Mirrors rover_slam_b200/csrc/weights.h.  Used by the oracle restatements and by tests.
"""
from __future__ import annotations

import os
import struct

import numpy as np

DEFAULT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "weights", "rover_fe.rfw")


def load(path: str = DEFAULT) -> dict:
    with open(path, "rb") as f:
        buf = f.read()
    magic, n, total = struct.unpack_from("<4sIQ", buf, 0)
    if magic != b"RFW1":
        raise ValueError("not an RFW1 blob")
    out = {}
    for i in range(n):
        name, nd, d0, d1, d2, d3, off, nb = struct.unpack_from("<80sI4IQQ", buf, 16 + 128 * i)
        name = name.rstrip(b"\0").decode()
        dims = [d0, d1, d2, d3][:nd]
        out[name] = np.frombuffer(buf, dtype="<f4", count=nb // 4, offset=off).reshape(dims).copy()
    return out

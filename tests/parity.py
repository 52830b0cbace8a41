"""The RGBA parity metric shared by every image comparison in tests/, smoke() and bench.py.

BASELINE.json north_star: "RGBA within 1e-4 of the CPU reference" (relative).  A purely relative
error is ill-defined where the image is (nearly) empty: coverage alpha is 1 - T with T a product of
~1e3 blend factors, so a pixel with alpha ~ 1e-6 carries the fp32 cancellation error of T (6e-8) as
a relative error of several percent in ANY two evaluation orders.  The metric therefore is

    |engine - oracle| <= RTOL * max(|oracle|, ABS_FLOOR),   RTOL = 1e-4, ABS_FLOOR = 1e-2

i.e. 1e-4 relative for every value >= 0.01 and 1e-6 absolute below that (the reference's own
output target is 8-bit, one LSB = 3.9e-3, VPR.cs:228).
"""
import numpy as np

RTOL = 1e-4
ABS_FLOOR = 1e-2


def rel_err(engine, oracle):
    return np.abs(engine - oracle) / np.maximum(np.abs(oracle), ABS_FLOOR)


def max_rel_err(engine, oracle):
    return float(rel_err(engine, oracle).max()) if engine.size else 0.0

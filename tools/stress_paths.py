#!/usr/bin/env python3
"""Randomised agreement check of the planned paths (clustered, general clustered, repeated-segment) against
the single-interval Chebyshev path: large seeded batches in the parameter ranges of SURVEY.md section 8d,
3- and 4-level, noisy and noiseless, identical and unequal atoms.  Prints one line per batch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np  # noqa: E402

import _cases as C  # noqa: E402
from noisyquantumsimulator_b200.engine import Evaluator  # noqa: E402
from oracle import rydcz_oracle as orc  # noqa: E402

ev = Evaluator(0)
rng = np.random.default_rng(20261018)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000


def batch(protocol, n, unequal):
    P = np.zeros((n, 12)); segs = []
    Om = 2 * np.pi * rng.uniform(1.5e6, 9e6, n)
    P[:, 0] = Om * 10 ** rng.uniform(0.6, 2.9, n)               # V / Omega from 4 to 800
    P[:, 1] = 2 * np.pi * rng.uniform(0, 3e5, n)
    P[:, 4] = rng.uniform(2e3, 5e3, n); P[:, 5] = rng.uniform(2e3, 9e3, n)
    P[:, 6] = 10 ** rng.uniform(2, 6, n); P[:, 7] = 10 ** rng.uniform(3, 5.5, n); P[:, 8] = rng.uniform(0, 5e3, n)
    if unequal:
        P[:, 2] = 2 * np.pi * 5e4 * rng.standard_normal(n); P[:, 3] = 2 * np.pi * 5e4 * rng.standard_normal(n)
    S = {"lp": 2, "bb": 5, "bb7": 7, "sjp": 300}[protocol]
    G = np.zeros((n, S, 4))
    if protocol == "lp":
        tau = rng.uniform(3.5, 5.5, n) / Om
        xi = np.exp(1j * rng.uniform(-np.pi, np.pi, n))
        det = rng.uniform(0.2, 0.5, n) * Om
        G[:, :, 0] = tau[:, None]; G[:, 0, 1] = Om; G[:, 1, 1] = (Om * xi).real; G[:, 1, 2] = (Om * xi).imag
        G[:, :, 3] = det[:, None]
    elif protocol in ("bb", "bb7"):
        ot = rng.uniform(5, 40, n)
        b = np.concatenate([np.zeros((n, 1)), np.sort(rng.uniform(0.01, 0.99, (n, S - 1)), 1), np.ones((n, 1))], 1)
        ph = rng.uniform(-np.pi, np.pi, (n, S))
        G[:, :, 0] = np.diff(b, axis=1) * (ot / Om)[:, None]
        G[:, :, 1] = Om[:, None] * np.cos(ph); G[:, :, 2] = Om[:, None] * np.sin(ph)
    else:
        T = rng.uniform(5, 25, n) / Om
        A, wm, ph0 = rng.uniform(0.05, 1, n) * np.pi, rng.uniform(0.5, 3, n) * Om, rng.uniform(0, 2 * np.pi, n)
        t_mid = (np.arange(S) + 0.5)[None, :] * (T / S)[:, None]
        ph = A[:, None] * np.cos(wm[:, None] * t_mid - ph0[:, None])
        G[:, :, 0] = (T / S)[:, None]
        G[:, :, 1] = Om[:, None] * np.cos(ph); G[:, :, 2] = Om[:, None] * np.sin(ph)
        G[:, :, 3] = (-rng.uniform(0.001, 0.1, n) * Om)[:, None]
    return P, G, np.full(n, S, dtype=np.int32)


worst = 0.0
for dim in (3, 4):
    for protocol, n in (("lp", N), ("bb", N), ("bb7", N), ("sjp", max(N // 10, 256))):
        for noisy in (True, False):
            for unequal in (False, True):
                P, G, ns = batch(protocol, n, unequal)
                ev.set_cluster(0)
                a, sa = ev.evaluate_host(dim, noisy, P, G, ns, want_states=True)
                ev.cluster_stats(); ev.repeat_stats()
                ev.set_cluster(1)
                b, sb = ev.evaluate_host(dim, noisy, P, G, ns, want_states=True)
                planned, back = ev.cluster_stats()
                rep = ev.repeat_stats()
                d_out, d_st = float(np.abs(a[:, :14] - b[:, :14]).max()), float(np.abs(sa - sb).max())
                worst = max(worst, d_out, d_st)
                print(f"dim {dim} {protocol:3s} noisy {int(noisy)} unequal {int(unequal)} n {n:6d}: clustered {planned:6d} "
                      f"repeated {rep:6d} handed back {back:3d}  max|d out| {d_out:.2e}  max|d state| {d_st:.2e}  "
                      f"operator applications {a[:, 14].sum() / b[:, 14].sum():5.1f}x fewer", flush=True)
print("worst difference", worst)

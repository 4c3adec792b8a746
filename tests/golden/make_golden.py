"""Generates the golden vectors under tests/golden/*.npz with the CPU oracle.

The reference's tests pin nothing (test/generative.jl:59-67), and Julia cannot run here, so
these vectors are ORACLE-generated: they pin the oracle against regressions and give the GPU
tests fixed inputs/outputs that do not depend on numpy's generator at test time.  Streams are
stored already guard-banded (see qarbom_oracle.GuardedStream).

    python tests/golden/make_golden.py        # rewrites the fixtures (deterministic)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import qarbom_oracle as qo  # noqa: E402
from tests.helpers import DenseStreams, chains_from_arrays, chains_to_arrays, f32_round, make_oracle_model, u24  # noqa: E402

CASES = [  # name, V, H, L, gaussian, algo, B, k, steps
    ("pcd_rbm_3x2", 3, 2, 0, False, "pcd", 2, 1, 4),              # test/generative.jl:22-38 shape
    ("cd_rbm_3x2_online", 3, 2, 0, False, "cd", 1, 3, 6),         # test/generative.jl:4-20 shape
    ("pcd_clf_3x2x1", 3, 2, 1, False, "pcd", 2, 1, 4),            # test/classification.jl shape
    ("cd_clf_3x2x1_online", 3, 2, 1, False, "cd", 1, 3, 6),
    ("pcd_rbm_16x8", 16, 8, 0, False, "pcd", 8, 2, 3),
    ("cd_rbm_64x32_b16", 64, 32, 0, False, "cd", 16, 2, 2),
    ("pcd_grbm_24x16", 24, 16, 0, True, "pcd", 8, 2, 2),
    ("cd_clf_50x20x10_b16", 50, 20, 10, False, "cd", 16, 1, 2),
]


def generate(name, V, H, L, gaussian, algo, B, k, steps, seed):
    rng = np.random.default_rng(seed)
    m = make_oracle_model(rng, V, H, L, gaussian)
    X = f32_round(rng.standard_normal((steps * B, V))) if gaussian else rng.integers(0, 2, (steps * B, V)).astype(np.float64)
    Y = np.eye(L)[rng.integers(0, L, steps * B)] if L else None
    out = dict(V=V, H=H, L=L, gaussian=gaussian, algo=algo, B=B, k=k, steps=steps, lr=0.05, llr=0.02,
               W0=m.W.copy(), a0=m.a.copy(), b0=m.b.copy(), X=X)
    if L:
        out.update(U0=m.U.copy(), c0=m.c.copy(), Y=Y)
    chains = None
    if algo == "pcd":
        cv0, ch0 = u24(rng, B, V), u24(rng, B, H)
        cy0 = u24(rng, B, L) if L else None
        chains = chains_from_arrays(cv0, ch0, cy0)
        out.update(cv0=cv0, ch0=ch0)
        if L:
            out["cy0"] = cy0
    for t in range(steps):
        ds = DenseStreams(rng, k, B, V, H, L, gaussian)
        su, sz = ds.serial(algo)
        st = qo.GuardedStream(su, sz)
        sl = slice(t * B, (t + 1) * B)
        if algo == "pcd":
            qo.persistent_contrastive_divergence(m, X, [range(t * B, (t + 1) * B)], chains, steps=k, learning_rate=0.05,
                                                 stream=st, y=Y, label_learning_rate=0.02)
            cv, ch, cy = chains_to_arrays(chains)
            out[f"cv_{t}"], out[f"ch_{t}"] = cv, ch
            if L:
                out[f"cy_{t}"] = cy
        else:
            qo.contrastive_divergence(m, X[sl], steps=k, learning_rate=0.05, stream=st, y=None if Y is None else Y[sl],
                                      label_learning_rate=0.02, batch_size=B)
        ds.absorb(st.u, algo)
        out[f"Uh_{t}"] = ds.U_h
        if ds.U_v is not None:
            out[f"Uv_{t}"] = ds.U_v
        if ds.U_y is not None:
            out[f"Uy_{t}"] = ds.U_y
        if ds.Z_v is not None:
            out[f"Zv_{t}"] = ds.Z_v
        out[f"W_{t}"], out[f"a_{t}"], out[f"b_{t}"] = m.W.copy(), m.a.copy(), m.b.copy()
        if L:
            out[f"U_{t}"], out[f"c_{t}"] = m.U.copy(), m.c.copy()
    out["mse_final"] = qo.evaluate_mse(m, X, qo.Stream(normals=np.zeros(X.size))) if not L else 0.0
    out["free_energy_final"] = qo.mean_free_energy(m, X, Y)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


if __name__ == "__main__":
    for i, c in enumerate(CASES):
        generate(*c, seed=1000 + i)
        print("wrote", c[0])

"""Golden vectors for one FastPCD epoch (tests/golden/fastpcd/*.npz), generated with the CPU oracle
(`fast_persistent_contrastive_divergence`, src/training/train_fast_pcd.jl:1-127).  They pin the restatement --
including the classifier's chain-1 pairing and the thresholded Gaussian chain step -- against regressions.

    python tests/golden/make_golden_fastpcd.py        # rewrites the fixtures (deterministic)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import qarbom_oracle as qo  # noqa: E402
from tests.helpers import chains_from_arrays, chains_to_arrays, f32_round, make_oracle_model, u24  # noqa: E402

CASES = [  # name, V, H, L, gaussian, B, k, n_batches
    ("fastpcd_rbm_12x8", 12, 8, 0, False, 4, 2, 3),
    ("fastpcd_clf_10x6x3", 10, 6, 3, False, 4, 1, 3),
    ("fastpcd_grbm_9x5", 9, 5, 0, True, 4, 2, 2),
    ("fastpcd_grbmclf_8x5x2", 8, 5, 2, True, 4, 1, 2),
]


def generate(name, V, H, L, gaussian, B, k, nb, seed):
    rng = np.random.default_rng(seed)
    m = make_oracle_model(rng, V, H, L, gaussian)
    n = nb * B
    X = f32_round(rng.standard_normal((n, V))) if gaussian else rng.integers(0, 2, (n, V)).astype(np.float64)
    Y = np.eye(L)[rng.integers(0, L, n)] if L else None
    cv0, ch0 = u24(rng, B, V), u24(rng, B, H)
    cy0 = u24(rng, B, L) if L else None
    # serial streams: per mini-batch, per Gibbs step, per chain: H uniforms, (V normals,) V uniforms, L uniforms
    U = u24(rng, nb * k * B * (H + V + L))
    Z = f32_round(rng.standard_normal(nb * k * B * V)) if gaussian else None
    chains = chains_from_arrays(cv0, ch0, cy0)
    init = qo.copy_rbm(m)
    st = qo.GuardedStream(U, Z)
    qo.fast_persistent_contrastive_divergence(m, X, qo._set_mini_batches(n, B), chains, steps=k, learning_rate=0.05,
                                              fast_learning_rate=0.1, stream=st, y=Y, label_learning_rate=0.04,
                                              fast_label_learning_rate=0.08)
    assert st.upos == U.size and (Z is None or st.zpos == Z.size)
    cv, ch, cy = chains_to_arrays(chains)
    out = dict(V=V, H=H, L=L, gaussian=gaussian, B=B, k=k, nb=nb, X=X, cv0=cv0, ch0=ch0, U=st.u, cv=cv, ch=ch,
               W0=init.W, a0=init.a, b0=init.b, W=m.W, a=m.a, b=m.b)
    if L:
        out.update(Y=Y, cy0=cy0, cy=cy, U0=init.U, c0=init.c, U1=m.U, c=m.c)
    if gaussian:
        out["Z"] = Z
    np.savez_compressed(os.path.join(HERE, "fastpcd", name + ".npz"), **out)


if __name__ == "__main__":
    for i, c in enumerate(CASES):
        generate(*c, seed=2000 + i)
        print("wrote", c[0])

"""The 4-CTA-cluster chain kernel (csrc/sm100_chain.cuh, gibbs_chain_kernel_quad: two CTA pairs per 256 x 256 tile, weights
TMA-multicast between the pairs, dependencies per 128-row block) against one launch per half-step -- the path
test_gpu_parity.py / test_gpu_baseline_shapes.py pin to the oracle.  Same GEMMs, same epilogues, same Philox numbers: chain
states must be bit-identical, parameters equal up to the order of the float atomics of the bias sums.  Run with -m gpu."""
import numpy as np
import pytest

from tests.helpers import bf16_round, f32_round

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def q():
    import qarbom_b200
    qarbom_b200.build()
    return qarbom_b200


CASES = [  # V, H, L, gaussian, B, k, algo
    (1024, 1024, 0, False, 512, 3, "PCD"),     # 4 unit tiles x 2 quad row blocks per half-step
    (784, 1000, 0, False, 256, 2, "PCD"),      # ragged unit counts: partial last unit tile, K not a multiple of 64
    (1024, 2048, 0, True, 512, 2, "CD"),       # Gaussian visibles (cfg3 model), CD: 2k + 1 forward contractions in one program
    (784, 1000, 10, False, 512, 1, "CD"),      # label units inside the last visible tile (cfg4 model)
    (4096, 4096, 0, False, 1024, 2, "PCD"),    # the 8-GPU per-rank shape of cfg5: 64 quad tiles per half-step
]


@pytest.mark.parametrize("V,H,L,gaussian,B,k,algo", CASES)
def test_quad_chain_equals_one_launch_per_half_step(q, V, H, L, gaussian, B, k, algo):
    rng = np.random.default_rng(V + H + B + k)
    nb, steps = 2, 3
    X = rng.standard_normal((nb * B, V)).astype(np.float32) if gaussian else (rng.random((nb * B, V)) < 0.4).astype(np.uint8)
    Y = np.eye(L, dtype=np.uint8)[rng.integers(0, L, nb * B)] if L else None
    out = {}
    for quad in (0, 2):
        r = np.random.default_rng(11)
        W = bf16_round(0.03 * r.standard_normal((V, H)))
        if L:
            rbm = q.RBMClassifier(V, H, L, W=W, U=bf16_round(0.03 * r.standard_normal((L, H))))
        else:
            rbm = (q.GRBM if gaussian else q.RBM)(V, H, W=W)
        rbm.a[:] = f32_round(0.03 * r.standard_normal(V)); rbm.b[:] = f32_round(0.03 * r.standard_normal(H))
        with q.DeviceRBM(rbm, max_batch=B, precision="bf16", seed=0xC4A1) as dev:
            dev.set_option("fused_chain", 2 if quad else 0)
            dev.set_option("chain_quad", quad)
            dev.set_data(X, Y)
            if algo == "PCD":
                dev.init_chains(B)
            dev.reset_counters()
            for t in range(steps):
                (dev.pcd_step if algo == "PCD" else dev.cd_step)(t % nb, B, k, 0.02, 0.02)
            hist = dev.gemm_histogram()
            if quad:
                assert any(key == ("2cta", 128, "chain") for key in hist), hist
            else:
                assert not any(key[2] == "chain" for key in hist), hist
            chains = dev.get_chains() if algo == "PCD" else None
            dev.pull_params()
            out[quad] = (rbm.W.copy(), rbm.a.copy(), rbm.b.copy(), chains)
    ref, got = out[0], out[2]
    if algo == "PCD":
        assert np.array_equal(ref[3][0], got[3][0]) and np.array_equal(ref[3][1], got[3][1]), "chain states differ"
    for a, b in zip(ref[:3], got[:3]):
        assert np.max(np.abs(a - b)) <= 1e-5 * max(1.0, float(np.max(np.abs(a)))), float(np.max(np.abs(a - b)))


WIDTH_CASES = [  # V, H, L, gaussian, B, k, algo, chain_bn (0 = automatic), expected width
    (4096, 4096, 0, False, 1024, 2, "PCD", 0, 128),    # one wave of 256 x 256 pair tiles or less -> automatic choice is 256 x 128
    (4096, 4096, 0, False, 2048, 1, "PCD", 0, 256),    # more than one wave -> 256 x 256
    (1024, 1024, 0, False, 512, 3, "PCD", 64, 64),
    (784, 1000, 0, False, 256, 2, "PCD", 128, 128),    # ragged unit counts
    (1024, 2048, 0, True, 512, 2, "CD", 128, 128),     # Gaussian visibles, CD
    (784, 1000, 10, False, 512, 1, "CD", 256, 256),    # label units
]


@pytest.mark.parametrize("V,H,L,gaussian,B,k,algo,chain_bn,width", WIDTH_CASES)
def test_pair_chain_at_every_tile_width_equals_one_launch_per_half_step(q, V, H, L, gaussian, B, k, algo, chain_bn, width):
    """gibbs_chain_kernel<64|128|256> (chain_tile_width in qb_api.cu picks 128 up to one wave of 256 x 256 tiles) against one
    launch per half-step: bit-identical chain states, parameters equal up to the order of the float atomics."""
    rng = np.random.default_rng(V + H + B + k + 1)
    nb, steps = 2, 2
    X = rng.standard_normal((nb * B, V)).astype(np.float32) if gaussian else (rng.random((nb * B, V)) < 0.4).astype(np.uint8)
    Y = np.eye(L, dtype=np.uint8)[rng.integers(0, L, nb * B)] if L else None
    out = {}
    for fused in (0, 2):
        r = np.random.default_rng(12)
        W = bf16_round(0.03 * r.standard_normal((V, H)))
        if L:
            rbm = q.RBMClassifier(V, H, L, W=W, U=bf16_round(0.03 * r.standard_normal((L, H))))
        else:
            rbm = (q.GRBM if gaussian else q.RBM)(V, H, W=W)
        rbm.a[:] = f32_round(0.03 * r.standard_normal(V)); rbm.b[:] = f32_round(0.03 * r.standard_normal(H))
        with q.DeviceRBM(rbm, max_batch=B, precision="bf16", seed=0xC4A2) as dev:
            dev.set_option("fused_chain", fused)
            dev.set_option("chain_bn", chain_bn)
            dev.set_data(X, Y)
            if algo == "PCD":
                dev.init_chains(B)
            dev.reset_counters()
            for t in range(steps):
                (dev.pcd_step if algo == "PCD" else dev.cd_step)(t % nb, B, k, 0.02, 0.02)
            hist = dev.gemm_histogram()
            if fused:
                assert any(key == ("2cta", width, "chain") for key in hist), hist
            else:
                assert not any(key[2] == "chain" for key in hist), hist
            chains = dev.get_chains() if algo == "PCD" else None
            dev.pull_params()
            out[fused] = (rbm.W.copy(), rbm.a.copy(), rbm.b.copy(), chains)
    ref, got = out[0], out[2]
    if algo == "PCD":
        assert np.array_equal(ref[3][0], got[3][0]) and np.array_equal(ref[3][1], got[3][1]), "chain states differ"
    for a, b in zip(ref[:3], got[:3]):
        assert np.max(np.abs(a - b)) <= 1e-5 * max(1.0, float(np.max(np.abs(a)))), float(np.max(np.abs(a - b)))

"""CPU oracle for the classical RBM training hot path of NITeQ/QARBoM.jl.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker / the CPU baseline.

PARITY UNPINNED: Julia is not installed in this image and the reference's own tests
(`test/generative.jl:59-67`, `test/classification.jl:69-77`) contain no assertions, seeds or
golden vectors, so this restatement cannot be checked against outputs of the reference
itself.  It is a line-by-line Float64, one-sample-at-a-time restatement of the files
cited on every function below (paths relative to /root/reference), with the task-local
random stream replaced by an *injected* stream consumed in the reference's draw order.

Third-party arithmetic restated from its published algorithm (versions unpinned by the
reference: `Project.toml:16-17` only pins `julia = "1.9"`, no Manifest):
  * LogExpFunctions.jl ``logistic``: ``e = exp(x); e / (1 + e)`` with the Float64 clamps
    x < -744.4400719213812 -> 0 and x > 36.7368005696771 -> 1.
  * LogExpFunctions.jl ``logsumexp``: restated as max-shifted log-sum-exp.
  * Distributions.jl ``rand(Normal(mu, 1.0))`` = ``mu + 1.0 * randn()``.
  * Random ``rand()``/``randn()`` (Xoshiro256++ / ziggurat) are NOT reproduced: streams
    are injected.

Layout: ``W`` is (V, H) exactly like the Julia ``Matrix{Float64}`` (element [i, j] =
visible i, hidden j); ``U`` is (L, H).
"""
from __future__ import annotations

import copy
import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

_LOGISTIC_LOWER = -744.4400719213812
_LOGISTIC_UPPER = 36.7368005696771


# --------------------------------------------------------------------------- streams
class Stream:
    """Injected replacement for Julia's task-local RNG.

    ``rand(n)`` hands out the next n uniforms and ``randn(n)`` the next n normals, in
    call order -- the serial consumption order documented in SURVEY.md section 8(a).
    """

    def __init__(self, uniforms: Optional[np.ndarray] = None, normals: Optional[np.ndarray] = None):
        self.u = np.zeros(0) if uniforms is None else np.asarray(uniforms, dtype=np.float64).ravel()
        self.z = np.zeros(0) if normals is None else np.asarray(normals, dtype=np.float64).ravel()
        self.upos = 0
        self.zpos = 0

    def rand(self, n: int) -> np.ndarray:
        if self.upos + n > self.u.size:
            raise IndexError(f"uniform stream exhausted: need {self.upos + n}, have {self.u.size}")
        out = self.u[self.upos:self.upos + n]
        self.upos += n
        return out

    def randn(self, n: int) -> np.ndarray:
        if self.zpos + n > self.z.size:
            raise IndexError(f"normal stream exhausted: need {self.zpos + n}, have {self.z.size}")
        out = self.z[self.zpos:self.zpos + n]
        self.zpos += n
        return out

    def bernoulli(self, p: np.ndarray) -> np.ndarray:
        """``[rand() < p_i ? 1 : 0 for p_i in p]`` (src/gibbs.jl:5,15,31,48)."""
        return (self.rand(len(p)) < p).astype(np.int64)


class GuardedStream(Stream):
    """Test helper: an injected stream whose near-ties are pushed away from the decision
    boundary.  Any uniform with |u - p| < guard is replaced IN PLACE by p -/+ 2*guard on the
    same side (rounded to fp32), so the sampled trajectory is unchanged while an fp32
    device evaluation of p (|dp| ~ 1e-6) cannot flip a state."""

    def __init__(self, uniforms=None, normals=None, guard: float = 1e-4):
        super().__init__(uniforms, normals)
        self.u = self.u.copy()
        self.guard = guard
        self.nudged = 0

    def bernoulli(self, p: np.ndarray) -> np.ndarray:
        n = len(p)
        lo = self.upos
        u = self.rand(n)
        out = u < p
        near = np.abs(u - p) < self.guard
        if near.any():
            tgt = np.where(out, p - 2 * self.guard, p + 2 * self.guard)
            tgt = np.clip(tgt, 0.0, 1.0 - 2.0 ** -24).astype(np.float32).astype(np.float64)
            ok = (tgt < p) == out  # clipping may fail at p ~ 0 or 1: leave those, tests guard-band them
            idx = np.nonzero(near & ok)[0]
            self.u[lo + idx] = tgt[idx]
            self.nudged += idx.size
        return out.astype(np.int64)


class NumpyStream(Stream):
    """Free-running stream for statistical runs (not for bit parity)."""

    def __init__(self, seed: int):
        super().__init__()
        self.g = np.random.Generator(np.random.PCG64(seed))

    def rand(self, n: int) -> np.ndarray:
        return self.g.random(n)

    def randn(self, n: int) -> np.ndarray:
        return self.g.standard_normal(n)

    def bernoulli(self, p: np.ndarray) -> np.ndarray:
        return (self.g.random(len(p)) < p).astype(np.int64)


# --------------------------------------------------------------------------- model
@dataclass
class Model:
    """`RBM` / `GRBM` / `RBMClassifier` / `GRBMClassifier` (src/rbm.jl:2-39)."""

    W: np.ndarray
    a: np.ndarray
    b: np.ndarray
    U: Optional[np.ndarray] = None
    c: Optional[np.ndarray] = None
    gaussian: bool = False
    # NOT part of the reference (None = the reference): rounding applied to a real-valued Gaussian visible sample when a Gibbs
    # step stores it -- lets tests of the device's bf16 mode follow a device that keeps its chain states in bf16
    v_store: Optional[object] = None

    @property
    def n_visible(self) -> int:
        return self.W.shape[0]

    @property
    def n_hidden(self) -> int:
        return self.W.shape[1]

    @property
    def n_classifiers(self) -> int:
        return 0 if self.U is None else self.U.shape[0]

    @property
    def is_classifier(self) -> bool:
        return self.U is not None


def make_model(V: int, H: int, L: int = 0, gaussian: bool = False, stream: Optional[Stream] = None,
               W: Optional[np.ndarray] = None, U: Optional[np.ndarray] = None) -> Model:
    """Constructors, src/rbm.jl:43-99.

    Without W: ``W = randn(V, H)`` (sigma = 1, column-major fill), then for classifiers
    ``U = randn(L, H)``; biases zero.  With W (3-/5-arg ctor): copies, biases zero --
    except GRBMClassifier's 5-arg ctor which sets biases to 1e-5 (rbm.jl:94-99).
    """
    user_w = W is not None
    if W is None:
        W = stream.randn(V * H).reshape((V, H), order="F").copy()
    else:
        W = np.array(W, dtype=np.float64).reshape(V, H).copy()
    a = np.zeros(V)
    b = np.zeros(H)
    if L == 0:
        return Model(W, a, b, None, None, gaussian)
    if U is None:
        U = stream.randn(L * H).reshape((L, H), order="F").copy()
    else:
        U = np.array(U, dtype=np.float64).reshape(L, H).copy()
    c = np.zeros(L)
    if gaussian and user_w:
        a = np.ones(V) * 1e-5
        b = np.ones(H) * 1e-5
        c = np.ones(L) * 1e-5
    return Model(W, a, b, U, c, gaussian)


def copy_rbm(m: Model) -> Model:
    """src/rbm.jl:231-233,244-246 (deep copy of all fields; see DESIGN.md for the
    reference's GRBMClassifier snapshot quirk, which we do not reproduce)."""
    return copy.deepcopy(m)


# --------------------------------------------------------------------------- conditionals
def logistic(x: np.ndarray) -> np.ndarray:
    """LogExpFunctions.logistic for Float64 (used at src/rbm.jl:165,171,182)."""
    x = np.asarray(x, dtype=np.float64)
    with np.errstate(over="ignore", under="ignore"):
        e = np.exp(x)
        out = e / (1.0 + e)
    out = np.where(x < _LOGISTIC_LOWER, 0.0, np.where(x > _LOGISTIC_UPPER, 1.0, out))
    return out


def conditional_prob_h(m: Model, v: np.ndarray, y: Optional[np.ndarray] = None) -> np.ndarray:
    """src/rbm.jl:165 (2-arg) and :170-171 (3-arg, classifier)."""
    pre = m.b + m.W.T @ np.asarray(v, dtype=np.float64)
    if y is not None:
        pre = pre + m.U.T @ np.asarray(y, dtype=np.float64)
    return logistic(pre)


def visible_mean(m: Model, h: np.ndarray) -> np.ndarray:
    """``a .+ W * h`` -- the shared pre-activation of src/rbm.jl:182 and :187."""
    return m.a + m.W @ np.asarray(h, dtype=np.float64)


def conditional_prob_v(m: Model, h: np.ndarray, stream: Optional[Stream] = None) -> np.ndarray:
    """src/rbm.jl:182 (Bernoulli mean) / :187 (GRBM: returns a *sample* mu + randn())."""
    mu = visible_mean(m, h)
    if m.gaussian:
        return mu + 1.0 * stream.randn(m.n_visible)
    return logistic(mu)


def conditional_prob_y_given_h(m: Model, h: np.ndarray) -> np.ndarray:
    """src/rbm.jl:196-205: softmax(c + U h) through logsumexp."""
    log_num = m.c + m.U @ np.asarray(h, dtype=np.float64)
    mx = np.max(log_num)
    log_den = mx + math.log(float(np.sum(np.exp(log_num - mx))))
    return np.exp(log_num - log_den)


def conditional_prob_y_given_v(m: Model, v: np.ndarray) -> np.ndarray:
    """src/rbm.jl:191-194."""
    return conditional_prob_y_given_h(m, conditional_prob_h(m, v))


def reconstruct(m: Model, v: np.ndarray, stream: Optional[Stream] = None) -> np.ndarray:
    """src/rbm.jl:220-224 (mean-field for RBM, noisy sample for GRBM)."""
    return conditional_prob_v(m, conditional_prob_h(m, v), stream)


def classify(m: Model, v: np.ndarray) -> np.ndarray:
    """src/rbm.jl:226-229: one-hot(s) of the maximum label probability."""
    y = conditional_prob_y_given_v(m, v)
    return (y == np.max(y)).astype(np.int64)


# --------------------------------------------------------------------------- Gibbs
def gibbs_sample_hidden(m: Model, v, stream: Stream, y=None) -> np.ndarray:
    """src/gibbs.jl:3-6 / :29-32: H uniforms in unit order, strict ``rand() < p``."""
    p = conditional_prob_h(m, v, y)
    return stream.bernoulli(p)


def gibbs_sample_visible(m: Model, h, stream: Stream) -> np.ndarray:
    """src/gibbs.jl:13-16 (V uniforms) / :18-20 (GRBM: V normals, real-valued)."""
    if m.gaussian:
        v = conditional_prob_v(m, h, stream)
        return v if m.v_store is None else m.v_store(v)
    p = conditional_prob_v(m, h)
    return stream.bernoulli(p)


def gibbs_sample_label(m: Model, h, stream: Stream) -> np.ndarray:
    """src/gibbs.jl:46-49: softmax probabilities, then INDEPENDENT Bernoulli per unit."""
    p = conditional_prob_y_given_h(m, h)
    return stream.bernoulli(p)


def _get_v_model(m: Model, v, n_gibbs: int, stream: Stream):
    """src/gibbs.jl:57-63."""
    for _ in range(n_gibbs):
        h = gibbs_sample_hidden(m, v, stream)
        v = gibbs_sample_visible(m, h, stream)
    return v


def _get_v_y_model(m: Model, v, y, n_gibbs: int, stream: Stream):
    """src/gibbs.jl:65-72: draw order H, V, L per step."""
    for _ in range(n_gibbs):
        h = gibbs_sample_hidden(m, v, stream, y)
        v = gibbs_sample_visible(m, h, stream)
        y = gibbs_sample_label(m, h, stream)
    return v, y


# --------------------------------------------------------------------------- persistent chains
@dataclass
class FantasyData:
    """src/fantasy_data.jl:1-10 (y is None for non-classifiers)."""

    v: np.ndarray
    h: np.ndarray
    y: Optional[np.ndarray] = None


def _init_fantasy_data(m: Model, batch_size: int, stream: Stream) -> List[FantasyData]:
    """src/fantasy_data.jl:66-86: per chain ``rand(V)``, ``rand(H)``[, ``rand(L)``] --
    continuous U[0,1) floats, not binary."""
    out = []
    for _ in range(batch_size):
        v = stream.rand(m.n_visible).copy()
        h = stream.rand(m.n_hidden).copy()
        y = stream.rand(m.n_classifiers).copy() if m.is_classifier else None
        out.append(FantasyData(v, h, y))
    return out


def _update_fantasy_data(m: Model, chains: List[FantasyData], steps: int, stream: Stream) -> None:
    """src/fantasy_data.jl:12-19 / :21-29: step-major, chain-minor; the stored h is the
    binary sample drawn from the PREVIOUS v and is not refreshed from the new v."""
    for _ in range(steps):
        for ch in chains:
            if m.is_classifier:
                ch.h = gibbs_sample_hidden(m, ch.v, stream, ch.y).astype(np.float64)
                ch.v = np.asarray(gibbs_sample_visible(m, ch.h, stream), dtype=np.float64)
                ch.y = gibbs_sample_label(m, ch.h, stream).astype(np.float64)
            else:
                ch.h = gibbs_sample_hidden(m, ch.v, stream).astype(np.float64)
                ch.v = np.asarray(gibbs_sample_visible(m, ch.h, stream), dtype=np.float64)


# --------------------------------------------------------------------------- updates
@dataclass
class OptState:
    """Extension (absent in the reference): momentum velocity.  momentum = weight_decay
    = 0 collapses ``apply_update`` to src/rbm.jl:120-136,152-163 exactly."""

    vW: Optional[np.ndarray] = None
    va: Optional[np.ndarray] = None
    vb: Optional[np.ndarray] = None
    vU: Optional[np.ndarray] = None
    vc: Optional[np.ndarray] = None


def apply_update(m: Model, dW, da, db, lr: float, dU=None, dc=None, llr: float = 0.0,
                 momentum: float = 0.0, weight_decay: float = 0.0, opt: Optional[OptState] = None) -> None:
    """``update_rbm!(rbm, dW[, dU], da, db[, dc], lr[, llr])`` (src/rbm.jl:120-136,152-163):
    ``W .+= dW .* lr`` etc.  Extension: vel = momentum*vel + lr*d - lr*wd*W (W, U only)."""
    if momentum == 0.0 and weight_decay == 0.0:
        m.W += dW * lr
        if dU is not None:
            m.U += dU * llr
        m.a += da * lr
        m.b += db * lr
        if dc is not None:
            m.c += dc * llr
        return
    assert opt is not None
    if opt.vW is None:
        opt.vW = np.zeros_like(m.W)
        opt.va = np.zeros_like(m.a)
        opt.vb = np.zeros_like(m.b)
        if m.is_classifier:
            opt.vU = np.zeros_like(m.U)
            opt.vc = np.zeros_like(m.c)
    opt.vW = momentum * opt.vW + lr * dW - lr * weight_decay * m.W
    opt.va = momentum * opt.va + lr * da
    opt.vb = momentum * opt.vb + lr * db
    m.W += opt.vW
    m.a += opt.va
    m.b += opt.vb
    if dU is not None:
        opt.vU = momentum * opt.vU + llr * dU - llr * weight_decay * m.U
        opt.vc = momentum * opt.vc + llr * dc
        m.U += opt.vU
        m.c += opt.vc


# --------------------------------------------------------------------------- CD
def contrastive_divergence(m: Model, x: Sequence, *, steps: int, learning_rate: float, stream: Stream,
                           y: Optional[Sequence] = None, label_learning_rate: float = 0.1,
                           batch_size: int = 1, momentum: float = 0.0, weight_decay: float = 0.0,
                           opt: Optional[OptState] = None, trace: Optional[list] = None) -> None:
    """One epoch of CD-k.

    batch_size == 1 is the reference: src/training/train_cd.jl:2-21 (and :23-43 for
    classifiers): online, W mutates after every sample, full ``lr`` (no 1/B); the
    negative hidden term is the PROBABILITY ``conditional_prob_h(rbm, v_model)`` and for
    classifiers the 2-arg form that ignores y_model (:34).

    batch_size > 1 is the declared mini-batch extension (SURVEY.md section 8 "semantic
    deltas"): the per-sample statistics above are summed over B consecutive samples with
    W frozen, then applied once with lr/B (PCD's scaling, train_pcd.jl:39); a trailing
    ragged batch is dropped like `_set_mini_batches` (utils.jl:13-26).  The random
    stream is consumed sample-major within the batch (sample i: all k steps), i.e. the
    same serial order as the reference's loop.
    """
    n = len(x)
    nb = n // batch_size
    for bi in range(nb):
        dW = np.zeros_like(m.W)
        da = np.zeros_like(m.a)
        db = np.zeros_like(m.b)
        dU = np.zeros_like(m.U) if m.is_classifier else None
        dc = np.zeros_like(m.c) if m.is_classifier else None
        for i in range(bi * batch_size, (bi + 1) * batch_size):
            v_data = np.asarray(x[i], dtype=np.float64)
            if m.is_classifier:
                y_data = np.asarray(y[i], dtype=np.float64)
                h_data = conditional_prob_h(m, v_data, y_data)
                v_model, y_model = _get_v_y_model(m, v_data, y_data, steps, stream)
                h_model = conditional_prob_h(m, v_model)
                dU += np.outer(y_data, h_data) - np.outer(y_model, h_model)
                dc += y_data - y_model
            else:
                h_data = conditional_prob_h(m, v_data)
                v_model = _get_v_model(m, v_data, steps, stream)
                h_model = conditional_prob_h(m, v_model)
            dW += np.outer(v_data, h_data) - np.outer(v_model, h_model)
            da += v_data - v_model
            db += h_data - h_model
            if trace is not None:
                trace.append(dict(v_model=np.array(v_model, dtype=np.float64), h_data=h_data.copy(),
                                  h_model=h_model.copy(),
                                  y_model=None if not m.is_classifier else np.array(y_model, dtype=np.float64)))
        apply_update(m, dW, da, db, learning_rate / batch_size, dU, dc, label_learning_rate / batch_size,
                     momentum, weight_decay, opt)


# --------------------------------------------------------------------------- PCD
def _set_mini_batches(training_set_length: int, batch_size: int) -> List[range]:
    """src/utils.jl:13-26: contiguous ranges, trailing remainder dropped, no shuffle."""
    last = training_set_length % batch_size
    if last > 0:
        training_set_length -= last
    nb = int(round(training_set_length / batch_size))
    return [range(i * batch_size, (i + 1) * batch_size) for i in range(nb)]


def persistent_contrastive_divergence(m: Model, x: Sequence, mini_batches: List[range],
                                      chains: List[FantasyData], *, steps: int, learning_rate: float,
                                      stream: Stream, y: Optional[Sequence] = None,
                                      label_learning_rate: float = 0.1, momentum: float = 0.0,
                                      weight_decay: float = 0.0, opt: Optional[OptState] = None,
                                      trace: Optional[list] = None) -> None:
    """One epoch of PCD-k.

    Non-classifier (src/training/train_pcd.jl:2-43): per mini-batch the chains advance k
    steps FIRST (:14-16), then sample i pairs with chain i (:30): dW += x h_d' - v_c h_c'
    with the chain's BINARY sampled h, then `update_rbm!` with lr/B (:39).

    Classifier (:46-95): gradient from the chains as they are, update, and the chains
    advance AFTER the update (:89-92) -- so the first mini-batch's negative statistics
    are the raw U[0,1) floats of `_init_fantasy_data`.
    """
    for mb in mini_batches:
        B = len(mb)
        if len(chains) != B:
            _pcd_minibatch_more_chains(m, x, mb, chains, steps, learning_rate, stream, y, label_learning_rate, momentum,
                                       weight_decay, opt)
            if trace is not None:
                trace.append(dict(v=np.stack([c.v for c in chains]), h=np.stack([c.h for c in chains]),
                                  y=None if not m.is_classifier else np.stack([c.y for c in chains]),
                                  W=m.W.copy(), a=m.a.copy(), b=m.b.copy()))
            continue
        if not m.is_classifier:
            _update_fantasy_data(m, chains, steps, stream)
        dW = np.zeros_like(m.W)
        da = np.zeros_like(m.a)
        db = np.zeros_like(m.b)
        dU = np.zeros_like(m.U) if m.is_classifier else None
        dc = np.zeros_like(m.c) if m.is_classifier else None
        for index, i in enumerate(mb):
            v_data = np.asarray(x[i], dtype=np.float64)
            ch = chains[index]
            if m.is_classifier:
                y_data = np.asarray(y[i], dtype=np.float64)
                h_data = conditional_prob_h(m, v_data, y_data)
                dU += np.outer(y_data, h_data) - np.outer(ch.y, ch.h)
                dc += y_data - ch.y
            else:
                h_data = conditional_prob_h(m, v_data)
            dW += np.outer(v_data, h_data) - np.outer(ch.v, ch.h)
            da += v_data - ch.v
            db += h_data - ch.h
        apply_update(m, dW, da, db, learning_rate / B, dU, dc, label_learning_rate / B,
                     momentum, weight_decay, opt)
        if m.is_classifier:
            _update_fantasy_data(m, chains, steps, stream)
        if trace is not None:
            trace.append(dict(v=np.stack([c.v for c in chains]), h=np.stack([c.h for c in chains]),
                              y=None if not m.is_classifier else np.stack([c.y for c in chains]),
                              W=m.W.copy(), a=m.a.copy(), b=m.b.copy()))


def _pcd_minibatch_more_chains(m, x, mb, chains, steps, learning_rate, stream, y, label_learning_rate, momentum,
                               weight_decay, opt) -> None:
    """Declared extension (BASELINE.json config 2: "1000 persistent chains, batch 256"; the reference
    ties n_chains to batch_size, train_pcd.jl:162): with Nc != B chains, every chain advances k steps per
    mini-batch and the negative statistics are averaged over ALL chains,
        dW/lr = (1/B) sum_data x h_d' - (1/Nc) sum_chains v h',
    implemented as  lr/Nc * ( (Nc/B) sum_data - sum_chains )  so that Nc == B is the reference update."""
    B, Nc = len(mb), len(chains)
    clf = m.is_classifier
    if not clf:
        _update_fantasy_data(m, chains, steps, stream)
    scale = Nc / B
    dW = np.zeros_like(m.W); da = np.zeros_like(m.a); db = np.zeros_like(m.b)
    dU = np.zeros_like(m.U) if clf else None
    dc = np.zeros_like(m.c) if clf else None
    for i in mb:
        v_data = np.asarray(x[i], dtype=np.float64)
        if clf:
            y_data = np.asarray(y[i], dtype=np.float64)
            h_data = conditional_prob_h(m, v_data, y_data)
            dU += scale * np.outer(y_data, h_data)
            dc += scale * y_data
        else:
            h_data = conditional_prob_h(m, v_data)
        dW += scale * np.outer(v_data, h_data)
        da += scale * v_data
        db += scale * h_data
    for ch in chains:
        dW -= np.outer(ch.v, ch.h)
        da -= ch.v
        db -= ch.h
        if clf:
            dU -= np.outer(ch.y, ch.h)
            dc -= ch.y
    apply_update(m, dW, da, db, learning_rate / Nc, dU, dc, label_learning_rate / Nc, momentum, weight_decay, opt)
    if clf:
        _update_fantasy_data(m, chains, steps, stream)


def _update_fantasy_data_fast(m: Model, eff: Model, chains: List[FantasyData], steps: int, stream: Stream) -> None:
    """`_update_fantasy_data!` with fast weights (src/fantasy_data.jl:31-64).  `eff` holds the effective
    parameters W + W_fast, ... (src/rbm.jl:167-168,173-189,207-218).  The fast `gibbs_sample_visible` has a
    single method for every model (src/gibbs.jl:22-25): it thresholds ``rand() < conditional_prob_v(...)``,
    and for Gaussian visibles that "probability" is the noisy sample mu + randn() (src/rbm.jl:188-189) --
    so a GRBM's chain visibles become BINARY here; draw order per chain: H uniforms, V normals, V uniforms
    [, L uniforms]."""
    for _ in range(steps):
        for ch in chains:
            if m.is_classifier:
                ch.h = gibbs_sample_hidden(eff, ch.v, stream, ch.y).astype(np.float64)
            else:
                ch.h = gibbs_sample_hidden(eff, ch.v, stream).astype(np.float64)
            ch.v = stream.bernoulli(conditional_prob_v(eff, ch.h, stream)).astype(np.float64)
            if m.is_classifier:
                ch.y = gibbs_sample_label(eff, ch.h, stream).astype(np.float64)


def fast_persistent_contrastive_divergence(m: Model, x: Sequence, mini_batches: List[range], chains: List[FantasyData], *,
                                           steps: int, learning_rate: float, fast_learning_rate: float, stream: Stream,
                                           y: Optional[Sequence] = None, label_learning_rate: float = 0.1,
                                           fast_label_learning_rate: float = 0.1, pair_chains: Optional[bool] = None) -> None:
    """One epoch of FastPCD (src/training/train_fast_pcd.jl:1-57, classifier :59-127; Tieleman & Hinton 2009).

    The fast weights start at zero EVERY epoch (:12-14).  Inside a mini-batch the reference is online:
    for each sample, h_data uses the current slow weights, the slow weights are updated right away with
    lr/B (`update_rbm!` per-sample form, src/rbm.jl:101-118,138-150) and the fast weights are decayed by
    19/20 and pushed by fast_lr/B times the same difference -- per SAMPLE (:34-43).  The chains then advance
    with the effective parameters (see `_update_fantasy_data_fast`).

    Classifier quirk: `batch_index` is initialised to 1 per mini-batch and never incremented (:77-119), so
    EVERY sample of a mini-batch is paired with chain 1.  `pair_chains=None` follows the reference (False for
    classifiers, True otherwise); `pair_chains=True` is the corrected pairing (extension)."""
    clf = m.is_classifier
    if pair_chains is None:
        pair_chains = not clf
    W_fast = np.zeros_like(m.W); a_fast = np.zeros_like(m.a); b_fast = np.zeros_like(m.b)
    if clf:
        U_fast = np.zeros_like(m.U); c_fast = np.zeros_like(m.c)
    for mb in mini_batches:
        B = len(mb)
        lr, flr = learning_rate / B, fast_learning_rate / B
        llr, fllr = label_learning_rate / B, fast_label_learning_rate / B
        for batch_index, i in enumerate(mb):
            v_data = np.asarray(x[i], dtype=np.float64)
            y_data = np.asarray(y[i], dtype=np.float64) if clf else None
            h_data = conditional_prob_h(m, v_data, y_data)
            ch = chains[batch_index if pair_chains else 0]
            dW = np.outer(v_data, h_data) - np.outer(ch.v, ch.h)
            m.W += lr * dW
            m.a += lr * (v_data - ch.v)
            m.b += lr * (h_data - ch.h)
            W_fast = W_fast * (19 / 20) + flr * dW
            a_fast = a_fast * (19 / 20) + flr * (v_data - ch.v)
            b_fast = b_fast * (19 / 20) + flr * (h_data - ch.h)
            if clf:
                dU = np.outer(y_data, h_data) - np.outer(ch.y, ch.h)
                m.U += llr * dU
                m.c += llr * (y_data - ch.y)
                U_fast = U_fast * (19 / 20) + fllr * dU
                c_fast = c_fast * (19 / 20) + fllr * (y_data - ch.y)
        eff = Model(m.W + W_fast, m.a + a_fast, m.b + b_fast, (m.U + U_fast) if clf else None,
                    (m.c + c_fast) if clf else None, m.gaussian)
        _update_fantasy_data_fast(m, eff, chains, steps, stream)


def persistent_qubo_minibatch(m: Model, x: Sequence, mb: range, v_model, h_model, *, learning_rate: float,
                              y: Optional[Sequence] = None, y_model=None, label_learning_rate: float = 0.1) -> None:
    """One mini-batch of `persistent_qubo!` (src/training/train_quantum_pcd.jl:10-37, classifier :55-83)
    with the sampler's output (v_model, h_model[, label_model]) injected: every sample of the batch is
    paired with the SAME model sample; update with lr / B."""
    B = len(mb)
    v_model = np.asarray(v_model, dtype=np.float64); h_model = np.asarray(h_model, dtype=np.float64)
    dW = np.zeros_like(m.W); da = np.zeros_like(m.a); db = np.zeros_like(m.b)
    dU = np.zeros_like(m.U) if m.is_classifier else None
    dc = np.zeros_like(m.c) if m.is_classifier else None
    for i in mb:
        v_data = np.asarray(x[i], dtype=np.float64)
        if m.is_classifier:
            y_data = np.asarray(y[i], dtype=np.float64)
            h_data = conditional_prob_h(m, v_data, y_data)
            dU += np.outer(y_data, h_data) - np.outer(y_model, h_model)
            dc += y_data - np.asarray(y_model, dtype=np.float64)
        else:
            h_data = conditional_prob_h(m, v_data)
        dW += np.outer(v_data, h_data) - np.outer(v_model, h_model)
        da += v_data - v_model
        db += h_data - h_model
    apply_update(m, dW, da, db, learning_rate / B, dU, dc, label_learning_rate / B)


# --------------------------------------------------------------------------- evaluation
def evaluate_mse(m: Model, x_dataset: Sequence, stream: Optional[Stream] = None) -> float:
    """`_evaluate(MeanSquaredError)` through `evaluate` (src/evaluation.jl:16-20,52-70):
    sum over samples of sum_units (x - reconstruct(x))^2 / N."""
    n = len(x_dataset)
    acc = 0.0
    for vis in x_dataset:
        vis = np.asarray(vis, dtype=np.float64)
        pred = reconstruct(m, vis, stream)
        acc += float(np.sum((vis - pred) ** 2)) / n
    return acc


def evaluate_accuracy(m: Model, x_dataset: Sequence, y_dataset: Sequence) -> float:
    """`_evaluate(Accuracy)` through the classifier `evaluate` (src/evaluation.jl:9-14,28-50)."""
    n = len(x_dataset)
    acc = 0.0
    for vis, lab in zip(x_dataset, y_dataset):
        pred = classify(m, np.asarray(vis, dtype=np.float64))
        tp = 1 if np.all(np.rint(np.asarray(lab)).astype(np.int64) == pred) else 0
        acc += tp / n
    return acc


def free_energy(m: Model, v: np.ndarray, y: Optional[np.ndarray] = None) -> float:
    """New metric (no reference function).  Sign convention from the QUBO objective
    src/qubo.jl:43-47 (:50-55 with labels): E(v,h) = -v'Wh - a'v - b'h [- y'Uh - c'y], so
    F(v[,y]) = -a'v [- c'y] - sum_j softplus(b_j + W_j'v [+ U_j'y]);
    Gaussian visibles (unit variance, rbm.jl:187): 0.5*||v - a||^2 replaces -a'v."""
    v = np.asarray(v, dtype=np.float64)
    pre = m.b + m.W.T @ v
    if y is not None:
        y = np.asarray(y, dtype=np.float64)
        pre = pre + m.U.T @ y
    softplus = np.logaddexp(0.0, pre)
    if m.gaussian:
        f = 0.5 * float(np.sum((v - m.a) ** 2))
    else:
        f = -float(m.a @ v)
    if y is not None:
        f -= float(m.c @ y)
    return f - float(np.sum(softplus))


def mean_free_energy(m: Model, x_dataset: Sequence, y_dataset: Optional[Sequence] = None) -> float:
    n = len(x_dataset)
    tot = 0.0
    for i, vis in enumerate(x_dataset):
        tot += free_energy(m, vis, None if y_dataset is None else y_dataset[i])
    return tot / n


def _diverged_mse(history: Sequence[float]) -> bool:
    """src/evaluation.jl:132-142."""
    if len(history) == 1:
        return False
    return history[-1] > history[-2] or history[-1] > min(history)


def _diverged_accuracy(history: Sequence[float]) -> bool:
    """src/evaluation.jl:156-168."""
    if len(history) == 1:
        return False
    return history[-1] < history[-2] or history[-1] < max(history)


# --------------------------------------------------------------------------- train!
def train(m: Model, x_train: Sequence, method: str, *, n_epochs: int, learning_rate: Sequence[float],
          stream: Stream, gibbs_steps: Optional[int] = None, batch_size: Optional[int] = None,
          label_train: Optional[Sequence] = None, label_learning_rate: Optional[Sequence[float]] = None,
          early_stopping: bool = False, store_best_rbm: bool = True, patience: int = 10,
          x_test_dataset: Optional[Sequence] = None, momentum: float = 0.0, weight_decay: float = 0.0) -> dict:
    """`train!` epoch loops: src/training/train_cd.jl:45-119,121-200 and
    src/training/train_pcd.jl:133-215,258-345 (MSE metric for generative models, accuracy
    for classifiers; logging and CSV are host text I/O and not restated).  Returns the
    merged metric history (initial evaluation first)."""
    method = method.upper()
    clf = m.is_classifier
    if gibbs_steps is None:
        gibbs_steps = 3 if method == "CD" else 1
    best = copy_rbm(m)
    initial_patience = patience
    eval_x = x_train if x_test_dataset is None else x_test_dataset

    def metric() -> float:
        if clf:
            return evaluate_accuracy(m, x_train, label_train)
        return evaluate_mse(m, eval_x, stream)

    initial = [metric()]
    hist: List[float] = []
    opt = OptState()
    if method == "PCD":
        mini_batches = _set_mini_batches(len(x_train), batch_size)
        chains = _init_fantasy_data(m, batch_size, stream)
    for epoch in range(n_epochs):
        lr = learning_rate[epoch]
        llr = label_learning_rate[epoch] if clf else 0.0
        if method == "CD":
            contrastive_divergence(m, x_train, steps=gibbs_steps, learning_rate=lr, stream=stream,
                                   y=label_train, label_learning_rate=llr,
                                   batch_size=1 if batch_size is None else batch_size,
                                   momentum=momentum, weight_decay=weight_decay, opt=opt)
        else:
            persistent_contrastive_divergence(m, x_train, mini_batches, chains, steps=gibbs_steps,
                                              learning_rate=lr, stream=stream, y=label_train,
                                              label_learning_rate=llr, momentum=momentum,
                                              weight_decay=weight_decay, opt=opt)
        hist.append(metric())
        diverged = _diverged_accuracy(hist) if clf else _diverged_mse(hist)
        if diverged:
            if early_stopping:
                if patience == 0:
                    break
                patience -= 1
        else:
            patience = initial_patience
            if store_best_rbm:
                best = copy_rbm(m)
    if store_best_rbm:
        m.W[...] = best.W
        m.a[...] = best.a
        m.b[...] = best.b
        if clf:
            m.U[...] = best.U
            m.c[...] = best.c
    return {("accuracy" if clf else "mse"): initial + hist}


# --------------------------------------------------------------------------- stream helpers
def serial_uniforms_pcd(U_h: np.ndarray, U_v: Optional[np.ndarray], U_y: Optional[np.ndarray]) -> np.ndarray:
    """Interleave dense injection arrays U_h[k][B][H], U_v[k][B][V], U_y[k][B][L] (the
    C-ABI layout) into the reference's serial order: step-major, chain-minor, (H, V, L)
    within a chain (fantasy_data.jl:12-29).  U_v is None for Gaussian visibles."""
    k, B, _ = U_h.shape
    parts = []
    for s in range(k):
        for i in range(B):
            parts.append(U_h[s, i])
            if U_v is not None:
                parts.append(U_v[s, i])
            if U_y is not None:
                parts.append(U_y[s, i])
    return np.concatenate([np.asarray(p, dtype=np.float64).ravel() for p in parts])


def serial_uniforms_cd(U_h: np.ndarray, U_v: Optional[np.ndarray], U_y: Optional[np.ndarray]) -> np.ndarray:
    """Same dense arrays, CD order: sample-major, step-minor (train_cd.jl:4-19 +
    gibbs.jl:57-72)."""
    k, B, _ = U_h.shape
    parts = []
    for i in range(B):
        for s in range(k):
            parts.append(U_h[s, i])
            if U_v is not None:
                parts.append(U_v[s, i])
            if U_y is not None:
                parts.append(U_y[s, i])
    return np.concatenate([np.asarray(p, dtype=np.float64).ravel() for p in parts])


def serial_normals(Z_v: np.ndarray, order: str) -> np.ndarray:
    """Z_v[k][B][V] -> serial normal stream ('pcd': step-major; 'cd': sample-major)."""
    if order == "pcd":
        return np.asarray(Z_v, dtype=np.float64).reshape(-1)
    return np.asarray(Z_v, dtype=np.float64).transpose(1, 0, 2).reshape(-1)


# --------------------------------------------------------------------------- batched (BLAS-3) formulation
def batched_step(m: Model, X: np.ndarray, Y: Optional[np.ndarray], algo: str, k: int, lr: float, llr: float = 0.0, *,
                 cv: Optional[np.ndarray] = None, ch: Optional[np.ndarray] = None, cy: Optional[np.ndarray] = None,
                 U_h=None, U_v=None, U_y=None, Z_v=None, rng: Optional[np.random.Generator] = None, v_store=None):
    """One mini-batch step of PCD (a20/a22) or mini-batch CD (a16/a17 with the batch extension) written with
    matrix-matrix products over the whole batch -- the formulation the device implements, and the BEST CASE for a
    CPU (threaded dgemm instead of the reference's per-sample gemv + outer products).  Used (1) by tests to show it
    equals the per-sample restatement on the same streams and (2) by bench.py as the `value_batched_blas` CPU figure.
    Random numbers: dense arrays U_h[k][B][H], U_v[k][B][V] (or Z_v), U_y[k][B][L], or a numpy Generator.
    `v_store` (tests of the device's bf16 mode only; None = the reference): rounding applied to a real-valued Gaussian
    visible sample when it is stored, so the oracle follows a device that keeps its states in bf16."""
    clf, B = m.is_classifier, X.shape[0]

    def uni(arr, s, shape):
        return arr[s] if arr is not None else rng.random(shape)

    def hidden_pre(v, y):
        pre = m.b + v @ m.W
        return pre + y @ m.U if (clf and y is not None) else pre

    def gibbs(v, y, s):
        h = (uni(U_h, s, (v.shape[0], m.n_hidden)) < logistic(hidden_pre(v, y))).astype(np.float64)
        mu = m.a + h @ m.W.T
        if m.gaussian:
            v2 = mu + (Z_v[s] if Z_v is not None else rng.standard_normal(mu.shape))
            if v_store is not None:
                v2 = v_store(v2)
        else:
            v2 = (uni(U_v, s, mu.shape) < logistic(mu)).astype(np.float64)
        y2 = None
        if clf:
            lp = m.c + h @ m.U.T
            lp = lp - lp.max(axis=1, keepdims=True)
            py = np.exp(lp - np.log(np.exp(lp).sum(axis=1, keepdims=True)))
            y2 = (uni(U_y, s, py.shape) < py).astype(np.float64)
        return v2, h, y2

    if algo.upper() == "PCD":
        if not clf:                                        # chains first (train_pcd.jl:15-17)
            for s in range(k):
                cv, ch, _ = gibbs(cv, None, s)
        hd = logistic(hidden_pre(X, Y))
        dW = X.T @ hd - cv.T @ ch
        da, db = (X - cv).sum(0), (hd - ch).sum(0)
        if clf:
            dU, dc = Y.T @ hd - cy.T @ ch, (Y - cy).sum(0)
    else:
        hd = logistic(hidden_pre(X, Y))
        v, y, h = X, Y, None
        for s in range(k):
            v, h, y = gibbs(v, y, s)
        hm = logistic(hidden_pre(v, None))                 # 2-arg form: ignores y~ (train_cd.jl:34)
        dW = X.T @ hd - v.T @ hm
        da, db = (X - v).sum(0), (hd - hm).sum(0)
        if clf:
            dU, dc = Y.T @ hd - y.T @ hm, (Y - y).sum(0)
    m.W += dW * (lr / B)
    m.a += da * (lr / B)
    m.b += db * (lr / B)
    if clf:
        m.U += dU * (llr / B)
        m.c += dc * (llr / B)
        if algo.upper() == "PCD":                          # classifier PCD: chains advance AFTER the update (train_pcd.jl:89-92)
            for s in range(k):
                cv, ch, cy = gibbs(cv, cy, s)
    return cv, ch, cy

"""``Gate`` (nodes/gate.py): a mixture written with a gating node, scalar and vector-valued, gated plate -1 and -2 —
bound trajectories and every node's moments against the reference (tests/golden/gate.npz, make_golden.py: gate_models)."""
import numpy as np
import pytest

from conftest import golden


def _check(g, nodes):
    for nm, node in nodes:
        for i in range(len(node.u)):
            np.testing.assert_allclose(np.asarray(node.u[i]), g["%s_u%d" % (nm, i)], rtol=1e-8, atol=1e-10,
                                       err_msg="%s.u[%d]" % (nm, i))


def test_gate_mixture_matches_the_reference(backend):
    from bayespy_b200.nodes import Dirichlet, Categorical, GaussianARD, Gamma, Gate
    from bayespy_b200.inference import VB
    g = golden("gate")
    N, K = len(g["y"]), 3
    alpha = Dirichlet(np.ones(K), name="alpha")
    Z = Categorical(alpha, plates=(N,), name="Z")
    Z.initialize_from_value(g["z_init"])
    mu = GaussianARD(0, 1e-3, plates=(K,), name="mu")
    mu.initialize_from_value(np.array([-1.0, 0.5, 2.0]))
    F = Gate(Z, mu, name="F")
    assert tuple(F.plates) == (N,)
    tau = Gamma(1e-3, 1e-3, name="tau")
    Y = GaussianARD(F, tau, name="Y")
    Y.observe(g["y"])
    Q = VB(Z, mu, alpha, tau, Y)
    Q.update(repeat=5, verbose=False, tol=0)
    np.testing.assert_allclose(Q.L[:5], g["L"], rtol=1e-9)
    _check(g, (("Z", Z), ("mu", mu), ("alpha", alpha), ("tau", tau)))
    for i in range(2):
        np.testing.assert_allclose(np.asarray(F.get_moments()[i]), g["F_u%d" % i], rtol=1e-8)


def test_gate_vector_valued_inner_plate(backend):
    from bayespy_b200.nodes import Categorical, GaussianARD, Gate
    from bayespy_b200.inference import VB
    g = golden("gate")
    K = 3
    X = GaussianARD(0, 1, plates=(K, 4), shape=(2,), name="X")
    X.initialize_from_value(g["X_init"])
    Z2 = Categorical(np.ones(K) / K, plates=(6, 1), name="Z2")
    Z2.initialize_from_value(g["z2_init"])
    G2 = Gate(Z2, X, gated_plate=-2, name="G2")
    assert tuple(G2.plates) == (6, 4)
    W = GaussianARD(G2, 1.5, name="W")
    W.observe(g["y2"])
    Q = VB(Z2, X, W)
    Q.update(repeat=3, verbose=False, tol=0)
    np.testing.assert_allclose(Q.L[:3], g["L2"], rtol=1e-9)
    np.testing.assert_allclose(np.asarray(G2.get_moments()[0]), g["G2_u0"], rtol=1e-8, atol=1e-10)
    _check(g, (("Z2", Z2), ("X", X)))


def test_gate_argument_checks(oracle_backend):
    from bayespy_b200.nodes import GaussianARD, Gate
    X = GaussianARD(0, 1, plates=(3,))
    with pytest.raises(ValueError):
        Gate([0, 1], X, gated_plate=0)
    with pytest.raises(ValueError):
        Gate([0, 1], X, gated_plate=-2)
    with pytest.raises(ValueError):
        Gate([0, 5], X)                       # label outside the gated plate
    F = Gate([0, 0, 2, 1], X)
    assert tuple(F.plates) == (4,)


def _lda_nodes(g, n_docs, n_vocab, n_topics, doc_idx, n, multiplier=None):
    from bayespy_b200 import nodes
    from bayespy_b200.inference.vmp.nodes.categorical import CategoricalMoments
    p_topic = nodes.Dirichlet(np.ones(n_topics), plates=(n_docs,), name="p_topic")
    p_word = nodes.Dirichlet(np.ones(n_vocab), plates=(n_topics,), name="p_word")
    document_indices = nodes.Constant(CategoricalMoments(n_docs), doc_idx, name="document_indices")
    kw = {} if multiplier is None else dict(plates_multiplier=(multiplier,))
    topics = nodes.Categorical(nodes.Gate(document_indices, p_topic), plates=(n,), name="topics", **kw)
    words = nodes.Categorical(nodes.Gate(topics, p_word), name="words")
    return p_topic, p_word, document_indices, topics, words


def test_lda_example_batch_and_stochastic(backend):
    """doc/source/examples/lda.rst scaled down: latent Dirichlet allocation from Dirichlet / Categorical / Gate and a
    constant categorical index node — batch VB (:84-135), then stochastic VI (:180-262): mini-batches through
    ``Constant.set_value`` + ``observe``, ``plates_multiplier`` inherited through the gates, natural-gradient steps of
    the two Dirichlet nodes — against the unmodified reference."""
    from bayespy_b200.inference import VB
    g = golden("lda_small")
    wd, corpus = g["word_documents"], g["corpus"]
    n_docs, n_topics = g["p_topic_init"].shape
    n_vocab = g["p_word_init"].shape[1]
    p_topic, p_word, document_indices, topics, words = _lda_nodes(g, n_docs, n_vocab, n_topics, wd, len(corpus))
    words.observe(corpus)
    # the reference's random initialisation, draw for draw (utils/random.py:329-347: normalised gamma draws)
    np.random.seed(5)
    p_topic.initialize_from_random()
    p_word.initialize_from_random()
    np.testing.assert_allclose(np.exp(np.asarray(p_topic.u[0])), g["p_topic_init"], rtol=1e-12)
    np.testing.assert_allclose(np.exp(np.asarray(p_word.u[0])), g["p_word_init"], rtol=1e-12)
    Q = VB(words, topics, p_word, p_topic, document_indices)
    iters = len(g["L"])
    Q.update(repeat=iters, verbose=False, tol=0)
    np.testing.assert_allclose(Q.L[:iters], g["L"], rtol=1e-9)
    for nm, node in (("p_topic", p_topic), ("p_word", p_word), ("topics", topics)):
        np.testing.assert_allclose(np.asarray(node.u[0]), g[nm + "_u0"], rtol=1e-8, atol=1e-10, err_msg=nm)
    # stochastic variational inference
    subsets = g["subsets"]
    S = subsets.shape[1]
    mult = len(corpus) / S
    p_topic, p_word, document_indices, topics, words = _lda_nodes(g, n_docs, n_vocab, n_topics, wd[:S], S, multiplier=mult)
    assert tuple(words.plates_multiplier) == (mult,) and tuple(p_word.plates_multiplier) == ()
    p_topic.initialize_from_value(g["p_topic_init"])
    p_word.initialize_from_value(g["p_word_init"])
    Q = VB(words, topics, p_word, p_topic, document_indices)
    Q.ignore_bound_checks = True
    for n in range(len(subsets)):
        Q["words"].observe(corpus[subsets[n]])
        Q["document_indices"].set_value(wd[subsets[n]])
        Q.update("topics", verbose=False)
        Q.gradient_step("p_topic", "p_word", scale=(n + 1) ** (-0.7))
        np.testing.assert_allclose(np.asarray(p_topic.u[0]), g["svi_p_topic"][n], rtol=1e-8, atol=1e-10,
                                   err_msg="p_topic, step %d" % n)
        np.testing.assert_allclose(np.asarray(p_word.u[0]), g["svi_p_word"][n], rtol=1e-8, atol=1e-10,
                                   err_msg="p_word, step %d" % n)
    np.testing.assert_allclose(Q.L[:Q.iter], g["svi_L"], rtol=1e-9)
    with pytest.raises(ValueError):
        Q["document_indices"].set_value(wd[:S - 1])

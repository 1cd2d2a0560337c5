"""The calls with which the reference's demo scripts (bayespy/demos/*.py) are run — small sizes, plotting off — both
when make_golden.py records the reference's printed bounds and when tests/test_reference_demos.py replays the
unmodified scripts through this package."""

CALLS = {
    "pca": lambda m: m.run(M=8, N=40, D_y=2, D=4, rotate=True, maxiter=12, plot=False),
    "lssm": lambda m: m.demo(M=4, N=40, D=3, maxiter=8, rotate=True, plot=False, monitor=False),
    "mog": lambda m: m.run(N=40, K=4, D=2),          # also prints the integral of the predictive density over a grid
    "categorical": lambda m: m.run(M=10, D=3),
    "gamma_shape": lambda m: m.run(),
    "hmm": lambda m: m.run(N=60, maxiter=5, plot=False),
    "annealing": lambda m: m.run(N=100, maxiter=15, plot=False),
    "pattern_search": lambda m: m.run(M=8, N=30, D_y=2, D=4, maxiter=15, plot=False),
    "stochastic_inference": lambda m: m.run(N=2000, N_batch=50, maxiter=8, plot=False),
    "lda": lambda m: m.run(n_documents=5, n_topics=3, n_vocabulary=8, n_words=300, maxiter=5, seed=1),
    # (lssm_tvd.demo() always monitors with plots: the same model and inference through its infer())
    "lssm_tvd": lambda m: m.infer(m.simulate_data(60)[0][None, :], 2, 2, maxiter=5, monitor=False, update_hyper=2),
    "lssm_sd": lambda m: m.demo(N=60, maxiter=5, D=2, K=2, plot=False, monitor=False),
}

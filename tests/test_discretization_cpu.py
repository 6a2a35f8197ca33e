"""uniform_bins (GPU-friendly min/max fit) against sklearn's KBinsDiscretizer, which the reference uses
(discretization.py:105-118), on CPU."""
import numpy as np
import torch

from bitswap_b200.discretization import uniform_bins


def test_uniform_bins_equals_sklearn_kbins():
    from sklearn.preprocessing import KBinsDiscretizer
    rs = np.random.RandomState(0)
    samples = rs.logistic(size=(4000, 12)).astype(np.float16)          # the reference stores samples in float16
    for q in (4, 6):
        est = KBinsDiscretizer(n_bins=1 << q, strategy="uniform")
        # The float16 samples are handed over as float64: with the reference's pinned stack (sklearn 0.20.1 on
        # numpy 1.x, README.md:91-97) np.linspace(float16 min, float16 max) already computed float64 edges; under
        # numpy >= 2 (NEP 50) the same call would round the edges to float16, which is not what produced the
        # reference's published bins.
        est.fit(samples.astype(np.float64))
        edges = np.array([np.array(a) for a in est.bin_edges_]).transpose()      # reference: :113-116
        centres = (edges[:-1, :] + edges[1:, :]) / 2
        want_e, want_c = edges[1:-1].transpose(), centres.transpose()
        e, c = uniform_bins(torch.from_numpy(samples.astype(np.float64)), q)
        assert e.shape == want_e.shape and c.shape == want_c.shape
        assert np.abs(e.numpy() - want_e).max() < 1e-12 and np.abs(c.numpy() - want_c).max() < 1e-12

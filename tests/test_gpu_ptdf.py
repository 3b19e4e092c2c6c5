"""calculate_ptdf on the GPU (dopf_calculate_ptdf, /root/reference/src/helpers/ptdf.jl:1-41) against the reference's known
answer, the oracle's Gauss-Jordan restatement and the host construction.  pytest -m gpu."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_three_node_known_answer(pkg):
    from dopf_b200.ptdf import ptdf_device
    # three_node.jl: L1 N2->N1 b 1, L2 N3->N1 b 1, L3 N2->N3 b 2, slack N3
    out = ptdf_device(3, [1, 2, 1], [0, 0, 2], [1, 1, 2], 2)
    np.testing.assert_allclose(out, [[-0.4, 0.2, 0.0], [-0.6, -0.2, 0.0], [0.4, 0.8, 0.0]], atol=1e-14)


@pytest.mark.parametrize("N,L,slack", [(40, 60, 0), (300, 450, 17), (2000, 3000, 0)])
def test_matches_oracle_and_host(pkg, oracle_mod, N, L, slack):
    from dopf_b200.ptdf import ptdf_device, ptdf_from_arrays
    d = pkg.cases.synthetic_arrays(N=N, L=L, G=4, S=1, T=2, seed=N)
    fr, to, b = d["line_from"], d["line_to"], d["susceptance"]
    dev = ptdf_device(N, fr, to, b, slack)
    host = ptdf_from_arrays(N, fr, to, b, slack)
    np.testing.assert_allclose(dev, host, rtol=0, atol=1e-10)
    assert not dev[:, slack].any()
    if N <= 300:
        np.testing.assert_allclose(dev, oracle_mod.ptdf(N, fr, to, b, slack), rtol=0, atol=1e-10)
    # a case built on the device PTDF runs like one built on the host PTDF
    d2 = pkg.cases.synthetic_arrays(N=N, L=L, G=4, S=1, T=2, seed=N, ptdf_fn=ptdf_device)
    np.testing.assert_allclose(d2["ptdf"], d["ptdf"] if slack == 0 else ptdf_from_arrays(N, fr, to, b, 0), atol=1e-10)


def test_disconnected_grid_is_reported(pkg):
    from dopf_b200.ptdf import ptdf_device
    with pytest.raises(RuntimeError, match="not connected|positive definite"):
        ptdf_device(4, [0, 2], [1, 3], [1.0, 1.0], 0)      # two islands

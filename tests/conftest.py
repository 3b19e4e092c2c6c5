import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    return g.load_package()


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def three_node_sorted(pkg):
    """three_node case with generators ordered by node (pv, gas @N1, wind @N2, coal @N3) + the
    permutation back to the reference order (pv, wind, coal, gas) used by the golden traces."""
    nodes, gens, stos, lines = pkg.cases.three_node()
    order = sorted(range(len(gens)), key=lambda i: nodes.index(gens[i].node))
    prob = pkg.Problem.from_structs(nodes, [gens[i] for i in order], stos, lines)
    return prob, order


GOLDEN = ["TNS", "big_gamma", "wrong_weight"]


def load_golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))

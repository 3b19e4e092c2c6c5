"""Case data: the reference's three-node system and the synthetic grids of SURVEY.md section 8(d)."""
import numpy as np

from .structures import Generator, Line, Node, Storage


def three_node():
    """/root/reference/src/cases/three_node.jl:1-22 -> (nodes, generators, storages, lines)."""
    node1 = Node("N1", [10, 250], False)
    node2 = Node("N2", [50, 70], False)
    node3 = Node("N3", [120, 200], True)
    nodes = [node1, node2, node3]
    lines = [Line("L1", node2, node1, 20, 1), Line("L2", node3, node1, 45, 1), Line("L3", node2, node3, 70, 2)]
    generators = [Generator("pv", 3, 80, "yellow", node1), Generator("wind", 4, 120, "lightblue", node2),
                  Generator("coal", 30, 300, "brown", node3), Generator("gas", 50, 120, "grey", node1)]
    storages = [Storage("battery", 1, 10, 20, "purple", node1)]
    return nodes, generators, storages, lines


def synthetic_arrays(N, L, G, S, T, seed=0, ptdf_fn=None, congest_frac=0.05, gamma=None):
    """Synthetic case as SoA arrays: random connected grid (spanning tree +
    chords, node 0 slack), integer-valued costs/capacities/demands like the reference's Int
    structs.  Returns a dict accepted by `Problem.from_arrays`.  `ptdf_fn(N, from, to, b, slack)`
    builds the PTDF (defaults to the host implementation in ptdf.py)."""
    from .ptdf import ptdf_from_arrays
    rng = np.random.default_rng(seed)
    assert L >= N - 1
    fr = np.empty(L, dtype=np.int32); to = np.empty(L, dtype=np.int32)
    perm = rng.permutation(N)
    for i in range(1, N):  # random spanning tree
        fr[i - 1] = perm[rng.integers(0, i)]; to[i - 1] = perm[i]
    for j in range(N - 1, L):  # chords
        a = int(rng.integers(0, N)); b = int(rng.integers(0, N - 1))
        if b >= a:
            b += 1
        fr[j] = a; to[j] = b
    susc = rng.integers(1, 6, size=L).astype(np.float64)
    ptdf = (ptdf_fn or ptdf_from_arrays)(N, fr, to, susc, 0)
    gen_node = np.sort(rng.integers(0, N, size=G)).astype(np.int32)
    gen_mc = rng.integers(1, 61, size=G).astype(np.float64)
    gen_pmax = rng.integers(10, 301, size=G).astype(np.float64)
    sto_node = np.sort(rng.integers(0, N, size=S)).astype(np.int32)
    sto_mc = np.ones(S)
    sto_pmax = rng.integers(5, 51, size=S).astype(np.float64)
    sto_emax = sto_pmax * rng.integers(2, 5, size=S)
    tot_cap = gen_pmax.sum()
    base = rng.integers(0, max(2, int(2 * tot_cap / (3 * N))) + 1, size=N).astype(np.float64)
    shape = 0.6 + 0.4 * np.sin(np.pi * np.arange(T) / 24.0) ** 2
    demand = np.floor(base[:, None] * shape[None, :])
    scale = min(1.0, 0.7 * tot_cap / max(demand.sum(axis=0).max(), 1.0))
    demand = np.floor(demand * scale)
    # line limits from a proportional dispatch: every generator at the same loading factor
    load = demand.sum(axis=0) / tot_cap
    inj = -demand.copy()
    np.add.at(inj, gen_node, gen_pmax[:, None] * load[None, :])
    flow = np.abs(ptdf @ inj).max(axis=1)
    fmax = np.ceil(1.2 * flow) + 1.0
    tight = rng.random(L) < congest_frac
    fmax[tight] = np.maximum(1.0, np.ceil(0.9 * flow[tight]))
    if gamma is None:
        gamma = min(0.3, 1.2 / max(G + S, 1))
    return dict(N=N, L=L, T=T, G=G, S=S, ptdf=ptdf, fmax=fmax, demand=demand,
                gen_mc=gen_mc, gen_pmax=gen_pmax, gen_node=gen_node,
                sto_mc=sto_mc, sto_pmax=sto_pmax, sto_emax=sto_emax, sto_node=sto_node,
                line_from=fr, line_to=to, susceptance=susc, slack=0, gamma=float(gamma))


def synthetic_scenarios(N, L, G, S, T, n_scen, seed=0, first_scenario=0, **kw):
    """Batch of `n_scen` independent scenarios on ONE synthetic grid (BASELINE configs[3]): grid, PTDF,
    line limits and the placement / capacities of the agents come from scenario `seed`; every scenario draws its own
    demand (node base loads rescaled by U(0.7, 1.3) and a shifted diurnal phase) and its own generator costs; the batch holds
    the scenarios first_scenario .. first_scenario + n_scen - 1 of that sequence."""
    d = synthetic_arrays(N=N, L=L, G=G, S=S, T=T, seed=seed, **kw)
    dem = np.empty((n_scen, N, T)); gmc = np.empty((n_scen, G))
    for c in range(n_scen):
        rng = np.random.default_rng([1000003 + seed, first_scenario + c])      # scenario k is the same draw whichever rank holds it
        scale = rng.uniform(0.7, 1.3, size=N)[:, None]
        dem[c] = np.floor(np.roll(d["demand"], int(rng.integers(0, T)), axis=1) * scale)
        gmc[c] = rng.integers(1, 61, size=G)
    out = dict(d)
    out.update(n_scen=n_scen, demand=dem, gen_mc=gmc, gen_pmax=np.tile(d["gen_pmax"], (n_scen, 1)),
               sto_mc=np.tile(d["sto_mc"], (n_scen, 1)), sto_pmax=np.tile(d["sto_pmax"], (n_scen, 1)), sto_emax=np.tile(d["sto_emax"], (n_scen, 1)))
    return out

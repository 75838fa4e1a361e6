"""Test grids restated from the reference's unit-test fixtures (values, not code).

three_bus_grid: tests/cpp_unit_tests/math_solver/test_math_solver_common.hpp:62-262 (SteadyStateSolverTestGrid) and
test_math_solver_pf.hpp:37-83 (PFSolverTestGrid): closed-form solution u = {1.08<-1deg, 0.97<-4deg, 0.90<-37deg}.
"""
import numpy as np

import oracle_lib as orc

DEG = np.pi / 180.0
A = np.exp(2j * np.pi / 3)
A2 = A * A


def _rot(x):
    """ComplexValue<asym>{x} = (x, x a^2, x a)"""
    return np.array([x, x * A2, x * A])


def _tensor_sm(s, m):
    return np.array([[s, m, m], [m, s, m], [m, m, s]], dtype=complex)


def three_bus_grid(sym=True, const_z=False, diverge=False, singular=False):
    """Returns (MathGrid, expected output dict)."""
    vref, v0, v1, v2 = 1.1, 1.08, 0.97, 0.90
    shift_val = np.pi / 6
    u0 = v0 * np.exp(-1j * DEG)
    u1 = v1 * np.exp(-4j * DEG)
    u2 = v2 * np.exp(-37j * DEG)
    shift = np.exp(1j * shift_val)
    y0, ys0, y1 = 1.0 - 2.0j, 0.05 + 0.2j, 3.0 - 4.0j
    branch0_i_f = (u0 - u1) * y0 + u0 * ys0
    branch0_i_t = (u1 - u0) * y0 + u1 * ys0
    branch1_i_f = (u1 - u2 * shift) * y1
    branch1_i_t = (u2 - u1 * np.conj(shift)) * y1
    uref, yref = vref, 10.0 - 50.0j
    source_inj = yref * (uref - u0)
    i0_load_inj = -source_inj + branch0_i_f
    i1_load_inj = branch0_i_t + branch1_i_f
    s0_load_inj = np.conj(i0_load_inj) * u0
    s1_load_inj = np.conj(i1_load_inj) * u1
    i2_shunt_inj = branch1_i_t
    ys = -i2_shunt_inj / u2

    B = 1 if sym else 3
    if sym:
        branch_param = np.array([[y0 + ys0, -y0, -y0, y0 + ys0], [y1, -y1 * shift, -y1 * np.conj(shift), y1]]).reshape(2, 4, 1, 1)
        shunt_param = np.array([ys]).reshape(1, 1, 1)
    else:
        y0_0 = 0.5 + 0.5j
        y0a = _tensor_sm(2.0 * y0 + y0_0, y0_0 - y0) / 3.0
        ys0a = _tensor_sm(ys0, 0.0)
        y1_1 = _tensor_sm(2.0 * y1, -y1) / 3.0
        y1_3 = np.array([[-y1, y1, 0.0], [0.0, -y1, y1], [y1, 0.0, -y1]], dtype=complex) / np.sqrt(3.0)
        branch_param = np.array([[y0a + ys0a, -y0a, -y0a, y0a + ys0a], [y1_1, y1_3, y1_3.T, y1_1]])
        ys_0 = ys * 0.2
        shunt_param = (_tensor_sm(2.0 * ys + ys_0, ys_0 - ys) / 3.0).reshape(1, 3, 3)
    source_param = np.array([[yref, yref]])
    sym_s_inj = np.array([
        s0_load_inj / 3.0, s0_load_inj / 3.0 / v0, s0_load_inj / 3.0 / v0 / v0,
        s1_load_inj / 3.0, s1_load_inj / 3.0 / v1, s1_load_inj / 3.0 / v1 / v1, 0.0,
    ])
    if const_z:
        z = np.zeros_like(sym_s_inj)
        z[2] = sym_s_inj[2] * 3.0
        z[5] = sym_s_inj[5] * 3.0
        sym_s_inj = z
    if diverge:
        sym_s_inj[6] = 1e6
    s_injection = sym_s_inj.reshape(-1, 1) if sym else np.repeat(sym_s_inj.reshape(-1, 1), 3, axis=1)
    if singular:
        branch_param = np.zeros_like(branch_param)
        shunt_param = np.zeros_like(shunt_param)
    grid = orc.MathGrid(
        sym=sym, phase_shift=[0.0, 0.0, -shift_val], branch_bus_idx=[[0, 1], [1, 2]],
        sources_per_bus=[0, 1, 1, 1], shunts_per_bus=[0, 0, 0, 1], load_gens_per_bus=[0, 3, 6, 7],
        load_gen_type=[0, 2, 1, 0, 2, 1, 0], branch_param=branch_param, shunt_param=shunt_param,
        source_param=source_param, source_u_ref=[vref], s_injection=s_injection,
    )

    # expected (sym values; asym = rotated phasors for u/i, repeated for s)
    def ph(x):  # phasor-like quantities
        return np.array([x]) if sym else _rot(x)

    def pw(x):  # power-like quantities
        return np.array([x]) if sym else np.array([x, x, x])

    branch = np.array([
        [pw(np.conj(branch0_i_f) * u0), pw(np.conj(branch0_i_t) * u1), ph(branch0_i_f), ph(branch0_i_t)],
        [pw(np.conj(branch1_i_f) * u1), pw(np.conj(branch1_i_t) * u2), ph(branch1_i_f), ph(branch1_i_t)],
    ])
    lg_s = [s0_load_inj / 3.0] * 3 + [s1_load_inj / 3.0] * 3 + [0.0]
    lg_i = [i0_load_inj / 3.0] * 3 + [i1_load_inj / 3.0] * 3 + [0.0]
    if const_z:
        lg_s = [x * 3.0 if k % 3 == 2 and k < 6 else 0.0 for k, x in enumerate(lg_s)]
        lg_i = [x * 3.0 if k % 3 == 2 and k < 6 else 0.0 for k, x in enumerate(lg_i)]
    expected = {
        "u": np.array([ph(u0), ph(u1), ph(u2)]),
        "branch": branch,
        "source": np.array([[pw(np.conj(source_inj) * u0), ph(source_inj)]]),
        "shunt": np.array([[pw(np.conj(i2_shunt_inj) * u2), ph(i2_shunt_inj)]]),
        "load_gen": np.array([[pw(s), ph(i)] for s, i in zip(lg_s, lg_i)]),
        "bus_injection": np.array([pw(np.conj(branch0_i_f) * u0), pw(np.conj(branch0_i_t) * u1 + np.conj(branch1_i_f) * u1), pw(0.0)]),
    }
    return grid, expected


def random_grid(n_node, n_extra_edges=0, seed=0, n_shunt=2, n_source=1):
    """Random symmetric math grid: a random tree plus `n_extra_edges` mesh edges, ordered by the (pinned) oracle
    symbolic stage so that radial grids have no fill-ins and meshed ones carry the reference's fill-in list."""
    rng = np.random.default_rng(seed)
    edges = [(int(rng.integers(0, i)), i) for i in range(1, n_node)]
    while len(edges) < n_node - 1 + n_extra_edges:
        a, b = (int(x) for x in rng.integers(0, n_node, 2))
        if a != b:
            edges.append((a, b))
    n_branch = len(edges)
    src_nodes = [0] + [int(x) for x in rng.integers(1, n_node, n_source - 1)]
    shunt_nodes = [int(x) for x in rng.integers(0, n_node, n_shunt)]
    lg_nodes = [int(x) for x in rng.integers(1, n_node, int(1.5 * n_node))]
    lg_types = rng.integers(0, 3, len(lg_nodes))
    bag = orc.topology(n_node, edges, [[1, 1]] * n_branch, [0.0] * n_branch, src_nodes, [1] * n_source,
                       shunt_node_idx=shunt_nodes, load_gen_node_idx=lg_nodes, load_gen_type=lg_types)
    assert bag.i64("n_math")[0] == 1
    n_lg = len(lg_nodes)
    # parameters in component order, then permuted into math order through the coupling
    z = rng.uniform(0.002, 0.02, n_branch) + 1j * rng.uniform(0.005, 0.05, n_branch)
    ys = 1.0 / z
    ysh = 1j * rng.uniform(0.0, 1e-3, n_branch)
    tap = np.where(rng.random(n_branch) < 0.1, rng.uniform(0.95, 1.05, n_branch), 1.0)
    bp_comp = np.stack([(ys + 0.5 * ysh) / tap**2, -ys / tap, -ys / tap, ys + 0.5 * ysh], axis=1)
    coup_branch = bag.i64("coup.branch").reshape(-1, 2)
    bp = np.zeros_like(bp_comp)
    bp[coup_branch[:, 1]] = bp_comp
    sh_comp = rng.uniform(0, 0.01, n_shunt) + 1j * rng.uniform(-0.02, 0.02, n_shunt)
    sh = np.zeros_like(sh_comp)
    sh[bag.i64("coup.shunt").reshape(-1, 2)[:, 1]] = sh_comp
    src_param = np.tile(np.array([[1.0 / (0.0005 + 0.005j), 1.0 / (0.001 + 0.01j)]]), (n_source, 1))
    grid = orc.MathGrid(
        sym=True, phase_shift=bag.f64("g0.phase_shift"), branch_bus_idx=bag.i64("g0.branch_bus_idx"),
        sources_per_bus=bag.i64("g0.sources_per_bus"), shunts_per_bus=bag.i64("g0.shunts_per_bus"),
        load_gens_per_bus=bag.i64("g0.load_gens_per_bus"), load_gen_type=bag.i64("g0.load_gen_type"),
        branch_param=bp, shunt_param=sh, source_param=src_param, source_u_ref=[1.02] * n_source,
        s_injection=np.zeros(n_lg), fill_in=bag.i64("g0.fill_in"),
    )
    return grid


def random_scenarios(grid, n_scn, seed=0, scale=0.6):
    """(n_scn, n_load_gen) complex per-unit injections (loads negative) and (n_scn, n_source) reference voltages"""
    rng = np.random.default_rng(seed + 1000)
    n_lg = len(grid.load_gen_type)
    p = -rng.uniform(0.0, scale / max(n_lg, 1), (n_scn, n_lg)) * rng.uniform(0.0, 2.0, (n_scn, 1))
    q = p * rng.uniform(-0.2, 0.6, (n_scn, n_lg))
    gen = rng.random((n_scn, n_lg)) < 0.1
    s = np.where(gen, -0.5 * (p + 1j * q), p + 1j * q)
    u_ref = rng.uniform(0.98, 1.06, (n_scn, len(grid.source_u_ref))) * np.exp(1j * rng.uniform(-0.05, 0.05, (n_scn, 1)))
    return s, u_ref


def regulated_benchmark_grid(rings, n_gen=12, seed=3, q_lim=None, option=None, n_scn=8):
    """Fictional benchmark grid plus `n_gen` const_pq generators under voltage regulators (PV buses): two buses carry two
    regulated generators each (shared u_ref), some regulators have one-sided limits; the batch switches generators off in some
    scenarios and varies one generator's P on top of the load profile.  Returns (FictionalGrid, input_data, update_data)."""
    import pgm_b200
    from pgm_b200 import structs

    option = option or pgm_b200.BENCHMARK_OPTION
    grid = pgm_b200.FictionalGrid(seed=0, has_mv_ring=rings, has_lv_ring=rings, **option)
    inp = dict(grid.input_data)
    rng = np.random.default_rng(seed)
    nodes = rng.choice(inp["sym_load"]["node"], size=n_gen, replace=False)
    nodes[-2:] = nodes[:2]
    max_id = max(int(a["id"].max()) for a in inp.values() if len(a))
    gen = structs.initialize_array("input", "sym_gen", n_gen)
    gen["id"] = max_id + 1 + np.arange(n_gen)
    gen["node"] = nodes
    gen["status"] = 1
    gen["type"] = 0
    gen["p_specified"] = rng.uniform(2e3, 2e4, n_gen)
    gen["q_specified"] = rng.uniform(-5e3, 5e3, n_gen)
    reg = structs.initialize_array("input", "voltage_regulator", n_gen)
    reg["id"] = max_id + 1 + n_gen + np.arange(n_gen)
    reg["regulated_object"] = gen["id"]
    reg["status"] = 1
    reg["u_ref"] = rng.uniform(0.99, 1.03, n_gen)
    reg["u_ref"][-2:] = reg["u_ref"][:2]
    if q_lim is not None:
        reg["q_min"] = -q_lim * rng.uniform(0.5, 1.5, n_gen)
        reg["q_max"] = q_lim * rng.uniform(0.5, 1.5, n_gen)
        reg["q_max"][2] = np.nan
        reg["q_min"][3] = np.nan
    inp["sym_gen"] = gen
    inp["voltage_regulator"] = reg
    upd = grid.batch_update(n_scn, seed=0)
    g = structs.initialize_array("update", "sym_gen", (n_scn, n_gen))
    g["id"] = gen["id"]
    g["status"][1::3, 4] = 0
    g["status"][2::4, -1] = 0
    g["p_specified"][:, 5] = rng.uniform(1e3, 3e4, n_scn)
    upd["sym_gen"] = g
    return grid, inp, upd

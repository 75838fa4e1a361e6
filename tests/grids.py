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

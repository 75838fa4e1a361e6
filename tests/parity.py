"""Comparison of output datasets of the CUDA path with the oracle's at the north_star tolerances (BASELINE.json):
iteration counts equal, node voltages within 1e-9 pu, currents and powers within 1e-6 relative.

Test infrastructure (used by tests/, bench.py's parity check of the timed batch and __graft_entry__.smoke()).
"relative" needs a floor for quantities that are physically zero (an open branch end, a const-power load's reactive part):
powers are compared relative to max(|ref|, 1 kW) (1e-3 pu at the 1 MVA base), currents relative to max(|ref|, 0.1 A), so a
difference of 1e-9 pu in a vanishing quantity does not count as a mismatch while every loaded element is held to 1e-6.
"""
import numpy as np

U_TOL = 1e-9      # pu (and rad for angles)
REL_TOL = 1e-6
P_FLOOR = 1e3     # W / var / VA
I_FLOOR = 0.1     # A

_POWER = ("p", "q", "s", "p_from", "q_from", "s_from", "p_to", "q_to", "s_to")
_CURRENT = ("i", "i_from", "i_to")


def compare_outputs(res, ref, components=None, check_ids=True):
    """res / ref: dict component -> structured array of identical shape.  Returns {"max_du_pu": .., "max_rel": ..};
    raises AssertionError naming the first attribute out of tolerance."""
    worst_u, worst_rel = 0.0, 0.0
    for c in components if components is not None else [k for k in res if isinstance(res[k], np.ndarray) and res[k].dtype.names]:
        a, b = res[c], ref[c]
        assert a.shape == b.shape, (c, a.shape, b.shape)
        if a.size == 0:
            continue
        names = a.dtype.names
        if check_ids:
            assert np.array_equal(a["id"], b["id"]), f"{c}.id"
            assert np.array_equal(a["energized"], b["energized"]), f"{c}.energized"
        for n in names:
            x, y = a[n], b[n]
            if n in ("id", "energized"):
                continue
            if x.dtype.kind in "iu":
                assert np.array_equal(x, y), f"{c}.{n}"
                continue
            both_nan = np.isnan(x) & np.isnan(y)
            assert np.array_equal(np.isnan(x), np.isnan(y)), f"{c}.{n}: NaN pattern differs"
            x = np.where(both_nan, 0.0, x)
            y = np.where(both_nan, 0.0, y)
            if n == "u_pu":
                d = float(np.max(np.abs(x - y)))
                worst_u = max(worst_u, d)
                assert d <= U_TOL, f"{c}.u_pu differs by {d:.3e} pu"
            elif n == "u":
                d = float(np.max(np.abs(x - y) / np.maximum(np.abs(y), 1e-300) * (np.abs(y) > 0)))
                assert d <= 1e-8, f"{c}.u relative difference {d:.3e}"
            elif n == "u_angle":
                dphi = np.abs(np.angle(np.exp(1j * (x - y))))
                live = np.asarray(b["u_pu"]) > 1e-6
                d = float(np.max(dphi * live))
                worst_u = max(worst_u, d)
                assert d <= U_TOL * 10, f"{c}.u_angle differs by {d:.3e} rad"
            elif n in _POWER:
                d = float(np.max(np.abs(x - y) / np.maximum(np.abs(y), P_FLOOR)))
                worst_rel = max(worst_rel, d)
                assert d <= REL_TOL, f"{c}.{n} relative difference {d:.3e}"
            elif n in _CURRENT:
                d = float(np.max(np.abs(x - y) / np.maximum(np.abs(y), I_FLOOR)))
                worst_rel = max(worst_rel, d)
                assert d <= REL_TOL, f"{c}.{n} relative difference {d:.3e}"
            elif n == "pf":
                # p / s: ill-conditioned where s vanishes; compare where the element carries power
                s_ref = np.asarray(b["s"])
                d = float(np.max(np.abs(x - y) * (s_ref > P_FLOOR)))
                assert d <= 1e-5, f"{c}.pf differs by {d:.3e}"
            else:  # loading and anything else: relative with a floor of 1e-3
                d = float(np.max(np.abs(x - y) / np.maximum(np.abs(y), 1e-3)))
                worst_rel = max(worst_rel, d)
                assert d <= REL_TOL, f"{c}.{n} relative difference {d:.3e}"
    return {"max_du_pu": worst_u, "max_rel": worst_rel}


def compare_batch(res, n_iter, status, ref, components=None):
    """full parity of one batch: status and iteration counts equal for every scenario, then compare_outputs"""
    ref_status = np.asarray(ref["status"])
    assert np.array_equal(np.asarray(status) != 0, ref_status != 0), "failure pattern differs"
    ok = ref_status == 0
    assert np.array_equal(np.asarray(n_iter)[ok], np.asarray(ref["n_iter"])[ok]), "iteration counts differ"
    comps = components if components is not None else [k for k in res if k in ref and isinstance(res[k], np.ndarray) and res[k].dtype.names]
    if ok.all():
        out = compare_outputs(res, ref, comps)
    else:
        out = compare_outputs({c: res[c][ok] for c in comps}, {c: ref[c][ok] for c in comps}, comps)
    out["scenarios"] = int(len(ref_status))
    out["n_iter_equal"] = True
    return out

"""Host side of the automatic tap changer (model_tap.cpp) without a GPU: the ranking of the regulated transformers against
the expectations of the reference's unit tests (tests/cpp_unit_tests/optimizer/test_tap_position_optimizer.cpp:499-760) and the
checks at model creation (main_core/input.hpp:168-241).  `tap_rank` rows: (kind: 0 transformer / 1 three-winding transformer,
index within the kind, rank group)."""
import numpy as np
import pytest

import pgm_b200
from pgm_b200.structs import initialize_array

FROM, TO = 0, 1


def _nodes(ids_u):
    a = initialize_array("input", "node", len(ids_u))
    a["id"], a["u_rated"] = [i for i, _ in ids_u], [u for _, u in ids_u]
    return a


def _trafos(rows):  # (id, from, to, tap_side[, from_status])
    a = initialize_array("input", "transformer", len(rows))
    for k, r in enumerate(rows):
        a["id"][k], a["from_node"][k], a["to_node"][k], a["tap_side"][k] = r[:4]
        a["from_status"][k] = r[4] if len(r) > 4 else 1
    a["to_status"], a["winding_from"], a["winding_to"], a["clock"] = 1, 1, 1, 0
    a["tap_min"], a["tap_max"], a["tap_nom"], a["tap_pos"] = -128 + 1, 127, 0, 0
    return a


def _trafo3w(id_, n1, n2, n3, tap_side):
    a = initialize_array("input", "three_winding_transformer", 1)
    a["id"], a["node_1"], a["node_2"], a["node_3"], a["tap_side"] = id_, n1, n2, n3, tap_side
    a["status_1"] = a["status_2"] = a["status_3"] = 1
    a["winding_1"] = a["winding_2"] = a["winding_3"] = 1
    a["clock_12"] = a["clock_13"] = 0
    a["tap_min"], a["tap_max"], a["tap_nom"], a["tap_pos"] = -127, 127, 0, 0
    return a


def _branches(kind, rows):
    a = initialize_array("input", kind, len(rows))
    for k, (i, f, t) in enumerate(rows):
        a["id"][k], a["from_node"][k], a["to_node"][k] = i, f, t
    a["from_status"] = a["to_status"] = 1
    return a


def _source(id_, node):
    a = initialize_array("input", "source", 1)
    a["id"], a["node"], a["status"], a["u_ref"] = id_, node, 1, 1.0
    return a


def _regulators(rows):  # (id, regulated_object, control_side)
    a = initialize_array("input", "transformer_tap_regulator", len(rows))
    for k, (i, obj, side) in enumerate(rows):
        a["id"][k], a["regulated_object"][k], a["control_side"][k] = i, obj, side
    a["status"] = 1
    return a


def _rank(data):
    return pgm_b200.PowerGridModel(data).math_index(0, "tap_rank").reshape(-1, 3).tolist()


def test_ranking_full_grid_2():
    """transformers behind links / lines / a three-winding transformer; 103 hangs on an open switch and 104 on an island
    without a source: neither is ranked"""
    data = {
        "node": _nodes([(0, 150e3), (1, 10e3), (2, 10e3), (3, 10e3), (4, 10e3), (5, 50e3), (6, 10e3), (7, 10e3), (8, 10e3), (9, 10e3),
                        (10, 10e3), (100, 10e3), (101, 10e3), (102, 10e3)]),
        "transformer": _trafos([(11, 0, 1, TO), (12, 0, 1, FROM), (13, 2, 3, FROM), (14, 6, 7, FROM), (15, 5, 8, FROM), (16, 9, 10, FROM),
                                (103, 9, 100, FROM, 0), (104, 101, 102, FROM)]),
        "three_winding_transformer": _trafo3w(17, 0, 4, 5, 1),
        "line": _branches("line", [(18, 4, 6), (19, 3, 10)]),
        "link": _branches("link", [(20, 1, 2), (21, 3, 7), (22, 8, 9)]),
        "source": _source(23, 0),
        "transformer_tap_regulator": _regulators([(24, 11, 1), (25, 12, 1), (26, 13, 1), (27, 14, 1), (28, 15, 1), (29, 16, 1), (30, 17, 1),
                                                  (105, 103, 0), (106, 104, 0)]),
    }
    assert _rank(data) == [[0, 0, 0], [0, 1, 0], [1, 0, 0], [0, 4, 0], [0, 2, 1], [0, 3, 1], [0, 5, 1]]


def test_ranking_meshed_grid_gives_the_bypassed_transformer_the_last_rank():
    data = {
        "node": _nodes([(0, 10e3), (1, 10e3), (2, 10e3), (3, 10e3)]),
        "transformer": _trafos([(11, 0, 1, TO), (12, 1, 2, FROM), (13, 2, 3, FROM)]),
        "line": _branches("line", [(21, 0, 2)]),
        "source": _source(31, 0),
        "transformer_tap_regulator": _regulators([(41, 11, 1), (42, 12, 1), (43, 13, 1)]),
    }
    assert _rank(data) == [[0, 0, 0], [0, 2, 0], [0, 1, 1]]


def test_regulating_towards_the_source_is_refused():
    data = {
        "node": _nodes([(0, 150e3), (1, 10e3), (2, 10e3), (3, 10e3), (4, 10e3), (5, 10e3)]),
        "transformer": _trafos([(10, 0, 1, FROM), (20, 1, 2, TO), (30, 1, 3, TO), (40, 2, 4, FROM), (50, 3, 5, FROM)]),
        "source": _source(60, 0),
        "transformer_tap_regulator": _regulators([(70, 10, 1), (71, 20, 0), (72, 30, 0), (73, 40, 1), (74, 50, 1)]),
    }
    with pytest.raises(pgm_b200.PgmB200Error, match="controlled from non-source side towards source side:\n  Transformer IDs: 20, 30$"):
        _rank(data)
    data = {
        "node": _nodes([(0, 150e3), (1, 10e3), (2, 10e3), (3, 10e3)]),
        "three_winding_transformer": _trafo3w(10, 0, 1, 2, 0),
        "source": _source(20, 0),
        "transformer_tap_regulator": _regulators([(30, 10, 0)]),
    }
    with pytest.raises(pgm_b200.PgmB200Error) as err:
        _rank(data)
    assert str(err.value) == ("Automatic tap changer has invalid configuration. The following transformer(s) are being controlled "
                              "from non-source side towards source side:\n  Transformer IDs: 10")


def test_a_switched_off_regulator_leaves_its_transformer_unranked():
    data = {
        "node": _nodes([(0, 150e3), (1, 10e3), (2, 10e3)]),
        "transformer": _trafos([(10, 0, 1, FROM), (11, 1, 2, FROM)]),
        "source": _source(20, 0),
        "transformer_tap_regulator": _regulators([(30, 10, 1), (31, 11, 1)]),
    }
    assert _rank(data) == [[0, 0, 0], [0, 1, 1]]
    data["transformer_tap_regulator"]["status"][0] = 0
    assert _rank(data) == [[0, 1, 0]]  # transformer 10 is a plain branch now: node 1 is as close to the source as node 0


def test_regulator_checks_at_model_creation():
    base = {
        "node": _nodes([(0, 150e3), (1, 10e3)]),
        "transformer": _trafos([(10, 0, 1, FROM)]),
        "line": _branches("line", [(15, 1, 1)]),
        "source": _source(20, 0),
    }
    with pytest.raises(pgm_b200.PgmB200Error, match="There are objects regulated by more than one regulator"):
        pgm_b200.PowerGridModel({**base, "transformer_tap_regulator": _regulators([(30, 10, 1), (31, 10, 1)])})
    with pytest.raises(pgm_b200.PgmB200Error, match="regulator is not supported for object with ID 15"):
        pgm_b200.PowerGridModel({**base, "transformer_tap_regulator": _regulators([(30, 15, 1)])})
    with pytest.raises(pgm_b200.PgmB200Error, match="The id cannot be found: 99"):
        pgm_b200.PowerGridModel({**base, "transformer_tap_regulator": _regulators([(30, 99, 1)])})
    with pytest.raises(pgm_b200.PgmB200Error, match="ControlSide #2"):
        pgm_b200.PowerGridModel({**base, "transformer_tap_regulator": _regulators([(30, 10, 2)])})
    model = pgm_b200.PowerGridModel({**base, "transformer_tap_regulator": _regulators([(30, 10, 1)])})
    assert model.math_index(0, "tap_rank").tolist() == [0, 0, 0]


@pytest.mark.parametrize("seed", range(60))
def test_ranking_equals_the_oracle_on_random_grids(seed):
    """random grids (tree + a few loops) of lines, links, transformers and three-winding transformers with regulators on random
    sides, some switched off, some branches open, one or two sources: the product's ranking (model_tap.cpp) and the oracle's
    (oracle/tap_optimizer.hpp) give the same groups in the same order, or both refuse the configuration"""
    import oracle_lib as orc

    rng = np.random.default_rng(1000 + seed)
    n_node = int(rng.integers(6, 40))
    data = {"node": _nodes([(i, 10e3) for i in range(n_node)])}
    next_id = [n_node]

    def new_id():
        next_id[0] += 1
        return next_id[0]

    edges = [(int(rng.integers(0, i)), i) for i in range(1, n_node)]
    for _ in range(int(rng.integers(0, 4))):
        a, b = (int(x) for x in rng.integers(0, n_node, 2))
        if a != b:
            edges.append((a, b))
    kinds = rng.choice(["line", "link", "transformer"], size=len(edges), p=[0.4, 0.15, 0.45])
    lines, links, trafos, regs = [], [], [], []
    for (a, b), kind in zip(edges, kinds):
        i = new_id()
        if kind == "line":
            lines.append((i, a, b))
        elif kind == "link":
            links.append((i, a, b))
        else:
            trafos.append((i, a, b, int(rng.integers(0, 2)), int(rng.random() > 0.1)))
            if rng.random() < 0.7:
                regs.append((new_id(), i, 1 if rng.random() < 0.9 else 0))  # mostly away from the source (node 0 feeds the tree)
    t3w = None
    if n_node >= 8 and rng.random() < 0.4:
        n1, n2, n3 = sorted(int(x) for x in rng.choice(n_node, 3, replace=False))
        t3w = _trafo3w(new_id(), n1, n2, n3, int(rng.integers(0, 3)))
        if rng.random() < 0.8:
            regs.append((new_id(), int(t3w["id"][0]), int(rng.integers(1, 3))))
    if lines:
        data["line"] = _branches("line", lines)
        data["line"]["from_status"][rng.random(len(lines)) < 0.1] = 0
    if links:
        data["link"] = _branches("link", links)
    if trafos:
        data["transformer"] = _trafos(trafos)
    if t3w is not None:
        data["three_winding_transformer"] = t3w
    sources = [_source(new_id(), 0)]
    if rng.random() < 0.3:
        sources.append(_source(new_id(), int(rng.integers(1, n_node))))
        sources[-1]["status"] = int(rng.random() < 0.7)
    data["source"] = np.concatenate(sources)
    if regs:
        data["transformer_tap_regulator"] = _regulators(regs)
        data["transformer_tap_regulator"]["status"][rng.random(len(regs)) < 0.2] = 0
    oracle = orc.Model(data)
    try:
        expected = oracle.tap_rank().tolist()
    except RuntimeError as err:
        with pytest.raises(pgm_b200.PgmB200Error) as mine:
            _rank(data)
        assert str(mine.value) == str(err)
        return
    assert _rank(data) == expected

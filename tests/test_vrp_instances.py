"""Host side of the CVRP instance files (tpc_b200/vrp_instances.py): TSPLIB parser, solution parser, the tables and the
per-episode sampling of `environment/custom/vrp/env.py:148-206,249-268`. With the reference checkout present (this
container only) the parser is also pinned to the reference's own pre-parsed JSON of every instance it ships."""
import json
import os

import numpy as np
import pytest

from tpc_b200 import vrp_instances as V

REF_DATA = "/root/reference/environment/custom/vrp/datasets"

MINI = """NAME : T-n6-k2
COMMENT : (hand made, Min no of trucks: 2, Optimal value: 99)
TYPE : CVRP
DIMENSION : 6
EDGE_WEIGHT_TYPE : EUC_2D
CAPACITY : 10
NODE_COORD_SECTION
 1 50 50
 2 10 20
 3 30 40
 4 70 80
 5 90 10
 6 20 90
DEMAND_SECTION
1 0
2 4
3 3
4 5
5 2
6 6
DEPOT_SECTION
 1
 -1
EOF
"""


def _write_mini(d):
    with open(os.path.join(d, "T-n6-k2.vrp"), "w") as f:
        f.write(MINI)
    with open(os.path.join(d, "opt-T-n6-k2"), "w") as f:
        f.write("Route #1: 1 2 4\nRoute #2: 3 5\ncost 99\n")


def test_parser_and_tables_on_a_hand_made_instance(tmp_path):
    _write_mini(tmp_path)
    prob, sol = V.load_problem(str(tmp_path), "T-n6-k2")
    assert prob["name"] == "T-n6-k2" and prob["dimension"] == 6 and prob["capacity"] == 10 and prob["type"] == "CVRP"
    assert prob["edge_weight_type"] == "EUC_2D" and prob["total_demand"] == 20
    assert prob["comment"] == "(hand made, Min no of trucks: 2, Optimal value: 99)"
    assert [n["type"] for n in prob["nodes"]] == ["depot"] + ["node"] * 5
    assert prob["nodes"][3] == {"id": 4, "x": 70, "y": 80, "demand": 5, "type": "node"}
    assert sol == {"cost": 99, "num_routes": 2, "routes": [[1, 2, 4], [3, 5]]}
    vehicles, nodes, eos = V.instance_tables(prob)
    assert vehicles.shape == (2, 3) and np.all(vehicles == np.array([50, 50, 10], np.float32))   # ceil(20 / 10) vehicles
    assert nodes.shape == (6, 3) and np.array_equal(nodes[5], np.array([20, 90, 6], np.float32))
    assert np.array_equal(eos, np.array([[50, 50, 0]], np.float32))
    assert V.load_problem(str(tmp_path), "no-such-instance") is None and V.load_problem(str(tmp_path / "nope"), "x") is None


def test_json_entry_wins_over_the_text_file(tmp_path):
    _write_mini(tmp_path)
    prob = V.parse_vrp(str(tmp_path / "T-n6-k2.vrp"))
    prob["capacity"] = 12                      # a marker: only the JSON carries it
    with open(tmp_path / "T-n6-k2.vrp.json", "w") as f:
        json.dump({"problem": prob, "solution": {"cost": 7, "num_routes": 0, "routes": []}}, f)
    got, sol = V.load_problem(str(tmp_path), "T-n6-k2")
    assert got["capacity"] == 12 and sol["cost"] == 7


def test_sampled_batch_layout(tmp_path):
    _write_mini(tmp_path)
    prob, _ = V.load_problem(str(tmp_path), "T-n6-k2")
    tables = V.instance_tables(prob)
    rng = np.random.default_rng(3)
    batch = V.sample_batch(tables, 16, 4, 2, rng)
    assert batch.shape == (16, 6, 3) and batch.dtype == np.float32
    assert np.all(batch[:, 0] == np.array([50, 50, 0], np.float32))              # the depot opens every episode
    assert np.all(batch[:, 4:] == np.array([50, 50, 10], np.float32))            # vehicles parked at the depot, full load
    customers = {tuple(r) for r in tables[1][1:].tolist()}
    for b in range(16):
        rows = [tuple(r) for r in batch[b, 1:4].tolist()]
        assert len(set(rows)) == 3 and set(rows) <= customers                    # distinct customers, never the depot
    assert len({tuple(map(tuple, batch[b, 1:4].tolist())) for b in range(16)}) > 1   # episodes differ
    again = V.sample_batch(tables, 16, 4, 2, np.random.default_rng(3))
    assert np.array_equal(batch, again)                                          # a seeded stream
    with pytest.raises(ValueError):
        V.sample_batch(tables, 1, 8, 2, rng)                                     # more customers than the instance has
    with pytest.raises(ValueError):
        V.sample_batch(tables, 1, 4, 3, rng)                                     # more vehicles than ceil(demand / capacity)


@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason="reference checkout not present")
def test_parser_equals_the_reference_json_of_every_shipped_instance():
    checked = 0
    for name in sorted(os.listdir(REF_DATA)):
        if not name.endswith(".vrp") or not os.path.exists(os.path.join(REF_DATA, name + ".json")):
            continue
        want = json.load(open(os.path.join(REF_DATA, name + ".json")))
        got = V.parse_vrp(os.path.join(REF_DATA, name))
        assert got == want["problem"], name
        sol = os.path.join(REF_DATA, "opt-" + name[:-4])
        if os.path.exists(sol):
            assert V.parse_solution(sol) == want["solution"], name
        checked += 1
    assert checked >= 20
    prob, sol = V.load_problem(REF_DATA, "A-n32-k5")                              # BASELINE config #5
    assert prob["dimension"] == 32 and prob["capacity"] == 100 and sol["cost"] == 784
    assert V.instance_tables(prob)[0].shape[0] == 5

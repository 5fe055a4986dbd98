"""Augerat / TSPLIB CVRP instance files on the host: what `environment/custom/vrp/datasets/parser.py` and
`CVRP.generate_dataset / generate_batch` (`environment/custom/vrp/env.py:67-146` is the step; :148-206 the batch, :249-268
the tables) do before the first device call.

The reference keeps, next to every `<name>.vrp`, a pre-parsed `<name>.vrp.json` = {"problem": {...}, "solution": {...}}
and its environment only ever reads the JSON (`parser.py:159-173`). `load_problem` here accepts either: the JSON when it
is there, else the `.vrp` text itself (+ an `opt-<name>` solution file when present). The instance files are not part of
this repository; point `env_config.dir_path` at a directory that holds them (e.g. the reference checkout's
`environment/custom/vrp/datasets`).

Everything here is NumPy / pure Python; the device side starts at `CVRP.load_batch` (env.py).
"""
from __future__ import annotations

import json
import math
import os

import numpy as np

NODE, DEPOT = "node", "depot"
_SECTIONS = ("NODE_COORD_SECTION", "DEMAND_SECTION", "DEPOT_SECTION")


def parse_vrp(path: str) -> dict:
    """One TSPLIB CVRP file -> the dictionary `parser.py:49-132` builds (same keys, same value types)."""
    spec = {"NAME": "", "COMMENT": "", "TYPE": "", "DIMENSION": -1, "EDGE_WEIGHT_TYPE": "", "CAPACITY": -1}
    coords: list[tuple[int, int, int]] = []
    demand: dict[int, int] = {}
    depots: list[int] = []
    section = None
    with open(path) as f:
        for raw in f:
            tok = raw.split()
            if not tok:
                continue
            if tok[-1] == "EOF":
                break
            if tok[0] in _SECTIONS:
                section = tok[0]
                continue
            key = tok[0].rstrip(":")
            if key in spec:                      # "KEY : value ..." header line
                rest = [t for t in tok[1:] if t != ":"]
                if key == "COMMENT":
                    spec[key] = " ".join(rest)
                elif key in ("DIMENSION", "CAPACITY"):
                    spec[key] = int(rest[-1])
                else:
                    spec[key] = rest[-1] if rest else ""
                continue
            if section == "NODE_COORD_SECTION":
                coords.append((int(tok[0]), int(tok[1]), int(tok[2])))
            elif section == "DEMAND_SECTION":
                demand[int(tok[0])] = int(tok[1])
            elif section == "DEPOT_SECTION" and tok[0] != "-1":
                depots.append(int(tok[0]))
    nodes = [{"id": i, "x": x, "y": y, "demand": demand.get(i, -1), "type": DEPOT if i in depots else NODE}
             for (i, x, y) in coords]
    return {"name": spec["NAME"], "comment": spec["COMMENT"], "type": spec["TYPE"], "dimension": spec["DIMENSION"],
            "edge_weight_type": spec["EDGE_WEIGHT_TYPE"], "capacity": spec["CAPACITY"],
            "total_demand": sum(demand.values()), "nodes": nodes}


def parse_solution(path: str) -> dict:
    """`opt-<name>` file (`Route #k: a b c` lines and a `cost N` line) -> {"cost", "num_routes", "routes"} (parser.py:24-47).
    Like the reference's parser this stops at the first blank line, so a file that opens with "Solution 1:" and a blank
    line (A-n45-k7) yields no routes and cost -1 -- which is what the reference's own JSON of that instance holds."""
    routes, cost = [], -1
    with open(path) as f:
        for raw in f:
            tok = raw.split()
            if not tok:
                break
            if tok[0] == "Route":
                routes.append([int(t) for t in tok[2:]])
            elif tok[0].lower() == "cost":
                cost = int(tok[-1])
    return {"cost": cost, "num_routes": len(routes), "routes": routes}


def load_problem(dir_path: str, prob_name: str):
    """(problem, solution) of the instance called `prob_name` under `dir_path` (parser.py:159-173), or None.
    `solution` is {"cost": -1, ...} when no optimum file sits next to a bare `.vrp`."""
    if not dir_path or not os.path.isdir(dir_path):
        return None
    names = sorted(os.listdir(dir_path))
    for n in names:
        if n.endswith(".json"):
            try:
                with open(os.path.join(dir_path, n)) as f:
                    entry = json.load(f)
            except (OSError, ValueError):
                continue
            if isinstance(entry, dict) and entry.get("problem", {}).get("name") == prob_name:
                return entry["problem"], entry.get("solution", {"cost": -1, "num_routes": 0, "routes": []})
    for n in names:
        if n.endswith(".vrp"):
            prob = parse_vrp(os.path.join(dir_path, n))
            if prob["name"] == prob_name:
                sol_path = os.path.join(dir_path, "opt-" + n[: -len(".vrp")])
                sol = parse_solution(sol_path) if os.path.exists(sol_path) else {"cost": -1, "num_routes": 0, "routes": []}
                return prob, sol
    return None


def instance_tables(problem: dict, num_features: int = 3):
    """(vehicles [V, F], nodes [n, F], depot row [1, F]) as `CVRP.generate_dataset` (vrp/env.py:249-268) builds them:
    V = ceil(total demand / capacity) vehicles parked at the depot with a full load, one row [x, y, demand] per node."""
    nodes = np.zeros((len(problem["nodes"]), num_features), dtype=np.float32)
    depot = None
    for i, nd in enumerate(problem["nodes"]):
        nodes[i, 0], nodes[i, 1], nodes[i, 2] = nd["x"], nd["y"], nd["demand"]
        if nd["type"] == DEPOT:
            depot = nd
    if depot is None:
        raise ValueError(f"instance {problem.get('name')!r} names no depot")
    n_vehicles = math.ceil(problem["total_demand"] / problem["capacity"])
    vehicles = np.zeros((n_vehicles, num_features), dtype=np.float32)
    vehicles[:, 0], vehicles[:, 1], vehicles[:, 2] = depot["x"], depot["y"], problem["capacity"]
    eos = np.array([[depot["x"], depot["y"], depot["demand"]]], dtype=np.float32)
    if num_features > 3:
        eos = np.pad(eos, ((0, 0), (0, num_features - 3)))
    return vehicles, nodes, eos


def sample_batch(tables, batch_size: int, node_sample_size: int, vehicle_sample_size: int, rng: np.random.Generator):
    """(B, node_sample_size + vehicle_sample_size, F) instance batch in the layout of `CVRP.generate_batch`
    (vrp/env.py:148-206): row 0 the depot, rows 1 .. Nn - 1 distinct customers drawn without replacement from the
    instance (the depot is node 0 of the file and never drawn), then the vehicles. The reference shuffles with NumPy's
    global generator; here the caller passes a seeded Generator (one stream per (seed, reset count))."""
    vehicles, nodes, eos = tables
    n_customers = nodes.shape[0] - 1
    if node_sample_size - 1 > n_customers:
        raise ValueError(f"node_sample_size {node_sample_size} exceeds the instance ({n_customers} customers + depot)")
    if vehicle_sample_size > vehicles.shape[0]:
        raise ValueError(f"vehicle_sample_size {vehicle_sample_size} exceeds the instance's {vehicles.shape[0]} vehicles")
    F = nodes.shape[1]
    batch = np.zeros((batch_size, node_sample_size + vehicle_sample_size, F), dtype=np.float32)
    batch[:, 0] = eos[0]
    for b in range(batch_size):
        picks = rng.permutation(n_customers)[: node_sample_size - 1] + 1
        batch[b, 1:node_sample_size] = nodes[picks]
        batch[b, node_sample_size:] = vehicles[rng.permutation(vehicles.shape[0])[:vehicle_sample_size]]
    return batch

"""CPU tests of the measurement helpers: the modelled HBM bytes per state that bench.py reports
(`config.hbm_bytes_per_state`) and the reducer of ncu launch lists behind profiles/*_launch_shares_*.csv."""
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_dataflow_bytes_follow_the_kernel_decomposition():
    import bench
    u = 128 * 4
    N, R = 51, 100
    kept = bench_kept(R)

    def enc(tokens, layers, dff):
        # per token-layer: forward 14 + 2m row units; backward 30 + 4m (LayerNorm 1 backward rides in the d(out1) GEMM)
        # minus 2 for every layer that is not the top of its stack (LayerNorm 2 backward rides in the dx GEMM above it)
        m = dff // 128
        return tokens * layers * (14 + 2 * m + 30 + 4 * m - 2 * (layers - 1) / layers) * u
    critic = enc(N + R + N + R, 3, 512)
    actor = enc(N + kept + N + kept, 1, 128) + (N + kept) * 12 * u
    full, rollout = bench.dataflow_bytes_per_state(N, R)
    assert full == int(critic + actor)
    # the figures quoted in DESIGN.md section 5
    assert 36.5e6 < full < 37.5e6 and 1.8e6 < rollout < 1.95e6
    # a one-layer critic has no fused LayerNorm-2 backward: 2 row units per token more than a layer below another one
    one, _ = bench.dataflow_bytes_per_state(N, R, critic=(1, 512))
    assert one == int(enc(2 * (N + R), 1, 512) + actor)


def bench_kept(R, B=1024, chunk=10240):
    """rule tokens an actor update chunk keeps on average (token windows), as bench.py counts them"""
    T = R
    n_chunks = max(1, -(-T * B // chunk))
    return sum(R - min(R - 1, (k * chunk) // B) for k in range(n_chunks)) / n_chunks


def test_launch_shares_reducer(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests", "gpu_checks"))
    import launch_shares
    hdr = ["ID", "Process ID", "Process Name", "Host Name", "Kernel Name", "Context", "Stream", "Block Size", "Grid Size",
           "Device", "CC", "Section Name", "Metric Name", "Metric Unit", "Metric Value"]

    def row(i, name, ns):
        return [str(i), "1", "python", "h", name, "1", "7", "(256, 1, 1)", "(8, 1, 1)", "0", "10.0", "s",
                "gpu__time_duration.sum", "ns", str(ns)]
    raw = tmp_path / "raw.csv"
    with open(raw, "w", newline="") as f:
        f.write("==PROF== Connected to process 1\n")
        w = csv.writer(f, quoting=csv.QUOTE_ALL)
        w.writerow(hdr)
        w.writerow(row(0, "tpc::<unnamed>::reset_kernel(int)", 1000))            # a short first segment (ignored)
        w.writerow(row(1, "tpc::<unnamed>::reset_kernel(int)", 1000))
        w.writerow(row(2, "void tpc::<unnamed>::gemm_tc_kernel<0>(CUtensorMap_st)", 3000))
        w.writerow(row(3, "void tpc::<unnamed>::gemm_tc_kernel<0>(CUtensorMap_st)", 5000))
        w.writerow(row(4, "tpc::<unnamed>::wgrad_tc_kernel(CUtensorMap_st)", 1000))
    out_l, out_s = tmp_path / "l.csv", tmp_path / "s.csv"
    launch_shares.main(str(raw), str(out_l), str(out_s))
    lst = list(csv.reader(open(out_l)))
    assert lst[0] == ["id", "kernel", "grid", "block", "duration_us"] and len(lst) == 6
    shares = {r[0]: r for r in list(csv.reader(open(out_s)))[1:]}
    g = shares["void tpc::<unnamed>::gemm_tc_kernel<0>"]
    assert g[1] == "2" and float(g[2]) == 8.0 and float(g[4]) == 80.0     # 8 of the step's 10 us
    assert float(shares["tpc::<unnamed>::reset_kernel"][4]) == 10.0

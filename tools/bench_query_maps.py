"""Host-only timing of make_query_map over a batch (the e2e-only `query_maps` stage of bench.py): no GPU needed.
python tools/bench_query_maps.py [--batch 1024] [--structures 2000] [--threads 0] [--repeats 20]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--structures", type=int, default=2000)
    ap.add_argument("--repeats", type=int, default=20)
    a = ap.parse_args()
    import bench
    from folddisco_b200 import host, synth
    db = synth.generate(a.structures, synth.SEED_BASE + 2)
    inputs = host.QueryInputs(*bench.query_inputs(db, a.batch, 0))
    best, tot = 1e9, 0.0
    for r in range(a.repeats + 3):
        t0 = time.perf_counter()
        qb = host.QueryBatch(host.HashParams())
        qb.add_prepared(inputs)
        dt = (time.perf_counter() - t0) * 1e3
        if r >= 3:
            best = min(best, dt)
            tot += dt
        del qb
    print("query_maps: batch %d, best %.3f ms, mean %.3f ms (%d host threads)" % (a.batch, best, tot / a.repeats, os.cpu_count()))


if __name__ == "__main__":
    main()

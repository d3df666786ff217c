"""One host core of the CPU walker baseline of bench.py -- TEST / BENCH INFRASTRUCTURE (see oracle.py's header).

    python oracle/walk_baseline.py <graph.npy> <lo> <hi> <n_hops> <alpha> <T> <sync_dir>

Loads the CSR (memory-mapped), warms up, waits until every worker is ready (a file barrier in <sync_dir>), then runs
the oracle's restatement of do_random_walks + sample_neighborhood_topt (pinsage_model.py:32-53, 88-107) on the sources
[lo, hi) and prints the seconds it took."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    path, lo, hi, n_hops, alpha, T, sync_dir = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5]), int(sys.argv[6]), sys.argv[7]
    from oracle import oracle
    z = np.load(path, mmap_mode="r")
    n = int(z[0])
    indptr, indices = np.asarray(z[1:n + 2]), np.asarray(z[n + 2:])
    oracle.topt_from_trace(oracle.do_random_walks_philox(indptr, indices, np.arange(lo, min(hi, lo + 32)), 8, alpha, 7), np.arange(lo, min(hi, lo + 32)), 4)
    open(os.path.join(sync_dir, f"ready.{lo}"), "w").close()
    t_wait = time.time()
    while not os.path.exists(os.path.join(sync_dir, "go")):
        if time.time() - t_wait > 300:
            sys.exit("no go signal")
        time.sleep(0.005)
    src = np.arange(lo, hi)
    t0 = time.perf_counter()
    trace = oracle.do_random_walks_philox(indptr, indices, src, n_hops, alpha, 7)
    oracle.topt_from_trace(trace, src, T)
    print(time.perf_counter() - t0, flush=True)


if __name__ == "__main__":
    main()
